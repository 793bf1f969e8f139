"""BPR on the GPU behind the reference's recommender class API.

``GpuBPRMixin`` carries the four hot-path methods -- ``initModel`` / ``buildModel`` / ``predict`` /
``evalRanking`` (+ the ``ranking_performance`` progress hook) -- written against the attributes
the reference's base classes provide (``self.data`` = Record, ``self.k``, ``self.maxIter``,
``self.lRate``, ``self.regU`` ...).  It is combined with

* ``yue_b200.host.recommender.IterativeRecommender`` here (``class BPR`` below), so the package
  runs stand-alone, and
* the reference's own ``base.IterativeRecommender`` in the drop-in shim
  ``dropin/recommender/cf/BPR.py``.

What replaces what (reference paths):
  buildModel   recommender/cf/BPR.py:31-62 -- the per-triplet SGD loop (numpy text; the shipped
               TF/Adam variant at 83-129 is documented in DESIGN.md and not reproduced)
  predict      recommender/cf/BPR.py:131-134
  evalRanking  base/IterativeRecommender.py:77-173, with the EXACT masked top-N instead of the
               lossy selection at 107-145
No CUDA context is created before initModel: under ``-cv`` the object is built in the parent and
executed in a child process (yue.py:92-105).
"""
import os
import random
from os.path import abspath
from time import localtime, strftime, time

import numpy as np

from .engine import MODE_HOGWILD, MODE_SERIAL, RANK_AUTO, Engine
from .host.config import LineConfig
from .host.fileio import FileIO
from .host.measure import Measure


class GpuBPRMixin(object):
    #: optional config keys understood on top of the reference's (all have defaults)
    #:   yue.device=<int>      CUDA device (default $YUE_DEVICE or 0)
    #:   yue.sgd=hogwild|serial  update schedule (default hogwild; serial = reference order)
    #:   yue.seed=<int>        sampler seed (default: drawn from `random`, unseeded like the reference)
    #:   yue.ingest=host|device   where the array form of the log (event CSR, play sets, test sets) is built: host =
    #:                         numpy on Record's dicts; device = yue_ingest_events on the numbered events (K0)
    #:   yue.devices=0,1,...   several CUDA devices, driven by threads of this process: evalRanking ranks a block of the test
    #:                         users on each; buildModel (Hogwild, num.factors 32/64/128) shards the users over them with the
    #:                         hot rows shared over peer memory (yue_b200/sharding.py: SharedHotTrainer).  The first one is
    #:                         the device everything else runs on.  yue.sub_epochs (32 up to 2 devices, 4 N^2 above) / yue.asynchrony (0.25) tune it.
    #:   yue.metrics=host|device  where evalRanking computes Precision/Recall/F1/MAP/Coverage (default host:
    #:                         the reference's own summation order; device = one kernel over the lists that are
    #:                         already on the GPU, for test sets where the Python set operations dominate)
    _engine = None

    # ---- plumbing --------------------------------------------------------------------------
    def _opt(self, key, default):
        return self.config[key] if self.config.contains(key) else default

    def _devices(self):
        v = self._opt('yue.devices', None)
        if v is None:
            return [int(self._opt('yue.device', os.environ.get('YUE_DEVICE', '0')))]
        devs = [int(x) for x in str(v).replace(' ', '').split(',') if x != '']
        if not devs:
            raise ValueError('yue.devices is empty')
        return devs

    def _get_engine(self):
        if self._engine is None:
            dev = self._devices()[0]
            self._engine = Engine(dev)
            arrays = getattr(self.data, 'interaction_arrays', None)
            self._test_on_device = False
            if hasattr(self.data, 'log'):                       # yue.ingest=arrays (ingest.ArrayRecord): numbered events -> K0
                self.data.log.upload(self._engine)
                self.data.test_indptr, self.data.test_items = self._engine.get_test_set()
                ev = LineConfig(self.config['evaluation.setup'])
                if ev.contains('-cold') or ev.contains('-sample'):         # base/recommender.py:22-49 on the CSR form
                    from .ingest import filter_test_rows
                    self.data.test_indptr, self.data.test_items = filter_test_rows(
                        self.data.log, self.data.test_indptr, self.data.test_items,
                        cold=int(ev['-cold']) if ev.contains('-cold') else None, sample=ev.contains('-sample'))
                    self._engine.set_test_set(self.data.test_indptr, self.data.test_items)
                self._test_on_device = True
                self._synced = (None, None)
                return self._engine
            if self._opt('yue.ingest', 'host') == 'device':
                # Record's containers as numbered events: every training event per user in file order
                # (userRecord, data/record.py:147-150), then the held-out pairs (testSet, 182-192)
                uid, tid = self.data.name2id['user'], self.data.name2id[self.recType]
                tr_u = [uid[u] for u, evs in self.data.userRecord.items() for _ in evs]
                tr_i = [tid[e[self.recType]] for evs in self.data.userRecord.values() for e in evs]
                te_u = [uid[u] for u, held in self.data.testSet.items() for _ in held]
                te_i = [tid[t] for held in self.data.testSet.values() for t in held]
                ev_user = np.array(tr_u + te_u, dtype=np.int32)
                ev_item = np.array(tr_i + te_i, dtype=np.int32)
                flag = np.zeros(len(ev_user), dtype=np.uint8)
                flag[len(tr_u):] = 1
                self._engine.ingest_events(self.m, self.n, ev_user, ev_item, flag)
                self._test_on_device = True
                self._synced = (None, None)
                return self._engine
            if arrays is not None:
                ev_indptr, ev_items, uq_indptr, uq_items = arrays(self.recType)
            else:                       # the reference's Record: build the arrays from its dicts
                from .host.record import interaction_arrays
                ev_indptr, ev_items, uq_indptr, uq_items = interaction_arrays(
                    self.data.name2id, self.data.userRecord, self.recType)
            self._engine.set_interactions(self.m, self.n, ev_indptr, ev_items, uq_indptr, uq_items)
            self._synced = (None, None)
        return self._engine

    #: names of the attributes that hold the user / track factor tables (WRMF calls them X / Y, WRMF.py:19-20)
    _user_table, _item_table = 'P', 'Q'

    def _push_factors(self):
        """Upload the host tables when they are not the arrays last synchronised."""
        eng = self._get_engine()
        P, Q = getattr(self, self._user_table), getattr(self, self._item_table)
        if self._synced[0] is not P or self._synced[1] is not Q:
            P = np.ascontiguousarray(P, dtype=np.float32)
            Q = np.ascontiguousarray(Q, dtype=np.float32)
            setattr(self, self._user_table, P)
            setattr(self, self._item_table, Q)
            eng.set_factors(P, Q)
            self._synced = (P, Q)
        return eng

    def _pull_factors(self):
        P, Q = self._engine.get_factors()
        setattr(self, self._user_table, P)
        setattr(self, self._item_table, Q)
        self._synced = (P, Q)

    # ---- the hot path ----------------------------------------------------------------------
    def initModel(self):
        super(GpuBPRMixin, self).initModel()             # P, Q ~ U[0,0.1) float32, loss = lastLoss = 0
        self.m = self.data.getSize('user')
        self.n = self.data.getSize(self.recType)
        self.train_size = len(self.data.trainingData)

    def buildModel(self):
        print('training...')
        eng = self._push_factors()
        mode = MODE_SERIAL if self._opt('yue.sgd', 'hogwild') == 'serial' else MODE_HOGWILD
        seed = int(self._opt('yue.seed', random.getrandbits(63)))
        devs = self._sharded_devices(mode)
        if devs:
            return self._build_sharded(eng, devs, seed)
        iteration = 0
        while iteration < self.maxIter:
            loss = eng.bpr_epoch(self.lRate, self.regU, self.regI, seed, iteration, mode)
            p2, q2 = eng.frob2()
            self.loss = loss + self.regU * p2 + self.regI * q2
            iteration += 1
            if self.isConverged(iteration):
                break
        self._pull_factors()

    def _sharded_session(self, eng, devs):
        """buildModel over several GPUs.  BPR.py:40-62 with the users dealt out to the devices (user u on devs[u % N], so
        that all devices walk the reference's user stream together), the same negatives as one GPU would draw (the sampler
        is a function of the global event index) and the schedule of sharding.SharedHotTrainer; one handle per device,
        each driven by its own thread.  Context manager: yields run_epoch(seed, iteration, apr=None) -> (sum of the ranks'
        -log losses, |P|^2, |Q|^2); on exit the trained tables are gathered into the host attributes."""
        import contextlib
        from . import sharding

        @contextlib.contextmanager
        def session():
            world = len(devs)
            ev_indptr, ev_items, uq_indptr, uq_items = eng.get_interactions()
            P, Q = getattr(self, self._user_table), getattr(self, self._item_table)
            shared = sharding.ThreadCtl.Shared(world)
            S, A = int(self._opt('yue.sub_epochs', 0)), float(self._opt('yue.asynchrony', 0))   # 0: sharding.default_sub_epochs / default_asynchrony

            def start(r, ctl):
                mine = sharding.interleaved_users(self.m, world, r)
                sh = sharding.local_shard_of_users(ev_indptr, ev_items, uq_indptr, uq_items, mine)
                e = Engine(devs[r])
                e.set_interactions(sh['m_local'], self.n, sh['ev_indptr'], sh['ev_items'], sh['uq_indptr'], sh['uq_items'])
                e.set_event_offsets(sh['event_offsets'])
                e.set_factors(np.ascontiguousarray(P[mine]), Q)
                tr = sharding.SharedHotTrainer(e, ctl, np.bincount(sh['ev_items'], minlength=self.n), sub_epochs=S, asynchrony=A)
                return mine, e, tr
            ranks = sharding.run_on_ranks(shared, start)

            def run_epoch(seed, iteration, apr=None):
                def epoch(r, ctl):
                    _, e, tr = ranks[r]
                    loss = tr.epoch(self.lRate, self.regU, self.regI, seed, iteration, want_loss=True, apr=apr, finalize=True)
                    p2, q2 = e.frob2()
                    return loss, p2, q2
                res = sharding.run_on_ranks(shared, epoch)
                return sum(x[0] for x in res), sum(x[1] for x in res), res[0][2]
            try:
                yield run_epoch

                def finish(r, ctl):
                    _, e, tr = ranks[r]
                    tr.close()
                    return e.get_factors()
                tables = sharding.run_on_ranks(shared, finish)
            finally:
                for _, e, _ in ranks:
                    e.close()
            Pn = np.array(P, dtype=np.float32, copy=True)
            for (mine, _, _), (Pl, _) in zip(ranks, tables):
                Pn[mine] = Pl
            setattr(self, self._user_table, Pn)
            setattr(self, self._item_table, tables[0][1])
            self._synced = (None, None)                # the main handle gets the trained tables on its next use
        return session()

    def _sharded_devices(self, mode):
        """The devices buildModel trains on when it shards the users, else None."""
        devs = self._devices()
        if len(devs) > 1 and mode == MODE_HOGWILD:
            if self.k in (32, 64, 128):
                return devs
            print('yue.devices: num.factors=%d trains on device %d alone (the shared hot rows of the multi-GPU trainer need 32, 64 '
                  'or 128 factors); ranking uses all devices' % (self.k, devs[0]))
        return None

    def _build_sharded(self, eng, devs, seed):
        with self._sharded_session(eng, devs) as run_epoch:
            iteration = 0
            while iteration < self.maxIter:
                loss, p2, q2 = run_epoch(seed, iteration)
                self.loss = loss + self.regU * p2 + self.regI * q2
                iteration += 1
                if self.isConverged(iteration):
                    break

    def predict(self, u):
        'invoked to rank all the items for the user'
        return self._push_factors().predict(self.data.getId(u, 'user'))

    def _topn_lists(self, users, N):
        """[user names] -> {user: [N track names]} through the fused score+mask+top-N kernel."""
        eng = self._push_factors()
        uid = np.array([self.data.getId(u, 'user') for u in users], dtype=np.int32)
        devs = self._devices()
        if len(devs) > 1 and len(uid) >= 2 * len(devs):
            ids, scores = self._rank_on_devices(eng, devs, uid, N)
        else:
            ids, scores = eng.rank_topn(uid, N, RANK_AUTO)
        id2name = self.data.id2name[self.recType]
        return {u: [id2name[int(t)] for t in row if t >= 0] for u, row in zip(users, ids)}, ids, scores

    _rank_engines = None

    def _rank_on_devices(self, eng, devs, uid, N):
        """Every device ranks a contiguous block of the users against its own copy of Q and of the play sets (SURVEY.md 8e:
        no collective, the host concatenates the lists).  Same kernels, same tables: the lists are those of one device."""
        from . import sharding
        if self._rank_engines is None:
            arrays = eng.get_interactions()
            self._rank_engines = [eng]
            for dv in devs[1:]:
                e = Engine(dv)
                e.set_interactions(self.m, self.n, *arrays)
                self._rank_engines.append(e)
            self._rank_synced = [None] * len(devs)
        P, Q = self._synced
        world = len(devs)
        bounds = [len(uid) * r // world for r in range(world + 1)]
        ids, scores = np.empty((len(uid), N), np.int32), np.empty((len(uid), N), np.float32)

        def rank(r, ctl):
            e = self._rank_engines[r]
            if r > 0 and self._rank_synced[r] is not P:
                e.set_factors(P, Q)
                self._rank_synced[r] = P
            lo, hi = bounds[r], bounds[r + 1]
            e.rank_topn(uid[lo:hi], N, RANK_AUTO, ids[lo:hi], scores[lo:hi])
        sharding.run_on_ranks(sharding.ThreadCtl.Shared(world), rank)
        return ids, scores

    def _eval_ranking_arrays(self, top, N):
        """evalRanking over ingest.ArrayRecord: every step on arrays -- ranking (K3/K5), the measures (K6) and the hit marks
        of the result lines (one sorted search); the only per-user Python left is joining a line's cells."""
        from .ingest import hit_mask, name_blob, ranking_measure, result_text
        eng = self._push_factors()
        data = self.data
        uid = np.flatnonzero(np.diff(data.test_indptr) > 0).astype(np.int32)
        devs = self._devices()
        if len(devs) > 1 and len(uid) >= 2 * len(devs):
            ids, _ = self._rank_on_devices(eng, devs, uid, N)
        else:
            ids, _ = eng.rank_topn(uid, N, RANK_AUTO)
        un, tn = data.log.names['user'], data.log.names[self.recType]
        hits = hit_mask(uid, ids, self.n, data.test_indptr, data.test_items)
        if getattr(data, '_track_blob', None) is None:
            data._track_blob = name_blob(tn)                        # the catalog's names as one blob, once per Record
        ub, uo = name_blob(un[uid])
        res = ['userId: recommendations in (itemId, ranking score) pairs, * means the item matches.\n',
               result_text(ub, uo, data._track_blob[0], data._track_blob[1], ids, hits)]
        self.rec_ids, self.rec_users = ids, uid
        if self._opt('yue.metrics', 'device') == 'device' and len(devs) == 1:
            measure, ndcg = self._format_device_measure(uid, top)      # K6 over the lists the call left on the device
        else:
            measure, ndcg = ranking_measure(ids, hits, np.diff(data.test_indptr)[uid], top, self.n)
        return res, measure, ndcg

    def evalRanking(self):
        top = [int(num) for num in self.ranking['-topN'].split(',')]
        N = max(top)
        if N > 100 or N < 0:
            print('N can not be larger than 100! It has been reassigned with 10')
            N = 10
        if hasattr(self.data, 'log'):
            res, self.measure, self.ndcg = self._eval_ranking_arrays(top, N)
            currentTime = strftime("%Y-%m-%d %H-%M-%S", localtime(time()))
            outDir = self.output['-dir']
            if self.isOutput:
                FileIO.writeFile(outDir, self.config['recommender'] + '@' + currentTime + '-top-' + self.ranking['-topN'] + 'items'
                                 + self.foldInfo + '.txt', res)
                print('The result has been output to ', abspath(outDir), '.')
            FileIO.writeFile(outDir, self.config['recommender'] + '@' + currentTime + '-measure' + self.foldInfo + '.txt', self.measure)
            print('The result of %s %s:\n%s' % (self.algorName, self.foldInfo, ''.join(self.measure)))
            return
        users = list(self.data.testSet.keys())
        recList, _, _ = self._topn_lists(users, N)
        res = ['userId: recommendations in (itemId, ranking score) pairs, * means the item matches.\n']
        for i, user in enumerate(users):
            if i % 100 == 0:
                print(self.algorName, self.foldInfo, 'progress:' + str(i) + '/' + str(len(users)))
            held = self.data.testSet[user]
            res.append(user + ':' + ''.join(item + ('*' if item in held else '') for item in recList[user]) + '\n')
        currentTime = strftime("%Y-%m-%d %H-%M-%S", localtime(time()))
        outDir = self.output['-dir']
        if self.isOutput:
            fileName = ''
            if self.ranking.contains('-topN'):
                fileName = self.config['recommender'] + '@' + currentTime + '-top-' + self.ranking['-topN'] \
                    + 'items' + self.foldInfo + '.txt'
            FileIO.writeFile(outDir, fileName, res)
            print('The result has been output to ', abspath(outDir), '.')
        fileName = self.config['recommender'] + '@' + currentTime + '-measure' + self.foldInfo + '.txt'
        self.recList = recList
        if self._opt('yue.metrics', 'host') == 'device':
            self.measure, self.ndcg = self._device_measure(users, top)
        else:
            self.measure = Measure.rankingMeasure(self.data.testSet, recList, top, self.data.getSize(self.recType))
            self.ndcg = {n: Measure.NDCG(self.data.testSet, recList, n) for n in top}
        FileIO.writeFile(outDir, fileName, self.measure)
        print('The result of %s %s:\n%s' % (self.algorName, self.foldInfo, ''.join(self.measure)))

    def _device_measure(self, users, top):
        """Measure.rankingMeasure's list of strings (evaluation/measure.py:16-41) from the device-side sums
        over the lists the last rank_topn call left on the GPU."""
        eng = self._engine
        if getattr(self, '_test_on_device', False):
            return self._format_device_measure(users, top)
        getId = self.data.getId
        rows = [sorted(set(getId(t, self.recType) for t in self.data.testSet[u])) for u in self.data.name2id['user']
                if u in self.data.testSet]
        by_user = dict(zip((u for u in self.data.name2id['user'] if u in self.data.testSet), rows))
        indptr = np.zeros(self.m + 1, dtype=np.int64)
        for u, r in by_user.items():
            indptr[getId(u, 'user') + 1] = len(r)
        np.cumsum(indptr, out=indptr)
        items = np.zeros(int(indptr[-1]), dtype=np.int32)
        for u, r in by_user.items():
            items[indptr[getId(u, 'user')]:indptr[getId(u, 'user') + 1]] = r
        eng.set_test_set(indptr, items)
        return self._format_device_measure(users, top)

    def _format_device_measure(self, users, top):
        eng = self._engine
        sums, distinct = eng.rank_metrics(top)
        B, itemCount = len(users), self.data.getSize(self.recType)
        print('rank measure...')
        measure, ndcg = [], {}
        for k, n in enumerate(top):
            prec, recall = float(sums[k, 0]) / (B * n), float(sums[k, 1]) / B
            measure += ['Top ' + str(n) + '\n', 'Precision:' + str(prec) + '\n', 'Recall:' + str(recall) + '\n',
                        'F1:' + str(Measure.F1(prec, recall)) + '\n', 'MAP:' + str(float(sums[k, 2]) / B) + '\n',
                        'Coverage:' + str(int(distinct[k]) / float(itemCount)) + '\n']
            ndcg[n] = float(sums[k, 3]) / B
        return measure, ndcg

    def ranking_performance(self):
        """Top-10 on the first 300 test users (IterativeRecommender.py:175-235), masking the
        TRAINING tracks -- the reference masks the test tracks there, which zeroes its own hit
        counts; the hook is kept, the defect is not."""
        sample = {}
        itemcount = 0
        for user in self.data.testSet:
            itemcount += len(self.data.testSet[user])
            if len(sample) == 300:
                break
            sample[user] = self.data.testSet[user]
        recList, _, _ = self._topn_lists(list(sample.keys()), 10)
        measure = Measure.rankingMeasure(sample, recList, [10], itemcount)
        print('-' * 80)
        print('Ranking Performance ' + self.foldInfo + ' (Top-10 On 300 sampled users)')
        for m in measure[1:]:
            print(m.strip())
        print('-' * 80)
        return measure


from .host.recommender import IterativeRecommender  # noqa: E402


class BPR(GpuBPRMixin, IterativeRecommender):
    """BPR: Bayesian Personalized Ranking from Implicit Feedback (Rendle et al.), GPU hot path."""

    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(BPR, self).__init__(conf, trainingSet, testSet, fold)
