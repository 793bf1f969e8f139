"""Yue: the data-loading / splitting dispatcher and recommender loader of the reference's
yue.py:10-135, for running the GPU recommenders without the reference tree.

    from yue_b200.host.config import Config
    from yue_b200.host.driver import Yue
    Yue(Config('config/BPR.conf')).execute()

`recommender=<Name>` is resolved in yue_b200's registry (BPR, APR, WRMF) instead of by importing
recommender.{baseline,cf,advanced}.<Name>; `-cv k` folds run one after the other on the same GPU
unless `-p` is given, in which case fold i uses CUDA device i % device_count (the reference runs
them as processes and divides MKL threads, yue.py:72-105).
"""
import os
from multiprocessing import get_context
from time import localtime, strftime, time

from .config import LineConfig
from .fileio import DataSplit, FileIO


def _registry():
    from ..bpr import BPR
    reg = {'BPR': BPR}
    try:
        from ..apr import APR
        reg['APR'] = APR
    except ImportError:
        pass
    from ..wrmf import WRMF
    reg['WRMF'] = WRMF
    return reg


def _run_fold(queue, cls_name, config, train, test, fold, order, device):
    os.environ['YUE_DEVICE'] = str(device)
    queue.put((order, _registry()[cls_name](config, train, test, fold).execute()))


class Yue(object):
    def __init__(self, config):
        self.trainingData, self.testData, self.measure = [], [], []
        self.config = config
        setup = LineConfig(config['record.setup'])
        columns = {}
        for col in setup['-columns'].split(','):
            name, pos = col.split(':')
            columns[name] = int(pos)
        delim = setup['-delim'] if setup.contains('-delim') else ''
        if not self.config.contains('evaluation.setup'):
            print('Evaluation is not well configured!')
            exit(-1)
        self.evaluation = LineConfig(config['evaluation.setup'])
        binarized = self.evaluation.contains('-b')
        bottom = float(self.evaluation['-b']) if binarized else 0

        def load(path):
            return FileIO.loadDataSet(path, columns=columns, binarized=binarized, threshold=bottom, delim=delim)
        self._array_folds = None
        if self.config.contains('yue.ingest') and self.config['yue.ingest'] == 'arrays':
            # (-b only rewrites the value of a `play` field, tool/file.py:40-44; no model on this path reads that field)
            # the log as numbered events, without a Python object per event (yue_b200/ingest.py)
            from ..ingest import cv_folds, load_numbered
            target = self.evaluation['-target'] if self.evaluation.contains('-target') else 'track'
            pure_cv = self.evaluation.contains('-cv') and not any(self.evaluation.contains(o) for o in ('-testSet', '-ap', '-byTime'))
            if pure_cv:                                 # every fold numbered from the coded columns of ONE read of the file
                self._array_folds = lambda k: ((log, []) for log in cv_folds(config['record'], columns, delim, k, target))
            if pure_cv or not self.evaluation.contains('-cv'):
                if not pure_cv:
                    self.trainingData = load_numbered(config['record'], columns, delim, self.evaluation, target)
                print('preprocessing...')
                return
            # -cv on top of another split: the dict form below (execute() folds whatever list the split left)
        if self.evaluation.contains('-testSet'):
            self.trainingData = load(config['record'])
            self.testData = load(self.evaluation['-testSet'])
        elif self.evaluation.contains('-ap'):
            self.trainingData, self.testData = DataSplit.dataSplit(load(config['record']),
                                                                   test_ratio=float(self.evaluation['-ap']))
        elif self.evaluation.contains('-byTime') or self.evaluation.contains('-cv'):
            self.trainingData = load(config['record'])
        print('preprocessing...')

    def execute(self):
        name = self.config['recommender']
        reg = _registry()
        if name not in reg:
            raise ImportError('yue_b200 accelerates %s; %s is not on the hot path' % (sorted(reg), name))
        if not self.evaluation.contains('-cv'):
            return reg[name](self.config, self.trainingData, self.testData).execute()
        k = int(self.evaluation['-cv'])
        if k <= 1 or k > 10:
            k = 3
        parallel = self.evaluation.contains('-p')
        ctx = get_context('spawn')                 # CUDA contexts do not survive fork
        queue = ctx.Queue()
        from ..engine import device_count
        ndev = max(1, device_count())
        procs = []
        folds = self._array_folds(k) if getattr(self, '_array_folds', None) else DataSplit.crossValidation(self.trainingData, k)
        for i, (train, test) in enumerate(folds, 1):
            p = ctx.Process(target=_run_fold, args=(queue, name, self.config, train, test, '[' + str(i) + ']', i,
                                                    (i - 1) % ndev if parallel else 0))
            p.start()
            procs.append(p)
            if not parallel:
                p.join()
        results = dict(queue.get() for _ in procs)
        for p in procs:
            p.join()
        self.measure = [results[i] for i in range(1, k + 1)]
        res = []
        for i, line in enumerate(self.measure[0]):
            if line[:3] == 'Top':
                res.append(line)
                continue
            label = line.split(':')[0]
            res.append(label + ':' + str(sum(float(m[i].split(':')[1]) for m in self.measure) / k) + '\n')
        currentTime = strftime("%Y-%m-%d %H-%M-%S", localtime(time()))
        outDir = LineConfig(self.config['output.setup'])['-dir']
        FileIO.writeFile(outDir, name + '@' + currentTime + '-' + str(k) + '-fold-cv' + '.txt', res)
        print('The result of %d-fold cross validation:\n%s' % (k, ''.join(res)))
        return res
