"""The steps either side of the hot path without per-event Python.

Before: log file -> numbered events.  The reference reads the log line by line into a list of dicts
(tool/file.py:23-52), splits it with one `random()` per event (tool/dataSplit.py:9-23) and hands out ids by first appearance
while walking those dicts (data/record.py:138-146, 182-188) -- microseconds of interpreter per event, minutes at config
C2's 50 M events.  Here the file goes through a C parser into columns, ids are first-appearance codes of whole columns
(`pandas.factorize`: the same order, training events first, test events extending the maps), and the numbered events go to
the device, where K0 (yue_ingest_events) builds the event CSR, the play sets and the test sets.

After: ranked ids -> result lines.  IterativeRecommender.py:145-155 builds `user:` + track names, `*` after a hit, one
Python loop per user and per item; here the hits are one sorted search over (user, track) keys and the lines are built
column-wise.

`ArrayRecord` gives the recommender classes the part of the reference's Record interface they use (name2id / id2name /
getId / getSize / testSet) on top of the arrays; the dict views are built only when somebody asks for them.
"""
import re

import numpy as np


class ArrayLog(object):
    """Numbered events of one log: ids by first appearance (training events first), file order kept."""

    def __init__(self, ev_user, ev_item, is_test, names, rec_type):
        self.ev_user, self.ev_item, self.is_test = ev_user, ev_item, is_test
        self.names, self.rec_type = names, rec_type            # names[kind] = array of names, index = id
        self.file_pos = None                                   # -byTime: position of every event in the file (the split reorders them)
        self.m, self.n = len(names['user']), len(names[rec_type])

    @property
    def train_size(self):
        return int(len(self.is_test) - int(self.is_test.sum()))

    def upload(self, engine):
        engine.ingest_events(self.m, self.n, self.ev_user, self.ev_item, self.is_test)


def read_columns(path, columns, delim=''):
    """The named columns of a log file as arrays of strings, in file order (tool/file.py:23-40: fields split on `delim`,
    default comma / blank / tab; `columns` = {name: field index} in record.setup's -columns order)."""
    import pandas as pd
    names = list(columns.keys())
    where = [int(columns[k]) for k in names]
    if len(names) < 2:
        print('The dataset needs more information or the record.setup setting has some problems...')
        exit(-1)
    sep = delim if delim != '' else ',| |\t'
    single = len(sep) == 1 and not re.escape(sep) != sep
    try:
        df = pd.read_csv(path, sep=sep, header=None, usecols=sorted(set(where)), dtype=str, engine='c' if single else 'python',
                         keep_default_na=False, skipinitialspace=False, quoting=3)
    except (ValueError, IndexError):
        print('The record file is not in a correct format.')
        exit(-1)
    return {name: df[ind].to_numpy(dtype=object) for name, ind in zip(names, where)}


def global_randoms(n):
    """n values of `random.random()` drawn from -- and advancing -- CPython's global Mersenne Twister, as one array: the
    generator's state is handed to numpy's MT19937 (the same recurrence and tempering), 2 n words are drawn and combined the
    way `random.random()` does ((a >> 5) * 2^26 + (b >> 6)) / 2^53, and the advanced state is written back, so whatever draws
    from `random` next continues the same stream.  Bit-identical to the loop (tests/test_ingest_arrays.py)."""
    import random
    ver, st, gauss = random.getstate()
    bg = np.random.MT19937()
    bg.state = {'bit_generator': 'MT19937', 'state': {'key': np.array(st[:624], dtype=np.uint32), 'pos': int(st[624])}}
    raw = bg.random_raw(2 * int(n))
    a = (raw[0::2] >> np.uint64(5)).astype(np.float64)
    b = (raw[1::2] >> np.uint64(6)).astype(np.float64)
    ns = bg.state['state']
    random.setstate((ver, tuple(int(x) for x in ns['key']) + (int(ns['pos']),), gauss))
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


def split_ap(count, test_ratio):
    """tool/dataSplit.py:9-23: event e goes to the test set when random() < ratio -- the same global `random` stream, one
    draw per event in file order, so a seeded run splits exactly like the reference's."""
    if test_ratio >= 1 or test_ratio <= 0:
        test_ratio = 0.3
    return global_randoms(count) < test_ratio


def number_events(train, test, rec_type='track', key_order=None):
    """train / test: {column name: array of strings} (test may be None).  Ids per column by first appearance over the
    training events, then the test events (data/record.py:138-146, 182-188).  Returns an ArrayLog whose events are the
    training events in order followed by the test events."""
    import pandas as pd
    names = {}
    codes = {}
    nt = len(train['user'])
    for kind in (key_order or train.keys()):
        if kind == 'time':
            continue
        col = train[kind] if test is None else np.concatenate([train[kind], test[kind]])
        c, uniq = pd.factorize(col, sort=False)
        codes[kind], names[kind] = c.astype(np.int32), np.asarray(uniq, dtype=object)
    is_test = np.zeros(len(codes['user']), dtype=np.uint8)
    is_test[nt:] = 1
    return ArrayLog(codes['user'], codes[rec_type], is_test, names, rec_type)


def by_time(cols, ratio):
    """data/record.py:108-123: per user (in order of first appearance) the events sorted by their `time` field AS STRINGS
    (the loader keeps fields as text), the first int(len * (1 - ratio)) for training, the rest for test.  Returns the two
    index arrays into the file's events, each in the reference's order (user by user)."""
    import pandas as pd
    ucode, _ = pd.factorize(cols['user'], sort=False)
    order = np.lexsort((np.arange(len(ucode)), cols['time'].astype(str), ucode))       # user, then time string, then file order
    u_sorted = ucode[order]
    start = np.flatnonzero(np.r_[True, u_sorted[1:] != u_sorted[:-1]])
    length = np.diff(np.r_[start, len(order)])
    rank_in_user = np.arange(len(order)) - np.repeat(start, length)
    cut = (length * (1 - ratio)).astype(np.int64)
    is_train = rank_in_user < np.repeat(cut, length)
    return order[is_train], order[~is_train]


def by_time_coded(ucodes, times, ratio):
    """by_time on a coded user column and the time column as an Arrow string array: the same two index arrays, without a
    Python string per event.  Users in order of first appearance in the file, inside a user by the time field AS A STRING,
    ties in file order (`sorted` is stable).  When every time field is a run of digits of one length (epoch seconds) the
    string order is the numeric order and the sort is a numpy lexsort; otherwise Arrow sorts the strings (byte order of
    UTF-8 = code-point order = Python's str order)."""
    import pyarrow as pa
    import pyarrow.compute as pcc
    n = len(ucodes)
    first = np.full(int(ucodes.max()) + 1 if n else 0, n, dtype=np.int64)
    first[ucodes[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)
    rank = np.empty(len(first), dtype=np.int64)
    rank[np.argsort(first, kind='stable')] = np.arange(len(first))
    u = rank[ucodes]                                               # user codes numbered by first appearance
    lens = pcc.min_max(pcc.utf8_length(times)).as_py() if n else {'min': 0, 'max': 0}
    if n and lens['min'] == lens['max'] and 0 < lens['max'] <= 18 and pcc.all(pcc.utf8_is_digit(times)).as_py():
        order = np.lexsort((pcc.cast(times, pa.int64()).to_numpy(zero_copy_only=False), u))      # stable: ties keep file order
    else:
        tbl = pa.table({'u': pa.array(u), 't': times, 'i': pa.array(np.arange(n, dtype=np.int64))})
        order = pcc.sort_indices(tbl, sort_keys=[('u', 'ascending'), ('t', 'ascending'), ('i', 'ascending')]).to_numpy(zero_copy_only=False)
    u_sorted = u[order]
    start = np.flatnonzero(np.r_[True, u_sorted[1:] != u_sorted[:-1]])
    length = np.diff(np.r_[start, len(order)])
    rank_in_user = np.arange(len(order)) - np.repeat(start, length)
    cut = (length * (1 - ratio)).astype(np.int64)
    is_train = rank_in_user < np.repeat(cut, length)
    return order[is_train], order[~is_train]


def read_coded(path, columns, delim, want_time=False):
    """The non-time columns of a log file as (codes int32[events], names[array of str]) per column WITHOUT creating a Python
    string per field: Arrow's multi-threaded CSV reader + dictionary encoding (codes in order of first appearance in the
    file).  None when that reader cannot take the file (a regex delimiter, pyarrow missing): the caller falls back to
    read_columns.  want_time: out['time'] is the time column as ONE Arrow string array (by_time_coded sorts it).
    `path` may be a list of files (-testSet: the log and the test file): one result per file, the names tables shared."""
    if len(delim) != 1 or (want_time and 'time' not in columns):
        return None
    try:
        import pyarrow as pa
        import pyarrow.csv as pc
    except ImportError:
        return None
    paths = [path] if isinstance(path, str) else list(path)
    used = sorted(set(int(v) for k, v in columns.items() if k != 'time'))
    # The columns are converted to dictionary type BY THE READER: every 8 MB block is parsed and encoded on its own thread
    # (encoding the finished columns afterwards is one thread per column: 1.0 s of the 2.4 s a 5 M-event file took), the
    # blocks' dictionaries are merged once per column.  Whatever order the merged dictionary has, number_coded re-assigns
    # the ids by first appearance.
    types = {'f%d' % i: pa.dictionary(pa.int32(), pa.string()) for i in used}
    if want_time:
        if int(columns['time']) in used:                           # the time field doubles as another column: object path
            return None
        types['f%d' % int(columns['time'])] = pa.string()
    tables = [pc.read_csv(p, read_options=pc.ReadOptions(autogenerate_column_names=True, block_size=8 << 20),
                          parse_options=pc.ParseOptions(delimiter=delim, quote_char=False),
                          convert_options=pc.ConvertOptions(include_columns=sorted(types), column_types=types, strings_can_be_null=False))
              for p in paths]
    tbl = tables[0] if len(tables) == 1 else pa.concat_tables(tables)
    bounds = np.cumsum([0] + [t.num_rows for t in tables])
    outs = [{} for _ in paths]
    if want_time:
        tm = tbl['f%d' % int(columns['time'])].combine_chunks()
        for k, o in enumerate(outs):
            o['time'] = tm.slice(int(bounds[k]), int(bounds[k + 1] - bounds[k]))
    for name, ind in columns.items():
        if name == 'time':
            continue
        d = tbl['f%d' % int(ind)].unify_dictionaries().combine_chunks()
        idx = d.indices                                            # int32, no nulls: read the buffer directly (to_numpy() pulls in pandas)
        codes = np.frombuffer(idx.buffers()[1], dtype=np.int32, count=len(idx), offset=idx.offset * 4).copy()
        names = np.asarray(d.dictionary.to_pylist(), dtype=object)
        for k, o in enumerate(outs):
            o[name] = (codes[bounds[k]:bounds[k + 1]], names)
    return outs[0] if isinstance(path, str) else outs


def number_coded(train, test, rec_type='track', key_order=None):
    """number_events on coded columns: train / test = {column: (codes, names)} with the SAME names table per column (codes of
    one dictionary).  Ids are re-assigned by first appearance over the training events, then the test events
    (data/record.py:138-146, 182-188) -- integer work only."""
    from concurrent.futures import ThreadPoolExecutor
    nt = len(train['user'][0])
    kinds = [k for k in (key_order or train.keys()) if k != 'time']

    def one(kind):
        c = train[kind][0] if test is None else np.concatenate([train[kind][0], test[kind][0]])
        table = train[kind][1]
        # where each code occurs first: written back to front, so that for a repeated code the LAST write -- its first
        # occurrence -- stays (numpy assigns repeated indices in order); no sort over the events
        first = np.full(len(table), len(c), dtype=np.int64)
        first[c[::-1]] = np.arange(len(c) - 1, -1, -1, dtype=np.int64)
        seen = np.flatnonzero(first < len(c))
        order = seen[np.argsort(first[seen], kind='stable')]        # the codes that occur, in order of first appearance
        remap = np.full(len(table), -1, dtype=np.int32)
        remap[order] = np.arange(len(order), dtype=np.int32)
        return remap[c], table[order]
    with ThreadPoolExecutor(max(1, len(kinds))) as pool:            # the columns are independent; numpy's indexing drops the GIL
        done = list(pool.map(one, kinds))
    codes = {k: d[0] for k, d in zip(kinds, done)}
    names = {k: d[1] for k, d in zip(kinds, done)}
    is_test = np.zeros(len(codes['user']), dtype=np.uint8)
    is_test[nt:] = 1
    return ArrayLog(codes['user'], codes[rec_type], is_test, names, rec_type)


def load_numbered(path, columns, delim, evaluation, rec_type='track'):
    """File -> ArrayLog under the reference's evaluation.setup options -ap r / -testSet file / -byTime r (yue.py:38-46)."""
    order = [k for k in columns.keys()]
    if not evaluation.contains('-byTime'):
        coded = read_coded(path, columns, delim)
        if coded is not None and evaluation.contains('-ap'):
            held = split_ap(len(coded['user'][0]), float(evaluation['-ap']))
            return number_coded({k: (c[~held], t) for k, (c, t) in coded.items()}, {k: (c[held], t) for k, (c, t) in coded.items()}, rec_type, order)
        if coded is not None and not evaluation.contains('-testSet'):
            return number_coded(coded, None, rec_type, order)
        if coded is not None:                                       # -testSet file: one names table per column over both files
            both = read_coded([path, evaluation['-testSet']], columns, delim)
            if both is not None:
                return number_coded(both[0], both[1], rec_type, order)
    elif not evaluation.contains('-testSet') and not evaluation.contains('-ap'):
        coded = read_coded(path, columns, delim, want_time=True)   # config/BPR.conf's own split
        if coded is not None:
            times = coded.pop('time')
            tr, te = by_time_coded(coded['user'][0], times, float(evaluation['-byTime']))
            log = number_coded({k: (c[tr], t) for k, (c, t) in coded.items()}, {k: (c[te], t) for k, (c, t) in coded.items()}, rec_type, order)
            log.file_pos = np.concatenate([tr, te])
            return log
    cols = read_columns(path, columns, delim)
    if evaluation.contains('-testSet'):
        return number_events(cols, read_columns(evaluation['-testSet'], columns, delim), rec_type, order)
    if evaluation.contains('-ap'):
        held = split_ap(len(cols['user']), float(evaluation['-ap']))
        return number_events({k: v[~held] for k, v in cols.items()}, {k: v[held] for k, v in cols.items()}, rec_type, order)
    if evaluation.contains('-byTime'):
        tr, te = by_time(cols, float(evaluation['-byTime']))
        log = number_events({k: v[tr] for k, v in cols.items()}, {k: v[te] for k, v in cols.items()}, rec_type, order)
        log.file_pos = np.concatenate([tr, te])
        return log
    return number_events(cols, None, rec_type, order)


def cv_folds(path, columns, delim, k, rec_type='track'):
    """-cv k (tool/dataSplit.py:26-38, yue.py:72-96): fold i holds out the events whose position in the file is i modulo k.
    Yields one ArrayLog per fold; the file is read once, every fold is numbered on its own (ids by first appearance over
    ITS training events, then its test events -- what Record does with the two lists the reference hands it)."""
    if k <= 1 or k > 10:
        k = 3
    order = [c for c in columns.keys()]
    coded = read_coded(path, columns, delim)
    cols = read_columns(path, columns, delim) if coded is None else None
    n = len(coded['user'][0]) if coded is not None else len(cols['user'])
    pos = np.arange(n)
    for i in range(k):
        held = pos % k == i
        if coded is not None:
            yield number_coded({c: (v[~held], t) for c, (v, t) in coded.items()}, {c: (v[held], t) for c, (v, t) in coded.items()}, rec_type, order)
        else:
            yield number_events({c: v[~held] for c, v in cols.items()}, {c: v[held] for c, v in cols.items()}, rec_type, order)


def filter_test_rows(log, test_indptr, test_items, cold=None, sample=False):
    """The two edits base/recommender.py:22-49 makes to the test set, on the CSR form (rows = users, sorted unique held-out
    tracks that are not training plays of the user -- what K0 / Record leave):
      -cold t   a held-out track with more than t TRAINING events is dropped, for users that have training events
                (recommender.py:22-42: len(trackRecord[item]) > threshold, user in userRecord); users left empty go;
      -sample   of the remaining test users, in the order in which they first appear among the test EVENTS (the key order of
                Record's testSet), the first int(0.9 * count) are dropped (45-49).
    Returns the new (test_indptr, test_items)."""
    test_indptr = np.asarray(test_indptr, dtype=np.int64)
    test_items = np.asarray(test_items, dtype=np.int32)
    m = len(test_indptr) - 1
    owner = np.repeat(np.arange(m, dtype=np.int64), np.diff(test_indptr))
    keep = np.ones(len(test_items), dtype=bool)
    train = log.is_test == 0
    if cold is not None and log.rec_type == 'track':               # trackRecord is keyed by TRACK names whatever the target
        plays = np.bincount(log.ev_item[train], minlength=log.n)
        has_train = np.bincount(log.ev_user[train], minlength=m) > 0
        keep &= ~(has_train[owner] & (plays[test_items] > int(cold)))
    if sample:
        left = np.bincount(owner[keep], minlength=m) > 0                 # users that still have a test row
        tu = log.ev_user[~train]
        first = np.full(m, len(tu), dtype=np.int64)
        first[tu[::-1]] = np.arange(len(tu) - 1, -1, -1, dtype=np.int64)
        users = np.flatnonzero(left)
        users = users[np.argsort(first[users], kind='stable')]          # Record's testSet key order
        dropped = np.zeros(m, dtype=bool)
        dropped[users[:int(len(users) * 0.9)]] = True
        keep &= ~dropped[owner]
    out_indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(owner[keep], minlength=m), out=out_indptr[1:])
    return out_indptr, np.ascontiguousarray(test_items[keep])


class _IdToName(object):
    """id2name[kind][id] over an array."""

    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, i):
        return self.arr[int(i)]

    def __len__(self):
        return len(self.arr)


class ArrayRecord(object):
    """The slice of data/record.py's interface the GPU recommenders use, over an ArrayLog."""

    def __init__(self, log):
        self.log = log
        self.trainingData = range(log.train_size)              # len() is all anybody asks of it (BPR.py:28)
        self.recordCount = log.train_size
        self.id2name = {k: _IdToName(v) for k, v in log.names.items()}
        self._name2id, self._test = {}, None
        self.test_indptr = self.test_items = None              # set by the recommender from the device (K0)

    class _Lazy(dict):
        def __init__(self, rec):
            dict.__init__(self)
            self.rec = rec

        def __missing__(self, kind):
            self[kind] = {name: i for i, name in enumerate(self.rec.log.names[kind])}
            return self[kind]

    @property
    def name2id(self):
        if not isinstance(self._name2id, ArrayRecord._Lazy):
            self._name2id = ArrayRecord._Lazy(self)
        return self._name2id

    def getSize(self, t):
        return len(self.log.names[t])

    def getId(self, obj, t):
        ids = self.name2id[t]
        if obj in ids:
            return ids[obj]
        print('No ' + t + ' ' + obj + ' exists!')
        exit(-1)

    def contains(self, obj, t):
        return obj in self.name2id[t]

    def printTrainingSize(self):
        for kind in ('user', 'artist', 'album', 'track'):
            if kind in self.log.names:
                print(kind + ' count:', len(self.log.names[kind]))
        print('Training set size:', self.recordCount)

    @property
    def testSet(self):
        """{user: {track: 1}} built from the device's test CSR the first time someone wants the dict form."""
        if self._test is None:
            un, tn = self.log.names['user'], self.log.names[self.log.rec_type]
            self._test = {}
            for u in np.flatnonzero(np.diff(self.test_indptr) > 0):
                self._test[un[u]] = {tn[t]: 1 for t in self.test_items[self.test_indptr[u]:self.test_indptr[u + 1]]}
        return self._test


def hit_mask(users, ids, n, test_indptr, test_items):
    """[B, N] bool: is ids[b, r] a held-out track of users[b]?  One sorted search over (user, track) keys."""
    users = np.asarray(users, dtype=np.int64)
    owner = np.repeat(np.arange(len(test_indptr) - 1, dtype=np.int64), np.diff(test_indptr))
    keys = owner * n + np.asarray(test_items, dtype=np.int64)              # sorted: the CSR is user-major, rows ascending
    q = users[:, None] * n + np.where(ids >= 0, ids, 0).astype(np.int64)
    pos = np.searchsorted(keys, q.ravel()).reshape(q.shape)
    pos = np.minimum(pos, max(len(keys) - 1, 0))
    return (keys[pos] == q) & (ids >= 0) if len(keys) else np.zeros(ids.shape, dtype=bool)


def name_blob(names):
    """(bytes, int64 offsets[len + 1]) of an array of names: the form yue_result_lines reads."""
    enc = [x.encode('utf-8') for x in names]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum(np.fromiter((len(x) for x in enc), dtype=np.int64, count=len(enc)), out=off[1:])
    return b''.join(enc), off


def result_text(user_blob, user_off, track_blob, track_off, ids, hits):
    """IterativeRecommender.py:145-155 for all ranked users at once, as ONE string of lines, built by the library's host
    code on all cores (yue_result_lines); user_blob / user_off describe the ranked users in order."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    hits = np.ascontiguousarray(hits, dtype=np.uint8)
    B, N = ids.shape
    need = C.c_int64(0)
    args = (user_blob, user_off.ctypes.data_as(C.POINTER(C.c_int64)), track_blob, track_off.ctypes.data_as(C.POINTER(C.c_int64)),
            C.c_int64(len(track_off) - 1), ids.ctypes.data_as(C.POINTER(C.c_int32)), hits.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(B), C.c_int(N))
    lib.yue_result_lines(*args, None, C.c_int64(0), C.byref(need))
    buf = C.create_string_buffer(max(need.value, 1))
    rc = lib.yue_result_lines(*args, C.cast(buf, C.c_void_p), C.c_int64(need.value), C.byref(need))
    if rc:
        raise _lib.YueError(rc, lib.yue_last_error(None).decode())
    return buf.raw[:need.value].decode('utf-8')


def result_lines(user_names, track_names, ids, hits):
    """IterativeRecommender.py:145-155: 'user:' + the names of the ranked tracks, '*' after a hit, newline -- column-wise."""
    ids = np.asarray(ids)
    cells = np.where(ids >= 0, np.asarray(track_names, dtype=object)[np.where(ids >= 0, ids, 0)], '')
    cells = cells + np.where(hits, '*', '').astype(object)
    lines = np.asarray(user_names, dtype=object) + ':'
    for r in range(ids.shape[1]):
        lines = lines + cells[:, r]
    return list(lines + '\n')


def ranking_measure(ids, hits, n_test, tops, item_count):
    """Measure.rankingMeasure's list of strings (evaluation/measure.py:16-41: hits 7-13, precision 51-53, recall 91-94, F1
    97-101, MAP 56-66, coverage 43-48) and {n: NDCG@n} from the ranked ids [B, N], their hit flags and the users' numbers
    of held-out tracks -- whole-array arithmetic (the lists may come from several devices).  Same sums as the reference's
    loops up to the order of float64 additions."""
    ids, hits = np.asarray(ids), np.asarray(hits, dtype=bool)
    B = ids.shape[0]
    n_test = np.asarray(n_test, dtype=np.float64)
    cum = np.cumsum(hits, axis=1)
    ranks = np.arange(1, ids.shape[1] + 1, dtype=np.float64)
    disc = 1.0 / np.log2(ranks + 1.0)
    print('rank measure...')
    measure, ndcg = [], {}
    for n in tops:
        h = cum[:, n - 1] if n <= ids.shape[1] else cum[:, -1]
        prec = float(h.sum()) / (B * n)
        recall = float((h / n_test).sum()) / float(B)
        ap = ((cum[:, :n] / ranks[:n]) * hits[:, :n]).sum(axis=1) / np.minimum(n_test, n)
        f1 = 2 * prec * recall / (prec + recall) if (prec + recall) != 0 else 0
        shown = ids[:, :n]
        distinct = np.unique(shown[shown >= 0]).size
        measure += ['Top ' + str(n) + '\n', 'Precision:' + str(prec) + '\n', 'Recall:' + str(recall) + '\n', 'F1:' + str(f1) + '\n',
                    'MAP:' + str(float(ap.sum()) / B) + '\n', 'Coverage:' + str(distinct / float(item_count)) + '\n']
        ideal = np.cumsum(disc)[np.minimum(n_test, n).astype(np.int64) - 1]
        ndcg[n] = float(((hits[:, :n] * disc[:n]).sum(axis=1) / ideal).sum()) / B
    return measure, ndcg
