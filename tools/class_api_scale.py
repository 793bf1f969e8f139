#!/usr/bin/env python
"""Wall time of Yue -> BPR.execute() through the class API on one log, with the log as lists of dicts (the reference's data
path: tool/file.py loader, tool/dataSplit.py, data/record.py, per-user result lines, Measure) and as arrays
(yue.ingest=arrays: yue_b200/ingest.py).  SURVEY 8f rows 1-2: what the steps either side of the hot path cost.
usage: python tools/class_api_scale.py [users tracks plays]"""
import io
import os
import random
import sys
import tempfile
import time
from contextlib import redirect_stdout

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yue_b200 import synth  # noqa: E402
from yue_b200.bpr import BPR  # noqa: E402
from yue_b200.host.config import Config  # noqa: E402
from yue_b200.host.driver import Yue  # noqa: E402

users, tracks, plays = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (100_000, 50_000, 2_000_000)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "log.txt")
t0 = time.time()
synth.write_csv_log(path, users, tracks, plays, seed=20260102)
print("log: %d users x %d tracks x %d plays written in %.1f s (%.0f MB)" % (users, tracks, plays, time.time() - t0, os.path.getsize(path) / 1e6), flush=True)
from yue_b200.engine import Engine  # noqa: E402
Engine(0).close()                                   # the CUDA context is created once, outside the timings
modes = ("dicts", "arrays") if plays <= 5_000_000 else ("arrays",)      # the dict path needs minutes and tens of GB beyond that
for name in modes:
    vals = {"record": path, "record.setup": "-columns user:1,track:2,artist:3,time:0 -delim ,", "recommender": "BPR",
            "evaluation.setup": "-target track -ap 0.2", "item.ranking": "-topN 10", "num.factors": "64", "num.max.iter": "4",
            "learnRate": "-init 0.02 -max 1", "reg.lambda": "-u 0.01 -i 0.01 -b 0.2 -s 0.2", "output.setup": "on -dir %s/%s/" % (tmp, name),
            "yue.seed": "77"}
    if name == "arrays":
        vals["yue.ingest"] = "arrays"
    random.seed(5)
    np.random.seed(11)
    t = [time.time()]
    with redirect_stdout(io.StringIO()):
        y = Yue(Config(values=vals)); t.append(time.time())
        model = BPR(y.config, y.trainingData, y.testData); t.append(time.time())
        model.readConfiguration(); model.initModel(); t.append(time.time())
        model.buildModel(); t.append(time.time())
        model.evalRanking(); t.append(time.time())
    d = np.diff(t)
    print("%-6s load+split %.2f s | Record %.2f s | init %.2f s | buildModel (upload + 4 epochs) %.2f s | evalRanking (%d users, lines, measures, files) %.2f s | total %.2f s | %s"
          % (name, d[0], d[1], d[2], d[3], len(model.rec_users) if hasattr(model, "rec_users") else len(model.data.testSet), d[4], t[-1] - t[0],
             "".join(model.measure[1:3]).replace("\n", " ")), flush=True)
