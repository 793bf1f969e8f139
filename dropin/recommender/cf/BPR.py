"""Drop-in replacement for the reference's recommender/cf/BPR.py.

Copy this file over recommender/cf/BPR.py in a checkout of 0411tony/Yue (or put `dropin/` ahead
of the checkout on sys.path) and make the `yue_b200` package importable.  `Yue.execute`
(yue.py:59-70) keeps doing `from recommender.cf.BPR import BPR`; the class below derives from the
reference's OWN base.IterativeRecommender, so configuration parsing, Record, Measure, the lr
schedule and the result files are the reference's code -- only initModel / buildModel / predict /
evalRanking / ranking_performance run on the GPU (yue_b200.bpr.GpuBPRMixin -> libyue_b200.so).
No tensorflow import: the reference's live TF/Adam path is replaced by the SGD loop its own source
keeps as a comment (BPR.py:31-62), see DESIGN.md.
"""
from base.IterativeRecommender import IterativeRecommender

from yue_b200.bpr import GpuBPRMixin


class BPR(GpuBPRMixin, IterativeRecommender):
    # BPR: Bayesian Personalized Ranking from Implicit Feedback
    # Steffen Rendle, Christoph Freudenthaler, Zeno Gantner and Lars Schmidt-Thieme

    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(BPR, self).__init__(conf, trainingSet, testSet, fold)
