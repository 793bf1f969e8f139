"""Drop-in replacement for the reference's recommender/cf/WRMF.py: same class name and constructor, derives from the
reference's own base.IterativeRecommender; the two ALS loops of WRMF.py:34-80 run on the GPU (yue_wrmf_sweep).  See
dropin/recommender/cf/BPR.py and INTEGRATION.md."""
from base.IterativeRecommender import IterativeRecommender

from yue_b200.wrmf import GpuWRMFMixin


class WRMF(GpuWRMFMixin, IterativeRecommender):
    # WRMF: Collaborative Filtering for Implicit Feedback Datasets (Hu, Koren, Volinsky)

    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(WRMF, self).__init__(conf, trainingSet, testSet, fold)
