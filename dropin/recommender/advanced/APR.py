"""Drop-in replacement for the reference's recommender/advanced/APR.py (which cannot be imported as
shipped: base/DeepRecommender has no .py suffix).  Derives from the reference's own
base.IterativeRecommender; see dropin/recommender/cf/BPR.py and INTEGRATION.md."""
from base.IterativeRecommender import IterativeRecommender

from yue_b200.apr import GpuAPRMixin


class APR(GpuAPRMixin, IterativeRecommender):
    # APR: Adversarial Personalized Ranking for Recommendation

    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(APR, self).__init__(conf, trainingSet, testSet, fold)
