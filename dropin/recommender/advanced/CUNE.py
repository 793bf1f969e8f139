"""Drop-in replacement for the reference's recommender/advanced/CUNE.py: the training loop (CUNE.py:118-178) runs in
yue_cune_epoch on the GPU, the embedding stage before it stays host-side and needs gensim exactly as the reference does.
Derives from the reference's own base.IterativeRecommender; see dropin/recommender/cf/BPR.py and INTEGRATION.md."""
from base.IterativeRecommender import IterativeRecommender

from yue_b200.cune import GpuCUNEMixin


class CUNE(GpuCUNEMixin, IterativeRecommender):
    # CUNE-BPR: Collaborative User Network Embedding for Social Recommender Systems

    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(CUNE, self).__init__(conf, trainingSet, testSet, fold)
