"""Drop-in replacement for the reference's recommender/advanced/LightGCN.py: the TF-1 graph and its training loop
(LightGCN.py:27-98) run as one cooperative CUDA launch per pass (yue_gcn_epoch), ranking on the propagated tables.
Derives from the reference's own base.IterativeRecommender (the reference's base.DeepRecommender has no .py extension
and cannot be imported; its batch_size key and its truncated-normal init are in the mixin); see INTEGRATION.md."""
from base.IterativeRecommender import IterativeRecommender

from yue_b200.lightgcn import GpuLightGCNMixin


class LightGCN(GpuLightGCNMixin, IterativeRecommender):

    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(LightGCN, self).__init__(conf, trainingSet, testSet, fold)
