"""WRMF -- implicit-feedback ALS (Hu, Koren, Volinsky) on the GPU behind the reference's class API
(recommender/cf/WRMF.py:13-88).  SURVEY.md 8f row 4: the first model after BPR/APR that reuses the tables, the
ingest, the ranking kernels and the metrics of the hot path; what is new is the half-sweep kernel (K7,
csrc/wrmf_als.cuh -> yue_wrmf_sweep).

Kept from the reference: X = P*10 and Y = Q*10 at init (19-20); confidence 1 + 10 r_ui with r_ui the play count
(28-33, 46-47); `reg.lambda -u` regularises BOTH sweeps (55, 79); exactly `num.max.iter` iterations, no convergence
test (84-85); the loss printed per iteration is the squared error over the played pairs measured during the user
sweep (49-50, 83); `predict = Y.dot(X[u])` (86-88).  Config keys as in config/WRMF.conf.
"""
from .bpr import GpuBPRMixin
from .host.recommender import IterativeRecommender


class GpuWRMFMixin(GpuBPRMixin):
    _user_table, _item_table = 'X', 'Y'
    alpha = 10.0

    def initModel(self):
        super(GpuWRMFMixin, self).initModel()
        self.X = self.P * 10
        self.Y = self.Q * 10

    def buildModel(self):
        print('training...')
        eng = self._push_factors()
        iteration = 0
        while iteration < self.maxIter:
            self.loss = eng.wrmf_sweep(0, self.regU, self.alpha, want_loss=True)
            eng.wrmf_sweep(1, self.regU, self.alpha)
            iteration += 1
            print('iteration:', iteration, 'loss:', self.loss)
        self._pull_factors()


class WRMF(GpuWRMFMixin, IterativeRecommender):
    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(WRMF, self).__init__(conf, trainingSet, testSet, fold)
