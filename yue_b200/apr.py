"""APR -- Adversarial Personalized Ranking (He et al. 2018) on the GPU behind the reference's class
API (recommender/advanced/APR.py:13-143).

Config keys as in config/APR.conf: `APR=-regA 2 -eps 0.5 -advEpoch 100`, `batch_size` (read and kept
for compatibility: the kernels work per triplet, not per mini-batch), `num.max.iter`, `learnRate`,
`reg.lambda`.  Like the reference's buildModel (APR.py:113-137) training has two phases: plain BPR
for `num.max.iter` rounds, then `-advEpoch` adversarial rounds; a round here is a pass over the
training events with 1 (phase 1) or `negativeCount = 3` (phase 2, APR.py:23, 95-111) negatives per
positive, where the reference takes one mini-batch of `batch_size` events x 3 negatives per round.
The update is SGD with the perturbation fused per triplet (oracle/apr_ref.py) instead of Adam on
batch-aggregated perturbations -- the TF-1 graph cannot be imported and is not the
target; DESIGN.md section 4 (K2a).
"""
import random

from .bpr import GpuBPRMixin
from .engine import MODE_HOGWILD, MODE_SERIAL
from .host.config import LineConfig
from .host.recommender import IterativeRecommender


class GpuAPRMixin(GpuBPRMixin):
    def readConfiguration(self):
        super(GpuAPRMixin, self).readConfiguration()
        args = LineConfig(self.config['APR'])
        self.eps = float(args['-eps'])
        self.regAdv = float(args['-regA'])
        self.advEpoch = int(args['-advEpoch'])
        self.negativeCount = 3
        self.batch_size = int(self.config['batch_size']) if self.config.contains('batch_size') else 512

    def buildModel(self):
        print('training...')
        eng = self._push_factors()
        mode = MODE_SERIAL if self._opt('yue.sgd', 'hogwild') == 'serial' else MODE_HOGWILD
        seed = int(self._opt('yue.seed', random.getrandbits(63)))
        devs = self._sharded_devices(mode)
        if devs:                                                # yue.devices=0,1,...: the same two phases, users sharded
            loss = 0.0
            with self._sharded_session(eng, devs) as run_epoch:
                for epoch in range(self.maxIter):
                    loss = run_epoch(seed, epoch)[0]
                    print('iteration:', epoch, 'loss:', loss)
                for epoch in range(self.advEpoch):
                    loss = sum(run_epoch(seed, self.maxIter + epoch, apr=(self.eps, self.regAdv, slot))[0] for slot in range(self.negativeCount))
                    print('iteration:', epoch, 'loss:', loss)
            self.loss = loss
            return
        for epoch in range(self.maxIter):                       # phase 1: BPR (APR.py:120-127)
            loss = eng.bpr_epoch(self.lRate, self.regU, self.regI, seed, epoch, mode)
            print('iteration:', epoch, 'loss:', loss)
        for epoch in range(self.advEpoch):                      # phase 2: adversarial (APR.py:129-137)
            loss = 0.0
            for slot in range(self.negativeCount):
                loss += eng.apr_epoch(self.lRate, self.regU, self.regI, self.eps, self.regAdv, seed,
                                      self.maxIter + epoch, slot, mode)
            print('iteration:', epoch, 'loss:', loss)
        self.loss = loss if (self.maxIter + self.advEpoch) else 0
        self._pull_factors()


class APR(GpuAPRMixin, IterativeRecommender):
    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        super(APR, self).__init__(conf, trainingSet, testSet, fold)
