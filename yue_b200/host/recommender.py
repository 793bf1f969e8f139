"""Recommender / IterativeRecommender: the template-method driver and the hyper-parameter,
learning-rate and convergence logic of the reference's base/recommender.py:9-174 and
base/IterativeRecommender.py:11-75, for running the GPU recommenders without the reference tree.
The scoring/ranking methods live in yue_b200/bpr.py (GPU); nothing here computes scores."""
from collections import defaultdict
from math import isnan
from os.path import abspath

import numpy as np

from .config import LineConfig
from .record import Record


class Recommender(object):
    def __init__(self, conf, trainingSet=None, testSet=None, fold='[1]'):
        self.config = conf
        self.isSaveModel = False
        self.isLoadModel = False
        self.isOutput = True
        from ..ingest import ArrayLog, ArrayRecord
        arrays = isinstance(trainingSet, ArrayLog)             # yue.ingest=arrays: numbered events instead of lists of dicts
        self.data = ArrayRecord(trainingSet) if arrays else Record(self.config, trainingSet, testSet)
        self.foldInfo = fold
        self.evalConfig = LineConfig(self.config['evaluation.setup'])
        self.recType = self.evalConfig['-target'] if self.evalConfig.contains('-target') else 'track'
        self.measure = []
        if arrays:
            return                                             # -cold / -sample: ingest.filter_test_rows, when the test CSR exists
        if self.evalConfig.contains('-cold'):
            # keep only held-out tracks with at most `threshold` training plays (recommender.py:22-42)
            threshold = int(self.evalConfig['-cold'])
            dropped = defaultdict(list)
            for user in self.data.testSet:
                if user in self.data.userRecord:
                    for item in self.data.testSet[user]:
                        if len(self.data.trackRecord[item]) > threshold:
                            dropped[user].append(item)
            for user, items in dropped.items():
                for item in items:
                    del self.data.testSet[user][item]
                if len(self.data.testSet[user]) == 0:
                    del self.data.testSet[user]
        if self.evalConfig.contains('-sample'):
            # evaluate on the last tenth of the test users (recommender.py:45-49)
            users = list(self.data.testSet.keys())
            for user in users[:int(len(users) * 0.9)]:
                del self.data.testSet[user]

    def readConfiguration(self):
        self.algorName = self.config['recommender']
        self.output = LineConfig(self.config['output.setup'])
        self.isOutput = self.output.isMainOn()
        self.ranking = LineConfig(self.config['item.ranking'])

    def printAlgorConfig(self):
        print('Algorithm:', self.config['recommender'])
        print('Training set:', abspath(self.config['record']))
        if self.evalConfig.contains('-testSet'):
            print('Test set:', abspath(self.evalConfig.getOption('-testSet')))
        self.data.printTrainingSize()
        print('=' * 80)

    def initModel(self):
        pass

    def buildModel(self):
        pass

    def saveModel(self):
        pass

    def loadModel(self):
        pass

    def predict(self, user):
        return []

    def evalRanking(self):
        raise NotImplementedError('list-style ranking of non-factor models is outside the hot path')

    def execute(self):
        """recommender.py:152-174: configuration -> model -> ranking evaluation."""
        self.readConfiguration()
        if self.foldInfo == '[1]':
            self.printAlgorConfig()
        if self.isLoadModel:
            print('Loading model %s...' % (self.foldInfo))
            self.loadModel()
        else:
            print('Initializing model %s...' % (self.foldInfo))
            self.initModel()
            print('Building Model %s...' % (self.foldInfo))
            self.buildModel()
        print('Predicting %s...' % (self.foldInfo))
        self.evalRanking()
        if self.isSaveModel:
            print('Saving model %s...' % (self.foldInfo))
            self.saveModel()
        return self.measure


class IterativeRecommender(Recommender):
    def readConfiguration(self):
        super(IterativeRecommender, self).readConfiguration()
        self.k = int(self.config['num.factors'])
        self.maxIter = int(self.config['num.max.iter'])
        rate = LineConfig(self.config['learnRate'])
        self.lRate = float(rate['-init'])
        self.maxLRate = float(rate['-max'])
        reg = LineConfig(self.config['reg.lambda'])
        self.regU, self.regI, self.regB = float(reg['-u']), float(reg['-i']), float(reg['-b'])

    def printAlgorConfig(self):
        super(IterativeRecommender, self).printAlgorConfig()
        print('Reduced Dimension:', self.k)
        print('Maximum Iteration:', self.maxIter)
        print('Regularization parameter: regU %.3f, regI %.3f, regB %.3f' % (self.regU, self.regI, self.regB))
        print('=' * 80)

    def initModel(self):
        # U[0,1) in float64 -> float32 -> /10, from the global numpy stream like the reference
        self.P = np.random.rand(self.data.getSize('user'), self.k).astype(np.float32) / 10
        self.Q = np.random.rand(self.data.getSize(self.recType), self.k).astype(np.float32) / 10
        self.loss, self.lastLoss = 0, 0

    def updateLearningRate(self, iter):
        if iter > 1:
            self.lRate *= 1.01 if abs(self.lastLoss) > abs(self.loss) else 0.5
        if self.maxLRate > 0 and self.lRate > self.maxLRate:
            self.lRate = self.maxLRate

    def isConverged(self, iter):
        if isnan(self.loss):
            print('Loss = NaN or Infinity: current settings does not fit the recommender! Change the settings and try again!')
            exit(-1)
        deltaLoss = self.lastLoss - self.loss
        print('%s %s iteration %d: loss = %.4f, delta_loss = %.5f learning_Rate = %.5f'
              % (self.algorName, self.foldInfo, iter, self.loss, deltaLoss, self.lRate))
        converged = abs(deltaLoss) < 1e-3
        if not converged:
            self.updateLearningRate(iter)
        self.lastLoss = self.loss
        return converged
