"""Log loader / result writer with the behaviour of the reference's tool/file.py:10-52 and the
splitters of tool/dataSplit.py:9-37.  Events are dicts keyed in ``-columns`` order."""
import os
import re
from random import random


class FileIO(object):
    @staticmethod
    def writeFile(dir, file, content, op='w'):
        if not os.path.exists(dir):
            os.makedirs(dir)
        with open(dir + file, op) as f:
            f.writelines(content)

    @staticmethod
    def deleteFile(filePath):
        if os.path.exists(filePath):
            os.remove(filePath)

    @staticmethod
    def loadDataSet(file, columns, binarized=False, threshold=3, delim=''):
        print('load dataset...')
        names = list(columns.keys())
        if len(names) < 2:
            print('The dataset needs more information or the record.setup setting has some problems...')
            exit(-1)
        where = [int(v) for v in columns.values()]
        splitter = re.compile(delim if delim != '' else ',| |\t')
        record = []
        with open(file) as f:
            for lineNo, line in enumerate(f, 1):
                fields = splitter.split(line.strip())
                try:
                    event = {name: fields[ind] for name, ind in zip(names, where)}
                except IndexError:
                    print('The record file is not in a correct format. Error Location: Line num %d' % lineNo)
                    exit(-1)
                if binarized and 'play' in event:
                    # the reference re-applies the threshold once per column from 'play' onwards
                    # (tool/file.py:42-47 sits inside the column loop); kept for identical output
                    for _ in range(len(names) - names.index('play')):
                        event['play'] = 1 if int(event['play']) >= threshold else 0
                record.append(event)
        return record


class DataSplit(object):
    @staticmethod
    def dataSplit(data, test_ratio=0.3, output=False, path='./', order=1):
        if test_ratio >= 1 or test_ratio <= 0:
            test_ratio = 0.3
        trainingSet, testSet = [], []
        for entry in data:
            (testSet if random() < test_ratio else trainingSet).append(entry)
        if output:
            FileIO.writeFile(path, 'testSet[' + str(order) + ']', testSet)
            FileIO.writeFile(path, 'trainingSet[' + str(order) + ']', trainingSet)
        return trainingSet, testSet

    @staticmethod
    def crossValidation(data, k):
        if k <= 1 or k > 10:
            k = 3
        for fold in range(k):
            yield ([e for ind, e in enumerate(data) if ind % k != fold],
                   [e for ind, e in enumerate(data) if ind % k == fold])
