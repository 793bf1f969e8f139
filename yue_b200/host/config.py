"""Config / LineConfig with the behaviour of the reference's tool/config.py:3-88.

``Config(path)`` reads ``key=value`` lines (a line must split on exactly one '='; other lines are
reported and skipped; a missing file raises IOError, tool/config.py:28-40).  ``LineConfig(text)``
is the mini command line inside a value: whitespace-separated ``-flag value...`` groups with an
optional leading ``on``/``off`` main switch (tool/config.py:44-65).  A token is a flag when it
starts with '-' and the rest is not all digits -- so ``-5`` is a value while ``-0.5`` is (mis)read
as a flag, exactly like the reference.  Unknown keys print ``parameter <k> is invalid!`` and
exit(-1) (tool/config.py:8-13, 68-72).
"""
import os


def _invalid(key):
    print('parameter ' + key + ' is invalid!')
    exit(-1)


class Config(object):
    def __init__(self, fileName=None, values=None):
        self.config = {}
        if values is not None:
            self.config.update(values)
        if fileName is not None:
            self.readConfiguration(fileName)

    def __getitem__(self, item):
        if item not in self.config:
            _invalid(item)
        return self.config[item]

    getOptions = __getitem__

    def contains(self, key):
        return key in self.config

    def readConfiguration(self, fileName):
        path = os.path.abspath(fileName)
        if not os.path.exists(path):
            print('config file is not found!')
            raise IOError
        with open(path) as f:
            for lineno, raw in enumerate(f):
                text = raw.strip()
                if not text:
                    continue
                parts = text.split('=')
                if len(parts) != 2:
                    print('config file is not in the correct format! Error Line:%d' % lineno)
                    continue
                self.config[parts[0]] = parts[1]


def _is_flag(token):
    return token.startswith('-') and not token[1:].isdigit()


class LineConfig(object):
    def __init__(self, content):
        self.line = content.strip().split(' ')
        self.mainOption = self.line[0] == 'on'
        self.options = {}
        flags = [i for i, tok in enumerate(self.line) if _is_flag(tok)]
        for a, i in enumerate(flags):
            end = flags[a + 1] if a + 1 < len(flags) else len(self.line)
            self.options[self.line[i]] = ' '.join(self.line[i + 1:end])

    def __getitem__(self, item):
        if item not in self.options:
            _invalid(item)
        return self.options[item]

    getOption = __getitem__

    def isMainOn(self):
        return self.mainOption

    def contains(self, key):
        return key in self.options
