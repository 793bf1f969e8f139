"""Test infrastructure: a minimal stand-in for the TensorFlow 1.x API that recommender/advanced/LightGCN.py,
recommender/advanced/APR.py and base/DeepRecommender use, built on torch (CPU, float64, autograd) -- so that the REFERENCE'S OWN graph-building and training
text can be executed unmodified in the build container (oracle/make_golden_lightgcn.py), where TensorFlow 1 does not exist.

What this pins and what it does not.  Executing the reference's text pins everything the reference WROTE: which tensors feed
which layer, the per-event adjacency entries, sum instead of mean, the loss and its regulariser, one optimiser over U and V,
the batch slices in file order, the five-draws-keep-the-last sampler.  The MEANING of each TensorFlow op is restated here
from TensorFlow's documentation and is not pinned by anything in the reference tree:
  * `sparse_tensor_dense_matmul`: every (row, col, value) entry adds value * B[col] to row `row`; duplicate entries add up;
  * `nn.l2_normalize(x, axis=1)`: x * rsqrt(max(sum(x^2, axis), 1e-12));   `nn.l2_loss(x)`: sum(x^2) / 2;
  * `train.AdamOptimizer(lr)`: beta1 0.9, beta2 0.999, epsilon 1e-8; lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);
    m, v moving averages of g and g^2; var -= lr_t m / (sqrt(v) + epsilon); every variable the loss depends on, every step;
  * `truncated_normal(stddev)`: normal draws beyond two standard deviations are redrawn (numpy's global generator here);
  * `reduce_sum` of a Python list of tensors stacks them first;
  * `gradients(y, [variables])`: dense d y / d variable (zero rows where the batch does not touch the variable);
    `nn.softplus(x)` = log(1 + e^x); `Variable.assign` takes effect when the returned node is run.
A graph is a tree of lazy nodes; `Session.run(fetches, feed_dict)` evaluates them with torch (one cache per run), the
`minimize` node runs autograd on the loss and applies Adam.  Only what those two files touch exists.
"""
import numpy as np
import torch

int32 = "int32"
float32 = "float32"
DT = torch.float64
_VARIABLES = []


class Node(object):
    def __init__(self, fn, *inputs):
        self.fn, self.inputs = fn, inputs

    def value(self, feeds, cache):
        if id(self) not in cache:
            cache[id(self)] = self.fn(*[_val(x, feeds, cache) for x in self.inputs])
        return cache[id(self)]

    def __add__(self, o): return Node(lambda a, b: a + b, self, o)
    def __radd__(self, o): return Node(lambda a, b: b + a, self, o)
    def __sub__(self, o): return Node(lambda a, b: a - b, self, o)
    def __rsub__(self, o): return Node(lambda a, b: b - a, self, o)
    def __mul__(self, o): return Node(lambda a, b: a * b, self, o)
    def __rmul__(self, o): return Node(lambda a, b: b * a, self, o)
    def __neg__(self): return Node(lambda a: -a, self)


def _val(x, feeds, cache):
    if isinstance(x, Node):
        return x.value(feeds, cache)
    if isinstance(x, (list, tuple)):
        return [_val(y, feeds, cache) for y in x]
    return x


class Placeholder(Node):
    def __init__(self, dtype, name):
        Node.__init__(self, None)
        self.dtype, self.name = dtype, name

    def value(self, feeds, cache):
        v = feeds[self]
        return torch.as_tensor(np.asarray(v), dtype=torch.int64 if self.dtype == int32 else DT)


class Variable(Node):
    def __init__(self, initial_value, name=None, dtype=None, trainable=True):
        Node.__init__(self, None)
        self.name, self.trainable = name, trainable
        self.tensor = torch.tensor(np.asarray(initial_value), dtype=DT, requires_grad=True)     # tf.gradients may ask for any variable
        if trainable:
            _VARIABLES.append(self)

    def value(self, feeds, cache):
        return self.tensor

    def assign(self, value):
        var = self

        def fn(v):
            with torch.no_grad():
                var.tensor.copy_(v)
            return var.tensor
        return Node(fn, value)


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, name)


def truncated_normal(shape, stddev=1.0):
    x = np.random.standard_normal(tuple(shape))
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = np.random.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (stddev * x).astype(np.float32)              # TF's variables are float32; the arithmetic below is float64


def set_random_seed(seed):
    pass                                                # graph-level seed: the initial values come from numpy here


def cast(x, dtype):
    return x


def concat(values, axis=0):
    return Node(lambda vs: torch.cat(vs, dim=axis), list(values))


class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        idx = torch.tensor(np.asarray(indices, dtype=np.int64).T)
        self.rows, self.cols = idx[0], idx[1]
        self.values = torch.tensor(np.asarray(values, dtype=np.float64), dtype=DT)
        self.shape = tuple(int(x) for x in dense_shape)


def sparse_tensor_dense_matmul(sp, b):
    def fn(dense):
        out = torch.zeros((sp.shape[0], dense.shape[1]), dtype=DT)
        return out.index_add(0, sp.rows, sp.values[:, None] * dense[sp.cols])       # one term per entry: duplicates add up
    return Node(fn, b)


def reduce_sum(x, axis=None):
    def fn(v):
        if isinstance(v, list):
            v = torch.stack(v)
        return v.sum() if axis is None else v.sum(dim=axis)
    return Node(fn, x)


def split(value, num_or_size_splits, axis=0):
    sizes = [int(s) for s in num_or_size_splits]
    whole = Node(lambda v: torch.split(v, sizes, dim=axis), value)
    return [Node(lambda parts, k=k: parts[k], whole) for k in range(len(sizes))]


def zeros(shape, dtype=None):
    return np.zeros(tuple(shape), dtype=np.float32)


def constant(value, dtype=None):
    return float(value)


def add(a, b):
    return Node(lambda x, y: x + y, a, b)


def subtract(a, b):
    return Node(lambda x, y: x - y, a, b)


def stop_gradient(x):
    return Node(lambda v: v.detach(), x)


def gradients(ys, xs):
    """d ys / d x for every x in xs, dense (TF returns IndexedSlices for gathered variables; APR.py densifies them)."""
    def one(x):
        def fn(y):
            g = torch.autograd.grad(y, x.tensor, retain_graph=True, allow_unused=True)[0]
            return torch.zeros_like(x.tensor) if g is None else g
        return Node(fn, ys)
    return [one(x) for x in xs]


def multiply(a, b):
    return Node(lambda x, y: x * y, a, b)


def log(x):
    return Node(torch.log, x)


def sigmoid(x):
    return Node(torch.sigmoid, x)


class _NN(object):
    @staticmethod
    def embedding_lookup(params, ids):
        return Node(lambda p, i: p[i], params, ids)

    @staticmethod
    def l2_normalize(x, axis=None, epsilon=1e-12):
        return Node(lambda v: v * torch.rsqrt(torch.clamp((v * v).sum(dim=axis, keepdim=True), min=epsilon)), x)

    @staticmethod
    def l2_loss(x):
        return Node(lambda v: (v * v).sum() / 2, x)

    @staticmethod
    def softplus(x):
        return Node(torch.nn.functional.softplus, x)


nn = _NN()


class _Adam(object):
    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = float(learning_rate), beta1, beta2, epsilon, 0
        self.m, self.v = {}, {}

    def minimize(self, loss):
        opt = self

        class Train(Node):
            def value(self, feeds, cache):
                lv = loss.value(feeds, cache)
                tensors = [v.tensor for v in _VARIABLES]
                grads = torch.autograd.grad(lv, tensors, allow_unused=True)
                opt.t += 1
                lr_t = opt.lr * np.sqrt(1 - opt.b2 ** opt.t) / (1 - opt.b1 ** opt.t)
                with torch.no_grad():
                    for var, g in zip(_VARIABLES, grads):
                        if g is None:
                            continue
                        m = opt.m.setdefault(id(var), torch.zeros_like(var.tensor))
                        v = opt.v.setdefault(id(var), torch.zeros_like(var.tensor))
                        m.mul_(opt.b1).add_((1 - opt.b1) * g)
                        v.mul_(opt.b2).add_((1 - opt.b2) * g * g)
                        var.tensor -= lr_t * m / (torch.sqrt(v) + opt.eps)
                return None
        return Train(None)


class _Train(object):
    AdamOptimizer = _Adam


train = _Train()


def global_variables_initializer():
    return Node(lambda: None)


class Session(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        feeds, cache = feed_dict or {}, {}
        many = isinstance(fetches, (list, tuple))
        out = []
        # the loss is evaluated BEFORE the update is applied, like TF's control flow for [train, loss] in one run call
        order = sorted(range(len(fetches)), key=lambda k: hasattr(fetches[k], 'inputs') and type(fetches[k]).__name__ == 'Train') if many else [0]
        res = {}
        for k in order:
            f = fetches[k] if many else fetches
            v = f.value(feeds, cache)
            res[k] = None if v is None else (v.detach().numpy().copy() if isinstance(v, torch.Tensor) else v)
        out = [res[k] for k in range(len(order))]
        return out if many else out[0]


def variables():
    return list(_VARIABLES)


def reset():
    del _VARIABLES[:]
