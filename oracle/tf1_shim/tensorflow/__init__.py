"""oracle/tf1_shim/tensorflow -- TEST INFRASTRUCTURE ONLY (never imported by the product, never shipped to users).

An eager stand-in for the handful of TensorFlow 1.x entry points that the reference's hot path calls, so that the
reference's OWN, UNMODIFIED source files /root/reference/dgcnn/ops.py and /root/reference/dgcnn/model.py can be imported
and executed in this container (TensorFlow 1.x itself cannot be installed: no network, no Python 3.12 build exists).
tests/golden/make_reference_golden.py puts this directory on sys.path, loads the two reference files by path and runs
them; what comes out is produced by the reference's code -- its index arithmetic, concat orders, scopes, residual and head
wiring -- with only the TF primitives below restated (each from its documented TF 1.x behaviour, stated next to it).
Tensors are torch CPU tensors (fp32 unless the caller feeds fp64), so torch autograd differentiates the reference's graph.

Entry points used by the reference (file:line of the call):
  tf.transpose / matmul / reduce_sum / square / nn.top_k          ops.py:12-18
  tf.shape / range / reshape / gather / expand_dims / tile / concat   ops.py:26-39
  tf.reduce_max / reduce_mean / squeeze / variable_scope / nn.relu    ops.py:56-58,91-96,134
  slim.conv2d / slim.batch_norm                                    ops.py:47-70,124-133,151-160; model.py:46-53,65-72,94-101
  gen_nn_ops.max_pool_v2                                           model.py:76-77
  tf.nn.dropout                                                    model.py:91
  tf.placeholder / device / name_scope / argmax / to_int64 / equal / cast / multiply / add_n / zeros_like / Variable /
  trainable_variables / train.AdamOptimizer / summary.* / nn.softmax / nn.sparse_softmax_cross_entropy_with_logits and
  Session.run                                                      trainval.py:14-129

Two modes.  Called on torch tensors every function computes at once (ops.py / model.py on concrete inputs).  Called on a
`tf.placeholder` (trainval.py builds its graph once and feeds it later) the same functions return `Node`s: the call is
recorded AND executed on an example value (dummy data of the placeholder's shape, unknown dimensions = 64 points), which
is what gives graph construction its static shapes and creates the variables; `Session.run(fetches, feed_dict)` re-executes
the recorded calls on the fed values, each node once per run.
"""
import contextlib
from collections import OrderedDict

import numpy as np
import torch

float32, int32, int64 = torch.float32, torch.int32, torch.int64

# --------------------------------------------------------------------------------------------------------- graph mode
import functools
import operator


class Node(object):
    """A recorded call: fn(*args, **kwargs) with Nodes (also inside lists / tuples) standing for values known at run time."""

    def __init__(self, fn, args=(), kwargs=None, example=None):
        self.fn, self.args, self.kwargs, self.example = fn, args, kwargs or {}, example

    def run(self, ctx):
        key = id(self)
        if key not in ctx["memo"]:
            ctx["memo"][key] = self.fn(*_resolve(self.args, ctx), **_resolve(self.kwargs, ctx))
        return ctx["memo"][key]

    @property
    def shape(self):
        return self.example.shape

    def __getitem__(self, i):
        return Node(operator.getitem, (self, i), example=self.example[i] if self.example is not None else None)

    def __iter__(self):                  # `_, idx = tf.nn.top_k(...)`
        return iter([self[i] for i in builtins_range(len(self.example))])

    def _bin(self, other, op, swap=False):
        a, b = (other, self) if swap else (self, other)
        return Node(op, (a, b), example=op(_example(a), _example(b)))

    def __add__(self, o): return self._bin(o, operator.add)
    def __radd__(self, o): return self._bin(o, operator.add, True)
    def __sub__(self, o): return self._bin(o, operator.sub)
    def __rsub__(self, o): return self._bin(o, operator.sub, True)
    def __mul__(self, o): return self._bin(o, operator.mul)
    def __rmul__(self, o): return self._bin(o, operator.mul, True)
    def __truediv__(self, o): return self._bin(o, operator.truediv)
    def __rtruediv__(self, o): return self._bin(o, operator.truediv, True)
    def __neg__(self): return Node(operator.neg, (self,), example=-self.example)


import builtins
builtins_range = builtins.range


def _walk(x, f):
    if isinstance(x, Node):
        return f(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_walk(v, f) for v in x)
    if isinstance(x, dict):
        return {k: _walk(v, f) for k, v in x.items()}
    return x


def _resolve(x, ctx):
    return _walk(x, lambda n: n.run(ctx))


def _example(x):
    return _walk(x, lambda n: n.example)


def _has_node(x):
    if isinstance(x, Node):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_node(v) for v in x)
    if isinstance(x, dict):
        return any(_has_node(v) for v in x.values())
    return False


def dual(fn):
    """eager on tensors; recorded (and run on the example values) when any argument is a Node"""
    @functools.wraps(fn)
    def wrapper(*a, **k):
        if _has_node(a) or _has_node(k):
            return Node(fn, a, k, example=fn(*_example(a), **_example(k)))
        return fn(*a, **k)
    return wrapper


# ------------------------------------------------------------------------------------------------ variables and traces
_scopes = []
VARIABLES = OrderedDict()      # TF variable name (without ":0") -> torch tensor, requires_grad for trainable ones
PRESET = {}                    # name -> initial value (2-D [Cin, Cout] or TF [1, 1, Cin, Cout] for weights)
TRACE = {"top_k": [], "top_k_input": [], "conv2d": OrderedDict()}
DROPOUT_MASK = None            # {0,1} keep mask used by nn.dropout instead of a random draw (parity runs)
_gen = torch.Generator().manual_seed(0)


def reset(seed=0):
    VARIABLES.clear()
    PRESET.clear()
    TRACE["top_k"], TRACE["top_k_input"], TRACE["conv2d"] = [], [], OrderedDict()
    _gen.manual_seed(seed)
    del _scopes[:]
    globals()["DROPOUT_MASK"] = None


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    """tf.variable_scope: nested names joined by '/'."""
    _scopes.append(name)
    try:
        yield name
    finally:
        _scopes.pop()


AUTO_REUSE = "AUTO_REUSE"


def _scope_name(leaf):
    return "/".join(_scopes + [leaf])


def _get_variable(name, shape, init, dtype):
    """create-or-reuse (the reference builds under reuse=tf.AUTO_REUSE, trainval.py:29)"""
    v = VARIABLES.get(name)
    if v is None:
        if name in PRESET:
            v = torch.as_tensor(np.asarray(PRESET[name])).to(dtype).reshape(shape).clone()
        elif init == "xavier":     # tf.contrib.layers.xavier_initializer(uniform=True): +-sqrt(6 / (fan_in + fan_out))
            fan_in, fan_out = shape[-2], shape[-1]
            lim = (6.0 / (fan_in + fan_out)) ** 0.5
            v = ((torch.rand(shape, generator=_gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
        else:
            v = torch.zeros(shape, dtype=dtype)
        v.requires_grad_(True)
        VARIABLES[name] = v
    return v


# ------------------------------------------------------------------------------------------------------- tensor ops
@dual
def transpose(a, perm):
    return a.permute(*perm)


@dual
def matmul(a, b):
    return torch.matmul(a, b)


@dual
def square(a):
    return a * a


@dual
def reduce_sum(a, axis=None, keepdims=False):
    return a.sum(dim=axis, keepdim=keepdims)


@dual
def reduce_max(a, axis=None, keepdims=False):
    return a.amax(dim=axis, keepdim=keepdims)


@dual
def reduce_mean(a, axis=None, keepdims=False):
    return a.mean() if axis is None else a.mean(dim=axis, keepdim=keepdims)


@dual
def shape(a):
    return [int(s) for s in a.shape]


@dual
def range(n):   # noqa: A001  (tf.range)
    return torch.arange(int(n), dtype=torch.int64)


@dual
def reshape(a, shp):
    return a.reshape([int(s) for s in shp])


@dual
def gather(params, indices):
    """tf.gather along axis 0 with an index tensor of any rank: result shape = indices.shape + params.shape[1:]."""
    return params[indices.long()]


@dual
def expand_dims(a, axis):
    return a.unsqueeze(axis)


@dual
def tile(a, multiples):
    return a.repeat(*[int(m) for m in multiples])


@dual
def concat(values, axis):
    return torch.cat(list(values), dim=axis)


@dual
def squeeze(a, axis=None):
    return a.squeeze(axis)


@dual
def argmax(a, axis=None):
    return a.argmax(dim=axis)


@dual
def to_int64(a):
    return a.long()


@dual
def equal(a, b):
    return a == b


@dual
def cast(a, dtype):
    return a.to(dtype)


@dual
def multiply(a, b):
    return a * b


@dual
def add_n(values):
    out = values[0]
    for v in values[1:]:
        out = out + v
    return out


@dual
def zeros_like(a):
    return torch.zeros_like(a)


@dual
def _relu(a):
    return torch.relu(a)


@dual
def _top_k(a, k=1, sorted=True):   # noqa: A002
    neg = (-a.detach()).numpy()
    idx = np.argsort(neg, axis=-1, kind="stable")[..., :int(k)]
    idx_t = torch.from_numpy(np.ascontiguousarray(idx)).long()
    TRACE["top_k_input"].append(a.detach())
    TRACE["top_k"].append(idx_t.to(torch.int32))
    return torch.gather(a, -1, idx_t), idx_t


_dropout_calls = [0]


@dual
def _dropout(x, keep_prob, noise_shape=None):
    mask = DROPOUT_MASK
    if isinstance(mask, (list, tuple)):          # one mask per dropout call of a run (one per tower), in call order
        mask = mask[_dropout_calls[0] % len(mask)]
        _dropout_calls[0] += 1
    if mask is None:
        mask = (torch.rand(x.shape, generator=_gen) < keep_prob).to(x.dtype)
    return x * mask.to(x.dtype) / keep_prob


@dual
def _softmax(logits):
    return torch.softmax(logits, dim=-1)


@dual
def _sparse_xent(labels=None, logits=None):
    """tf.nn.sparse_softmax_cross_entropy_with_logits: -log softmax(logits)[label] per position, shape = labels.shape."""
    flat = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1).long(), reduction="none")
    return flat.reshape(labels.shape)


class _NN(object):
    relu = staticmethod(_relu)

    @staticmethod
    def top_k(a, k=1, sorted=True):   # noqa: A002
        """tf.nn.top_k: the k largest entries of the last axis in descending order; "if two elements are equal, the
        lower-index element appears first" (TF 1.x API documentation).  A stable sort of the negated values is that rule
        (torch.topk gives no tie guarantee)."""
        return _top_k(a, k=k, sorted=sorted)

    @staticmethod
    def dropout(x, keep_prob, noise_shape=None):
        """tf.nn.dropout(x, keep_prob): x / keep_prob where a uniform draw < keep_prob, else 0."""
        return _dropout(x, keep_prob, noise_shape)

    @staticmethod
    def softmax(logits):
        return _softmax(logits)

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(labels=None, logits=None):
        return _sparse_xent(labels=labels, logits=logits)


nn = _NN()


# ------------------------------------------------------------------------------ what trainval.py needs on top (graph mode)
UNKNOWN_DIM = 64     # example size of a `None` placeholder dimension (points per cloud); >= any k the tests use


def placeholder(dtype, shape=None):
    """tf.placeholder: fed at Session.run.  The example value (random data / zero labels of the declared shape) exists only
    to give the graph under construction concrete shapes."""
    shp = [UNKNOWN_DIM if d is None else int(d) for d in shape]
    ex = torch.randn(shp, generator=_gen).to(dtype) if dtype.is_floating_point else torch.zeros(shp, dtype=dtype)
    node = Node(None, example=ex)
    node.fn = lambda: (_ for _ in ()).throw(RuntimeError("placeholder was not fed"))
    node.is_placeholder, node.dtype = True, dtype
    return node


@contextlib.contextmanager
def device(name):
    yield


@contextlib.contextmanager
def name_scope(name):
    yield name


class Variable(Node):
    """tf.Variable / the objects tf.trainable_variables() returns: a Node that reads the live tensor."""

    def __init__(self, initial_value=None, trainable=True, name=None, _tensor=None):
        if _tensor is None:
            init = _example(initial_value)
            _tensor = init.detach().clone()
        self.tensor, self.name, self.trainable = _tensor, name, trainable
        Node.__init__(self, None, example=_tensor)

    def run(self, ctx):
        return self.tensor

    def initialized_value(self):
        return self

    def assign(self, value):
        def fn(v):
            with torch.no_grad():
                self.tensor.copy_(v)
            return self.tensor.detach().clone()
        return Node(fn, (value,), example=self.tensor)

    def assign_add(self, value):
        def fn(v):
            with torch.no_grad():
                self.tensor.add_(v)
            return self.tensor.detach().clone()
        return Node(fn, (value,), example=self.tensor)


def trainable_variables():
    """creation order, like the TRAINABLE_VARIABLES collection"""
    return [Variable(_tensor=t, name=n + ":0") for n, t in VARIABLES.items()]


class _AdamOptimizer(object):
    """tf.train.AdamOptimizer (TF 1.x docstring): lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);
    m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;  variable -= lr_t * m / (sqrt(v) + epsilon)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t, self.slots = float(learning_rate), beta1, beta2, epsilon, 0, {}

    def compute_gradients(self, loss):
        tvars = trainable_variables()

        def grads(l):
            got = torch.autograd.grad(l, [v.tensor for v in tvars], retain_graph=True, allow_unused=True)
            return tuple(torch.zeros_like(v.tensor) if g is None else g for g, v in zip(got, tvars))
        allg = Node(grads, (loss,), example=tuple(torch.zeros_like(v.tensor) for v in tvars))
        return [(allg[i], v) for i, v in enumerate(tvars)]

    def apply_gradients(self, grads_and_vars):
        pairs = list(grads_and_vars)

        def fn(*gvals):
            self.t += 1
            lr_t = self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)
            with torch.no_grad():
                for g, (_, var) in zip(gvals, pairs):
                    m, v = self.slots.setdefault(var.name, (torch.zeros_like(var.tensor), torch.zeros_like(var.tensor)))
                    m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                    v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                    var.tensor.sub_(lr_t * m / (v.sqrt() + self.eps))
            return None
        return Node(fn, tuple(g for g, _ in pairs))


class _Train(object):
    AdamOptimizer = _AdamOptimizer


train = _Train()


class _Summary(object):
    @staticmethod
    def scalar(name, tensor):
        return None

    @staticmethod
    def merge_all():
        return Node(lambda: b"", example=b"")


summary = _Summary()


class Session(object):
    """sess.run(fetches, feed_dict): every recorded call is executed once on the fed values; nested lists keep their
    structure; tensors come back as numpy arrays like in TF."""

    def run(self, fetches, feed_dict=None):
        ctx = {"memo": {}}
        for ph, val in (feed_dict or {}).items():
            t = torch.as_tensor(np.asarray(val)).to(ph.dtype)
            ctx["memo"][id(ph)] = t
        TRACE["top_k"], TRACE["top_k_input"] = [], []
        _dropout_calls[0] = 0
        out = _resolve(fetches, ctx)

        def to_np(v):
            if isinstance(v, torch.Tensor):
                return v.detach().numpy()
            if isinstance(v, (list, tuple)):
                return type(v)(to_np(x) for x in v)
            return v
        return to_np(out)
