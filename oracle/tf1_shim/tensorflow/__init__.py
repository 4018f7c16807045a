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
"""
import contextlib
from collections import OrderedDict

import numpy as np
import torch

float32, int32, int64 = torch.float32, torch.int32, torch.int64

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
def transpose(a, perm):
    return a.permute(*perm)


def matmul(a, b):
    return torch.matmul(a, b)


def square(a):
    return a * a


def reduce_sum(a, axis=None, keepdims=False):
    return a.sum(dim=axis, keepdim=keepdims)


def reduce_max(a, axis=None, keepdims=False):
    return a.amax(dim=axis, keepdim=keepdims)


def reduce_mean(a, axis=None, keepdims=False):
    return a.mean(dim=axis, keepdim=keepdims)


def shape(a):
    return [int(s) for s in a.shape]


def range(n):   # noqa: A001  (tf.range)
    return torch.arange(int(n), dtype=torch.int64)


def reshape(a, shp):
    return a.reshape([int(s) for s in shp])


def gather(params, indices):
    """tf.gather along axis 0 with an index tensor of any rank: result shape = indices.shape + params.shape[1:]."""
    return params[indices.long()]


def expand_dims(a, axis):
    return a.unsqueeze(axis)


def tile(a, multiples):
    return a.repeat(*[int(m) for m in multiples])


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def squeeze(a, axis=None):
    return a.squeeze(axis)


class _NN(object):
    @staticmethod
    def relu(a):
        return torch.relu(a)

    @staticmethod
    def top_k(a, k=1, sorted=True):   # noqa: A002
        """tf.nn.top_k: the k largest entries of the last axis in descending order; "if two elements are equal, the
        lower-index element appears first" (TF 1.x API documentation).  A stable sort of the negated values is that rule
        (torch.topk gives no tie guarantee)."""
        neg = (-a.detach()).numpy()
        idx = np.argsort(neg, axis=-1, kind="stable")[..., :int(k)]
        idx_t = torch.from_numpy(np.ascontiguousarray(idx)).long()
        TRACE["top_k_input"].append(a.detach())
        TRACE["top_k"].append(idx_t.to(torch.int32))
        return torch.gather(a, -1, idx_t), idx_t

    @staticmethod
    def dropout(x, keep_prob, noise_shape=None):
        """tf.nn.dropout(x, keep_prob): x / keep_prob where a uniform draw < keep_prob, else 0."""
        mask = DROPOUT_MASK
        if mask is None:
            mask = (torch.rand(x.shape, generator=_gen) < keep_prob).to(x.dtype)
        return x * mask.to(x.dtype) / keep_prob

    @staticmethod
    def softmax(logits):
        return torch.softmax(logits, dim=-1)


nn = _NN()
