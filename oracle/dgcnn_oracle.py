"""oracle/dgcnn_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch-CPU tensors, fp32 by default, fp64 on request) of the
reference's hot path and of the model around it, op for op:

    /root/reference/dgcnn/ops.py:8-163    k_nn, edges, edge_conv, repeat_edge_conv,
                                          repeat_residual_edge_conv, fc
    /root/reference/dgcnn/model.py:9-106  build
    /root/reference/dgcnn/trainval.py:38-52  softmax / accuracy / loss

PIN: the reference has no tests, golden vectors or fixtures, and its arithmetic lives in TensorFlow 1.x
(`tensorflow >= v1.3`, README.md:6), which cannot be installed here (no network, no py3.12 build).  What CAN run here is
the reference's own source: /root/reference/dgcnn/ops.py and model.py are executed UNMODIFIED on top of oracle/tf1_shim
(an eager stand-in for the ~25 TF entry points they call) by tests/golden/make_reference_golden.py, and the committed
tests/golden/ref_*.npz hold what they produce -- k_nn / edges on exact-arithmetic clouds, and three whole models
(dgcnn, residual-dgcnn with a shortcut conv, residual-dgcnn-nofc) with every EdgeConv tensor, logits, loss and every
parameter gradient, in fp32 and in fp64, and (trainval.py, through the shim's graph mode) a two-tower trainer over two
optimizer steps of two micro-steps: losses, accumulated gradients, variables after apply_gradient, inference().  tests/test_oracle_vs_reference.py holds this module to those vectors: indices
bit for bit, tensors / logits to 2e-5 / 5e-5, fp64 logits and gradients to 1e-6 relative (i.e. the same formulae), and
regenerates them from the sources wherever /root/reference exists.  So the reference's composition -- index arithmetic,
concat orders, scopes and variable names, residual and head wiring -- is pinned by its own code.  NOT pinned by anything
executable: the semantics of the TF primitives themselves (slim.conv2d / batch_norm defaults, tf.nn.top_k's tie rule,
tf.nn.dropout's scaling, max_pool_v2), which shim and oracle both restate from the TF 1.x API documentation, and the
accumulation order inside TF's kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.  The product package (dynamic-gcnn_b200/dgcnn) never does.

Two k_nn flavours:
  * k_nn(..., exact=True)  -> the C restatement in knn_oracle.c (fixed fp32 operation
    order, ties -> lower index).  This defines "bit-exact".
  * k_nn(..., exact=False) -> literally the TF graph: matmul, norms, top_k on a
    materialised [B,N,N] matrix (MKL accumulation order, torch.topk tie order).  This is
    the path that is *timed* as the reference-equivalent CPU baseline.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libknn_oracle.so")
_lib = None

BN_EPS = 1e-3  # slim.batch_norm default epsilon [TF-default]
DROPOUT_KEEP = 0.7  # model.py:91  tf.nn.dropout(net, 0.7): second positional arg = keep_prob
CONV1_WIDTH = 64  # ops.py:63 hard-coded


def build_c_oracle(force: bool = False) -> str:
    """Compile knn_oracle.c (gcc) into oracle/_build/.  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "knn_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def _c():
    global _lib
    if _lib is None:
        build_c_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        f32p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        lib.oracle_knn.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p]
        lib.oracle_pairwise_distance.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        lib.oracle_topk_rows.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, i32p]
        lib.oracle_sqnorm.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p]
        for fn in (lib.oracle_knn, lib.oracle_pairwise_distance, lib.oracle_topk_rows, lib.oracle_sqnorm):
            fn.restype = None
        _lib = lib
    return _lib


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# ----------------------------------------------------------------------------- k_nn
def pairwise_distance(points) -> torch.Tensor:
    """ops.py:11-16 with the fixed fp32 order of knn_oracle.c.  -> [B,N,N] fp32."""
    x = _f32(points)
    B, N, C = x.shape
    D = np.empty((B, N, N), dtype=np.float32)
    _c().oracle_pairwise_distance(_p(x, ctypes.c_float), B, N, C, _p(D, ctypes.c_float))
    return torch.from_numpy(D)


def topk_rows(D, k: int) -> torch.Tensor:
    """ops.py:18 on a materialised matrix: k smallest per row, ascending, ties -> lower index."""
    d = _f32(D)
    N = d.shape[-1]
    rows = d.size // N
    idx = np.empty(d.shape[:-1] + (k,), dtype=np.int32)
    _c().oracle_topk_rows(_p(d, ctypes.c_float), rows, N, k, _p(idx, ctypes.c_int32))
    return torch.from_numpy(idx)


def k_nn(points, k: int, exact: bool = True) -> torch.Tensor:
    """ops.py:8-19.  -> idx [B,N,k] int32, nearest first (self included)."""
    if exact:
        x = _f32(points)
        B, N, C = x.shape
        assert 1 <= k <= N
        idx = np.empty((B, N, k), dtype=np.int32)
        _c().oracle_knn(_p(x, ctypes.c_float), B, N, C, k, _p(idx, ctypes.c_int32))
        return torch.from_numpy(idx)
    M = points
    inner = torch.matmul(M, M.transpose(1, 2))  # ops.py:12-13
    sq = torch.sum(M * M, dim=-1, keepdim=True)  # ops.py:14
    nn_dist = sq + sq.transpose(1, 2) - 2 * inner  # ops.py:15-16
    _, idx = torch.topk(-nn_dist, k=k, dim=-1, largest=True, sorted=True)  # ops.py:18
    return idx.to(torch.int32)


def knn_pure_python(points, k: int) -> np.ndarray:
    """Slow, loop-level restatement used only on tiny cases to cross-check the C oracle."""
    x = _f32(points)
    B, N, C = x.shape
    out = np.zeros((B, N, k), dtype=np.int32)
    f = np.float32
    for b in range(B):
        s = []
        for i in range(N):
            acc = f(0)
            for c in range(C):
                acc = f(acc + f(x[b, i, c] * x[b, i, c]))
            s.append(acc)
        for i in range(N):
            row = []
            for j in range(N):
                p = f(0)
                for c in range(C):
                    # exact product in fp64 (24x24 bits) + one rounding == fmaf when the
                    # fp64 sum is exact; tiny cases use small-integer / short-mantissa data
                    p = f(np.float64(x[b, i, c]) * np.float64(x[b, j, c]) + np.float64(p))
                row.append((f(f(s[i] + s[j]) - f(f(2) * p)), j))
            row.sort(key=lambda t: (t[0], t[1]))
            out[b, i] = [j for _, j in row[:k]]
    return out


# ----------------------------------------------------------------------------- edges
def edges(points: torch.Tensor, k: int = 20, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ops.py:21-40 (= get_edge_feature).  -> [B,N,k,2C]."""
    if idx is None:
        idx = k_nn(points, k)
    B, N, C = points.shape
    base = (torch.arange(B) * N).reshape(B, 1, 1)  # ops.py:30-31
    flat = points.reshape(-1, C)  # ops.py:33
    nbr = flat[(idx.long() + base).reshape(-1)].reshape(B, N, k, C)  # ops.py:34
    ctr = points.unsqueeze(-2).expand(B, N, k, C)  # ops.py:35-37
    return torch.cat([ctr, nbr - ctr], dim=-1)  # ops.py:39


# ----------------------------------------------------------------------------- layers
def bn_train(t: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """slim.batch_norm defaults [TF-default]: is_training=True, center=True (beta), scale=False,
    epsilon=1e-3; batch mean and BIASED batch variance over every axis but channels."""
    dims = tuple(range(t.dim() - 1))
    mean = t.mean(dim=dims, keepdim=True)
    var = ((t - mean) ** 2).mean(dim=dims, keepdim=True)
    return (t - mean) / torch.sqrt(var + BN_EPS) + beta


class _BF16MatMul(torch.autograd.Function):
    """matmul whose operands (forward: input and weights; backward: the incoming gradient too) are rounded to bf16 and
    whose accumulation is fp32 -- what a bf16 tensor-core execution of the reference's 1x1 convolutions computes."""

    @staticmethod
    def forward(ctx, a, w):
        ab, wb = a.bfloat16().to(a.dtype), w.bfloat16().to(w.dtype)
        ctx.save_for_backward(ab, wb)
        return ab @ wb

    @staticmethod
    def backward(ctx, g):
        ab, wb = ctx.saved_tensors
        gb = g.bfloat16().to(g.dtype)
        return gb @ wb.t(), ab.reshape(-1, ab.shape[-1]).t() @ gb.reshape(-1, gb.shape[-1])


_bf16_matmul = False


class bf16_matmul(object):
    """with bf16_matmul(): every conv_bn inside runs its matmul on bf16-rounded operands (the yardstick for the
    reduced-precision variant of BASELINE.json configs[2]: how far a bf16 execution of the reference graph itself is
    from the fp32 one)."""

    def __enter__(self):
        global _bf16_matmul
        self._old, _bf16_matmul = _bf16_matmul, True

    def __exit__(self, *a):
        global _bf16_matmul
        _bf16_matmul = self._old


def conv_bn(t: torch.Tensor, P: Dict[str, torch.Tensor], scope: str, relu: bool = True) -> torch.Tensor:
    """slim.conv2d(kernel 1, stride 1, VALID, normalizer_fn=batch_norm): 1x1 conv == matmul on the
    channel axis, no bias, BN, then activation_fn (default relu) AFTER BN [TF-default]."""
    z = _BF16MatMul.apply(t, P[scope + "/weights"]) if _bf16_matmul else torch.matmul(t, P[scope + "/weights"])
    y = bn_train(z, P[scope + "/BatchNorm/beta"])
    return torch.relu(y) if relu else y


def edge_conv(point_cloud, k, P, scope, relu_out=True, idx=None) -> List[torch.Tensor]:
    """ops.py:42-73 -> [net_max, net_mean, net], each [B,N,1,*]."""
    net = edges(point_cloud, k=k, idx=idx)  # ops.py:45
    net = conv_bn(net, P, scope + "/conv0", relu=True)  # ops.py:47-54
    net_max = net.amax(dim=-2, keepdim=True)  # ops.py:56
    net_mean = net.mean(dim=-2, keepdim=True)  # ops.py:57
    net = torch.cat([net_max, net_mean], dim=-1)  # ops.py:58
    net = conv_bn(net, P, scope + "/conv1", relu=relu_out)  # ops.py:62-70
    return [net_max, net_mean, net]


def _listify(v, repeat, what):
    """ops.py:78-87: scalar -> list broadcast, wrong length -> ValueError."""
    if not isinstance(v, list):
        return [int(v)] * repeat
    if len(v) != repeat:
        raise ValueError("Length of %s != repeat" % what)
    return v


def repeat_edge_conv(point_cloud, repeat, k, num_filters, P, idx_list=None, idx_out=None):
    """ops.py:75-98."""
    repeat = int(repeat)
    k = _listify(k, repeat, "k")
    num_filters = _listify(num_filters, repeat, "num_filters")
    net, tensors = point_cloud, []
    for i in range(repeat):
        idx = idx_list[i] if idx_list is not None else k_nn(net, k[i])
        if idx_out is not None:
            idx_out.append(idx)
        tensors += edge_conv(net, k[i], P, "EdgeConv%d" % i, idx=idx)
        net = tensors[-1].squeeze(-2)
    return tensors


def repeat_residual_edge_conv(point_cloud, repeat, k, num_filters, P, idx_list=None, idx_out=None):
    """ops.py:100-140."""
    repeat = int(repeat)
    k = _listify(k, repeat, "k")
    num_filters = _listify(num_filters, repeat, "num_filters")
    net, tensors, shortcut = point_cloud, [], None
    for i in range(repeat):
        scope = "EdgeConv%d" % i
        idx = idx_list[i] if idx_list is not None else k_nn(net, k[i])
        if idx_out is not None:
            idx_out.append(idx)
        if shortcut is None:
            tensors += edge_conv(net, k[i], P, scope, idx=idx)
        else:
            tensors += edge_conv(net, k[i], P, scope, relu_out=False, idx=idx)
            if num_filters[i] != num_filters[i - 1]:
                shortcut = conv_bn(shortcut, P, scope + "/shortcut", relu=False)  # ops.py:124-133
            tensors[-1] = torch.relu(shortcut + tensors[-1])  # ops.py:134
        net = tensors[-1]
        shortcut = tensors[-1]
        net = net.squeeze(-2)
    return tensors


def fc(net, repeat, num_filters, P):
    """ops.py:142-163."""
    repeat = int(repeat)
    num_filters = _listify(num_filters, repeat, "num_filters")
    for i in range(repeat):
        net = conv_bn(net, P, "FC%d" % i, relu=True)
    return net


# ----------------------------------------------------------------------------- model
def make_flags(**kw) -> SimpleNamespace:
    """Defaults of /root/reference/dgcnn/flags.py:9-45 (only what build()/trainval read)."""
    d = dict(NUM_CLASS=2, MODEL_NAME="dgcnn", TRAIN=True, KVALUE=20, DEBUG=False, EDGE_CONV_LAYERS=3,
             EDGE_CONV_FILTERS=64, FC_LAYERS=2, FC_FILTERS=[512, 256], LEARNING_RATE=0.001, GPUS=[0],
             MINIBATCH_SIZE=1, NUM_CHANNEL=3, WEIGHT_KEY="", SEED=0)
    d.update(kw)
    return SimpleNamespace(**d)


def param_shapes(flags, C0: int) -> Dict[str, tuple]:
    """Variable inventory, SURVEY.md Appendix A (TF scope names; conv weights stored 2-D [Cin,Cout])."""
    L = int(flags.EDGE_CONV_LAYERS)
    filt = _listify(flags.EDGE_CONV_FILTERS, L, "num_filters")
    shapes: Dict[str, tuple] = {}

    def conv(scope, cin, cout):
        shapes[scope + "/weights"] = (cin, cout)
        shapes[scope + "/BatchNorm/beta"] = (cout,)

    cin = C0
    for i in range(L):
        conv("EdgeConv%d/conv0" % i, 2 * cin, filt[i])
        conv("EdgeConv%d/conv1" % i, 2 * filt[i], CONV1_WIDTH)
        if flags.MODEL_NAME != "dgcnn" and i > 0 and filt[i] != filt[i - 1]:
            conv("EdgeConv%d/shortcut" % i, CONV1_WIDTH, filt[i])
        cin = CONV1_WIDTH
    if flags.MODEL_NAME == "residual-dgcnn-nofc":
        conv("Final", CONV1_WIDTH, int(flags.NUM_CLASS))
        return shapes
    conv("MergedEdgeConv", CONV1_WIDTH * L, 1024)
    width = 1024 + sum(2 * f + CONV1_WIDTH for f in filt) + 1024
    nfc = int(flags.FC_LAYERS)
    fcf = _listify(flags.FC_FILTERS, nfc, "num_filters")
    for j in range(nfc):
        conv("FC%d" % j, width, fcf[j])
        width = fcf[j]
    conv("Final", width, int(flags.NUM_CLASS))
    return shapes


def init_params(flags, C0: int, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """xavier_initializer() = uniform +-sqrt(6/(Cin+Cout)) for weights, zeros for beta [TF-default]."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in param_shapes(flags, C0).items():
        if name.endswith("/weights"):
            lim = math.sqrt(6.0 / (shp[0] + shp[1]))
            P[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
        else:
            P[name] = torch.zeros(shp, dtype=dtype)
    return P


def build(point_cloud, flags, P, idx_list=None, idx_out=None, dropout_mask=None, tensors_out=None):
    """model.py:9-106 -> logits [B,N,NUM_CLASS] (ReLU'd: reference quirk, model.py:94-101).

    idx_list: optional per-layer kNN indices to use instead of recomputing (teacher forcing for
    parity tests: a near-tie flip in feature space is a discontinuity, not an error).
    dropout_mask: the {0,1} keep mask [B,N,1,width] when flags.TRAIN; None -> drawn with torch RNG.
    """
    L = int(flags.EDGE_CONV_LAYERS)
    k = int(flags.KVALUE)
    net = point_cloud
    B, N = net.shape[0], net.shape[1]
    if flags.MODEL_NAME == "dgcnn":
        tensors = repeat_edge_conv(net, L, k, flags.EDGE_CONV_FILTERS, P, idx_list, idx_out)
    elif flags.MODEL_NAME in ("residual-dgcnn", "residual-dgcnn-nofc"):
        tensors = repeat_residual_edge_conv(net, L, k, flags.EDGE_CONV_FILTERS, P, idx_list, idx_out)
    else:
        raise NotImplementedError("Unsupported MODEL_NAME: %s" % flags.MODEL_NAME)  # model.py:41-43
    if tensors_out is not None:
        tensors_out.extend(tensors)
    if flags.MODEL_NAME == "residual-dgcnn-nofc":
        return conv_bn(tensors[-1], P, "Final", relu=True).squeeze(-2)  # model.py:45-58
    concat = torch.cat([tensors[3 * i + 2] for i in range(L)], dim=-1)  # model.py:60-63
    net = conv_bn(concat, P, "MergedEdgeConv", relu=True)  # model.py:65-72
    tensors = tensors + [net]  # model.py:74
    g = net.amax(dim=1, keepdim=True)  # model.py:77 max_pool_v2 ksize [1,N,1,1]
    g = g.reshape(B, -1, 1, 1024).expand(B, N, 1, 1024)  # model.py:80-81
    net = torch.cat([g] + tensors, dim=3)  # model.py:83-85
    net = fc(net, flags.FC_LAYERS, flags.FC_FILTERS, P)  # model.py:88
    if bool(flags.TRAIN):  # model.py:90-91
        if dropout_mask is None:
            dropout_mask = (torch.rand(net.shape) < DROPOUT_KEEP).to(net.dtype)
        net = net * dropout_mask / DROPOUT_KEEP
    net = conv_bn(net, P, "Final", relu=True)  # model.py:94-101
    return net.squeeze(-2)  # model.py:104


def softmax_loss_accuracy(logits, labels, weight=None):
    """trainval.py:39-52: softmax, accuracy = mean(argmax==label), loss = mean(xent [* weight])."""
    softmax = torch.softmax(logits, dim=-1)
    acc = (logits.argmax(dim=2) == labels.long()).to(torch.float32).mean()
    xent = torch.nn.functional.cross_entropy(
        logits.reshape(-1, logits.shape[-1]), labels.reshape(-1).long(), reduction="none"
    ).reshape(labels.shape)
    if weight is not None:
        xent = xent * weight
    return softmax, acc, xent.mean()


def adam_tf_step(p, g, m, v, t: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update [TF-default]: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    p -= lr_t * m / (sqrt(v) + eps)   (epsilon OUTSIDE the bias correction)."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    p.sub_(lr_t * m / (v.sqrt() + eps))


def train_step_reference(points, labels, flags, P, weight=None, dropout_mask=None, exact_knn=False):
    """One fwd+bwd of the whole graph the way trainval.py:38-54 builds it (single tower).
    exact_knn=False is the TF-literal matmul+top_k path that the CPU baseline times."""
    for t in P.values():
        t.requires_grad_(True)
        t.grad = None
    idx_list = None
    if not exact_knn:
        global k_nn
        saved = k_nn
        k_nn = lambda pts, kk, exact=True: saved(pts.detach(), kk, exact=False)  # noqa: E731
    try:
        logits = build(points, flags, P, idx_list=idx_list, dropout_mask=dropout_mask)
    finally:
        if not exact_knn:
            k_nn = saved
    _, acc, loss = softmax_loss_accuracy(logits, labels, weight)
    loss.backward()
    return loss.item(), acc.item(), {n: t.grad for n, t in P.items()}
