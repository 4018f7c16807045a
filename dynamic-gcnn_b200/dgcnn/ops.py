"""dgcnn.ops -- B200-native mirror of /root/reference/dgcnn/ops.py (same names, arguments, error behaviour).

    k_nn(points, k)                       ops.py:8    -> idx [B,N,k] int32
    edges(points, k=20)                   ops.py:21   -> [B,N,k,2C]
    edge_conv(point_cloud, k, num_filters, trainable, activation=relu, debug=False)   ops.py:42
    repeat_edge_conv(...)                 ops.py:75
    repeat_residual_edge_conv(...)        ops.py:100
    fc(net, repeat, num_filters, trainable, debug=False)                              ops.py:142
plus the names BASELINE.json's north_star uses for the same surface:
    pairwise_distance(points) -> [B,N,N] ; knn(adj_matrix, k) -> idx ; get_edge_feature(points, nn_idx, k)

Every hot-path op calls hand-written sm_100a CUDA through the C ABI in include/dgcnn_b200.h
(ctypes binding in _native.py).  Inputs are torch CUDA fp32 tensors; there is no CPU path.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import _native as nv
from .variables import default_store, variable_scope

class PlaneSinks:
    """Producer-side filling of the head's tensor-core operands (model.py:60-63,83-85: the concats).

    dgcnn.model.build pre-allocates the bf16 hi/lo operand planes of MergedEdgeConv, FC0 and FC1 and publishes, per
    producer, WHERE its output lives inside them.  The producing kernels (EdgeConv gather apply, BN apply) then write
    their result there as well, and the consuming layer skips the split pass for every source that `done` lists.
    `done` maps (data_ptr, columns, row stride) of a source tensor to (data_ptr of the operand planes, column)."""

    def __init__(self):
        self.targets = {}   # producer key -> list of (planes [2,P,K] bf16, column)
        self.planes = {}    # consumer scope -> planes
        self.done = {}
        self.where = {}     # (data_ptr, columns, row stride) -> (planes, column) of one operand that holds the tensor
        self.layer = 0

    def sink_args(self, key):
        """-> (n, ptr array, ld array, plane array) for the C ABI, plus the raw list"""
        import ctypes
        lst = self.targets.get(key, [])
        n = len(lst)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[pl.data_ptr() + 2 * col for pl, col in lst])
        lds = (ctypes.c_int * max(n, 1))(*[pl.shape[2] for pl, _ in lst])
        pes = (ctypes.c_int64 * max(n, 1))(*[_pe(pl) for pl, _ in lst])
        return n, ptrs, lds, pes, lst

    def mark(self, t2d_ptr, cols, ld, lst):
        for pl, col in lst:
            self.done[(t2d_ptr, cols, ld, pl.data_ptr())] = col
            self.where[(t2d_ptr, cols, ld)] = (pl, col)

    def find(self, t2d):
        """-> (planes, column) if some operand already holds the bf16 planes of this [rows, cols] tensor"""
        return self.where.get((t2d.data_ptr(), t2d.shape[1], t2d.stride(0)))


_sinks = None   # set by dgcnn.model.build for the duration of one forward pass


relu = "relu"  # stand-in for tf.nn.relu as the `activation` argument (None = linear, ops.py:121)

# Arithmetic mode of the 1x1 convolutions, set by dgcnn.model.build from flags.DTYPE for the duration of a forward pass:
#   "f32"  : fp32-faithful -- tensor-core operands are two bf16 planes (hi, lo), 3 MMAs per product; uv / conv1 forward
#            on the exact fp32 SIMT GEMM; fp32 uv table
#   "bf16" : BASELINE.json configs[2] -- single-plane bf16 operands, 1 MMA per product, every 1x1 conv on the tensor
#            cores, bf16 uv table (half the gather bytes).  Statistics, activations and k_nn stay fp32.
_precision = "f32"
_bf16_table = torch.float16   # bf16 mode: storage type of the gathered uv table (None: fp32)
_bf16_uv_gemm = True    # bf16 mode: the uv GEMM itself in bf16 (else exact fp32)


def _npl() -> int:
    return 1 if _precision == "bf16" else 2


def _pe(planes: torch.Tensor) -> int:
    """distance of the lo plane in elements (0 = the operand has a hi plane only)"""
    return planes.shape[1] * planes.shape[2] if planes.shape[0] == 2 else 0

BN_EPS = 1e-3
CONV1_WIDTH = 64  # ops.py:63


# test hooks (never set by the product): _knn_trace collects each layer's kNN indices; _knn_forced supplies them
_knn_trace = None
_knn_input_trace = None   # collects the tensor each layer's kNN was computed on (the GPU's own activations)
_knn_forced = None
# bench.py hooks: when a list, k_nn() / the head's forward GEMM bracket their C-ABI call with CUDA events on the
# launching stream (eager mode only: a captured graph cannot carry timing events)
_knn_events = None
_gemm_events = None
_ec_events = None
_ec_bwd_events = None


def _layer_knn(x, k, hint=None):
    if _knn_forced is not None:
        idx = next(_knn_forced).to(x.device, torch.int32).contiguous()
    else:
        idx = k_nn(x, k, hint=hint)
    if _knn_trace is not None:
        _knn_trace.append(idx)
    if _knn_input_trace is not None:
        _knn_input_trace.append(x.detach())
    return idx


# =============================================================================== weight gradients on a side stream
# A weight gradient (X^T . g) has no consumer inside the backward pass: only the optimizer reads it.  When the trainer
# enables _async_dw it is launched on a second stream, forked after the kernel that produced g: the tensor-bound dW
# GEMMs of the head then overlap the HBM-bound BatchNorm-backward / EdgeConv-gather kernels of the layers below instead
# of extending the critical path (in a captured CUDA graph the fork / join become graph edges).  The caller of backward
# joins with join_side_stream() before it reads any gradient; generic autograd users leave the switch off.
_async_dw = False
_side_streams = {}
_side_keep = []       # tensors the side stream still reads: kept alive until the join (no allocator reuse under it)


class _SideStream(object):
    """with _SideStream(dev, keep...): the body's launches go to the side stream, ordered after everything already
    queued on the current stream.  No-op unless _async_dw."""

    def __init__(self, dev, *keep):
        self.dev, self.keep, self.ctx = dev, keep, None

    def __enter__(self):
        if _async_dw:
            key = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
            side = _side_streams.get(key)
            if side is None:
                side = _side_streams[key] = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            _side_keep.extend(t for t in self.keep if t is not None)
            self.ctx = torch.cuda.stream(side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def ws_tag(base: str) -> str:
    """scratch buffers are per stream: the side stream must not share the main stream's split-K workspace"""
    return base + "_side" if (_async_dw and torch.cuda.current_stream() in _side_streams.values()) else base


def join_side_stream(dev) -> None:
    """current stream waits for every weight gradient launched on the side stream"""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    side = _side_streams.get(key)
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)
    del _side_keep[:]


# =============================================================================== raw kernels
def k_nn(points: torch.Tensor, k: int, hint: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ops.py:8-19.  Fused distance + top-k; the [B,N,N] matrix is never written.  Not differentiable
    (only the indices are used downstream, ops.py:19,34).

    hint (optional, [B,N,>=k] int32, DISTINCT indices per row -- e.g. the previous layer's result) warm-starts the
    selection threshold of the SIMT path (small clouds / more than 64 channels); the tensor-core path derives a tighter
    bound from its own first sweep.  The result is identical with or without it."""
    x = nv.require_cuda(points.detach(), "points")
    if x.dim() != 3:
        raise ValueError("k_nn: points must be [B,N,C]")
    B, N, C = x.shape
    k = int(k)
    L = nv.lib()
    need = L.dgcnn_knn_workspace_bytes(B, N, C)
    ws = nv.workspace(x.device, need, "knn")
    idx = torch.empty((B, N, k), dtype=torch.int32, device=x.device)
    ev = None
    if _knn_events is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    hp = 0
    if hint is not None and hint.shape[-1] >= k and hint.shape[:2] == (B, N):
        hint = nv.require_cuda(hint if hint.shape[-1] == k else hint[:, :, :k], "hint", torch.int32)
        hp = hint.data_ptr()
    nv.check(L.dgcnn_knn_mode(x.data_ptr(), hp, idx.data_ptr(), B, N, C, k, _KNN_FILTER_MODE, ws.data_ptr(), ws.numel(),
                              nv.stream_ptr(x.device)), "k_nn")
    if ev is not None:
        ev[1].record()
        _knn_events.append((B, N, C, k, ev[0], ev[1]))
    return idx


def pairwise_distance(points: torch.Tensor) -> torch.Tensor:
    """ops.py:11-16 materialised: D[b,i,j] = (s_i + s_j) - 2 x_i.x_j  -> [B,N,N]."""
    x = nv.require_cuda(points.detach(), "points")
    if x.dim() != 3:
        raise ValueError("pairwise_distance: points must be [B,N,C]")
    B, N, C = x.shape
    L = nv.lib()
    ws = nv.workspace(x.device, L.dgcnn_knn_workspace_bytes(B, N, C), "knn")
    D = torch.empty((B, N, N), dtype=torch.float32, device=x.device)
    nv.check(L.dgcnn_pairwise_distance(x.data_ptr(), D.data_ptr(), B, N, C, ws.data_ptr(), ws.numel(),
                                       nv.stream_ptr(x.device)), "pairwise_distance")
    return D


def knn(adj_matrix: torch.Tensor, k: int = 20) -> torch.Tensor:
    """ops.py:18 alone (north-star name): k smallest entries per row of a distance matrix, ascending,
    ties -> lower index.  adj_matrix [..., N] -> idx [..., k] int32."""
    D = nv.require_cuda(adj_matrix.detach(), "adj_matrix")
    N = D.shape[-1]
    rows = D.numel() // N
    idx = torch.empty(D.shape[:-1] + (int(k),), dtype=torch.int32, device=D.device)
    nv.check(nv.lib().dgcnn_topk_rows(D.data_ptr(), idx.data_ptr(), rows, N, int(k), nv.stream_ptr(D.device)), "knn")
    return idx


class _EdgeFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        x = nv.require_cuda(points, "points")
        ix = nv.require_cuda(idx, "idx", torch.int32)
        B, N, C = x.shape
        k = ix.shape[-1]
        out = torch.empty((B, N, k, 2 * C), dtype=torch.float32, device=x.device)
        nv.check(nv.lib().dgcnn_edge_feature(x.data_ptr(), ix.data_ptr(), out.data_ptr(), B, N, C, k,
                                             nv.stream_ptr(x.device)), "edges")
        ctx.save_for_backward(ix)
        ctx.shape = (B, N, C, k)
        return out

    @staticmethod
    def backward(ctx, g):
        (ix,) = ctx.saved_tensors
        B, N, C, k = ctx.shape
        g = nv.require_cuda(g, "grad")
        gx = torch.empty((B, N, C), dtype=torch.float32, device=g.device)
        nv.check(nv.lib().dgcnn_edge_feature_bwd(g.data_ptr(), ix.data_ptr(), gx.data_ptr(), B, N, C, k,
                                                 nv.stream_ptr(g.device)), "edges_bwd")
        return gx, None


def get_edge_feature(point_cloud: torch.Tensor, nn_idx: torch.Tensor, k: Optional[int] = None) -> torch.Tensor:
    """ops.py:30-39 with the indices supplied (north-star name) -> [B,N,k,2C]."""
    if k is not None and nn_idx.shape[-1] != int(k):
        raise ValueError("get_edge_feature: nn_idx has k=%d, argument says %d" % (nn_idx.shape[-1], int(k)))
    return _EdgeFeature.apply(point_cloud, nn_idx)


def edges(points: torch.Tensor, k: int = 20) -> torch.Tensor:
    """ops.py:21-40."""
    return _EdgeFeature.apply(points, k_nn(points, k))


# =============================================================================== differentiable blocks
def _gemm_raw(A, Bm, M, N, K, tA, tB):
    L = nv.lib()
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    need = L.dgcnn_gemm_workspace_bytes(M, N, K, tA, tB)
    ws = nv.workspace(A.device, need, ws_tag("gemm")) if need else None
    nv.check(L.dgcnn_gemm(A.data_ptr(), Bm.data_ptr(), out.data_ptr(), M, N, K, tA, tB, nv.ptr(ws),
                          ws.numel() if ws is not None else 0, nv.stream_ptr(A.device)), "gemm")
    return out


class _Conv1x1(torch.autograd.Function):
    """slim.conv2d(kernel_size=1, stride=1, VALID) on channels-last data = [P,Cin] x [Cin,Cout]."""

    @staticmethod
    def forward(ctx, a, w, exact=False):
        a = nv.require_cuda(a, "conv input")
        w = nv.require_cuda(w, "conv weights")
        ctx.save_for_backward(a, w)
        # if a producer already wrote this input as bf16 planes into some head operand, the weight gradient reuses them
        ctx.a_planes = _sinks.find(a) if _sinks is not None else None
        ctx.npl = _npl()
        P, Cin = a.shape
        Cout = w.shape[1]
        have = ctx.a_planes is not None and ctx.a_planes[0].shape[0] == ctx.npl
        if _tc_ok(P, Cin, Cout) and not exact and (ctx.npl == 1 or (have and _TC_FORWARD)):
            # bf16 mode: every 1x1 conv takes the tensor cores (one bf16 MMA per product).
            # f32 mode: exact SIMT kernel, unless the A/B switch _TC_FORWARD is on (see its comment).
            pw = _split(w, ctx.npl)
            if have:
                pl, col = ctx.a_planes
                return _tc_gemm_a_slice_raw(pl, col, pw, P, Cout, Cin, 0, 0)
            return _tc_gemm_raw(_split(a, ctx.npl), pw, P, Cout, Cin, 0, 0)
        return _gemm_raw(a, w, P, Cout, Cin, 0, 0)

    @staticmethod
    def backward(ctx, g):
        a, w = ctx.saved_tensors
        g = nv.require_cuda(g, "grad")
        P, Cin = a.shape
        Cout = w.shape[1]
        ga = gw = None
        npl = ctx.npl
        if _tc_ok(P, Cin, Cout):
            # f32 mode: the forward stays exact fp32 (the next layer's kNN is built on it); gradients feed no kNN, so
            # they take the tensor-core path like every other gradient GEMM of the model
            pg = _split(g, npl)
            if ctx.needs_input_grad[0]:
                ga = _tc_gemm_raw(pg, _split(w, npl), P, Cin, Cout, 0, 1)
            if ctx.needs_input_grad[1]:
                have = ctx.a_planes is not None and ctx.a_planes[0].shape[0] == npl
                with _SideStream(g.device, pg, a, ctx.a_planes[0] if have else None):
                    if have:
                        pl, col = ctx.a_planes
                        gw = _tc_gemm_a_slice_raw(pl, col, pg, Cin, Cout, P, 1, 0)
                    else:
                        gw = _tc_gemm_raw(_split(a, npl), pg, Cin, Cout, P, 1, 0)
            return ga, gw, None
        if ctx.needs_input_grad[1]:
            with _SideStream(g.device, a, g):
                gw = _gemm_raw(a, g, Cin, Cout, P, 1, 0)  # A^T . g  (split over the P points)
        if ctx.needs_input_grad[0]:
            ga = _gemm_raw(g, w, P, Cin, Cout, 0, 1)  # g . W^T
        return ga, gw, None


# ---- tensor-core path: tcgen05 GEMM on pre-split bf16 hi/lo planes (fp32-faithful, csrc/tc_gemm.cu) -------------
# A/B switch of the k_nn distance filter (host side; the library keeps no state): DGCNN_KNN_FINE=0|1 forces the coarse /
# fine precision mode, default = the library's rule.  Same result either way.
_KNN_FILTER_MODE = {"0": 0, "1": 1}.get(os.environ.get("DGCNN_KNN_FINE", ""), -1)

TC_MIN_ROWS = 1024  # below this the SIMT kernel wins (launch + split overhead)
# A/B switch (default OFF): conv1 forward on the fp32-faithful tensor-core GEMM, reading the planes its producer left in
# the FC0 operand.  Measured at configs[1]: 0.03 ms/step faster, but the 2^-16 operand error of `net` is amplified by the
# next layer's neighbour differences (x_j - x_i): same-graph logits 9e-5 -> 7e-4 from the fp32 oracle, gradients 6x
# further from fp64 than fp32 arithmetic is.  The EdgeConv stack therefore stays on exact fp32 GEMMs in f32 mode.
_TC_FORWARD = os.environ.get("DGCNN_TC_FORWARD", "0") == "1"
_FUSE_POOL = os.environ.get("DGCNN_FUSE_POOL", "1") != "0"     # A/B switch: global max pool fused into MergedEdgeConv's BN


def _tc_ok(P: int, *dims: int) -> bool:
    return P >= TC_MIN_ROWS and P % 8 == 0 and all(d % 8 == 0 and d >= 8 for d in dims)


def _split_into(x2d: torch.Tensor, planes: torch.Tensor, col: int) -> None:
    """fp32 [rows, c] -> columns [col, col+c) of the bf16 operand planes [2, rows, ld]."""
    rows, c = x2d.shape
    ld = planes.shape[2]
    dst = planes.data_ptr() + 2 * col
    nv.check(nv.lib().dgcnn_split_bf16(x2d.data_ptr(), rows, c, x2d.stride(0), dst, ld, _pe(planes),
                                       nv.stream_ptr(x2d.device)), "split_bf16")


def _split(x2d: torch.Tensor, npl: Optional[int] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 operand planes [npl, rows, cols] (npl: 2 = hi/lo, 1 = plain bf16; default: the mode)"""
    x2d = nv.require_cuda(x2d, "operand")
    planes = torch.empty((npl or _npl(),) + tuple(x2d.shape), dtype=torch.bfloat16, device=x2d.device)
    _split_into(x2d, planes, 0)
    return planes


def _tc_gemm_raw(pa, pb, M, N, K, tA, tB):
    L = nv.lib()
    out = torch.empty((M, N), dtype=torch.float32, device=pa.device)
    need = L.dgcnn_tc_gemm_workspace_bytes(M, N, K)
    ws = nv.workspace(pa.device, need, ws_tag("gemm")) if need else None
    assert pa.shape[0] == pb.shape[0], "tc_gemm: both operands must be in the same precision mode"
    nv.check(L.dgcnn_tc_gemm(pa.data_ptr(), pb.data_ptr(), out.data_ptr(), M, N, K, tA, tB, pa.shape[0], nv.ptr(ws),
                             ws.numel() if ws is not None else 0, nv.stream_ptr(pa.device)), "tc_gemm")
    return out


def _tc_gemm_a_slice_raw(pl, col, pb, M, N, K, tA, tB):
    """op(A) = the column slice starting at `col` of the wider operand `pl` (filled by its producer)"""
    L = nv.lib()
    out = torch.empty((M, N), dtype=torch.float32, device=pb.device)
    need = L.dgcnn_tc_gemm_workspace_bytes(M, N, K)
    ws = nv.workspace(pb.device, need, ws_tag("gemm")) if need else None
    assert pl.shape[0] == pb.shape[0], "tc_gemm_a_slice: both operands must be in the same precision mode"
    nv.check(L.dgcnn_tc_gemm_a_slice(pl.data_ptr() + 2 * col, pl.shape[2], _pe(pl), pb.data_ptr(), out.data_ptr(), M, N, K,
                                     tA, tB, pb.shape[0], nv.ptr(ws), ws.numel() if ws is not None else 0,
                                     nv.stream_ptr(pb.device)), "tc_gemm_a_slice")
    return out


def _tc_dx_sources(pg, pw, P, K, Cout, widths, needs):
    """g . W^T for a multi-source (concatenated) operand -> one dense gradient tensor per source."""
    if not any(needs):
        return [None] * len(widths)
    if len(widths) > 1 and len(widths) <= 32 and all(c % 32 == 0 for c in widths) and \
            nv.lib().dgcnn_tc_gemm_workspace_bytes(P, K, Cout) == 0:
        import ctypes
        n = len(widths)
        bufs = [torch.empty((P, c), dtype=torch.float32, device=pg.device) for c in widths]
        starts = (ctypes.c_int * n)(*[sum(widths[:i]) for i in range(n)])
        cw = (ctypes.c_int * n)(*widths)
        ptrs = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bufs])
        nv.check(nv.lib().dgcnn_tc_gemm_grouped(pg.data_ptr(), pw.data_ptr(), P, K, Cout, 0, 1, pg.shape[0], n, starts, cw,
                                                ptrs, nv.stream_ptr(pg.device)), "tc_gemm_grouped")
        return [b if needs[i] else None for i, b in enumerate(bufs)]
    gcat = _tc_gemm_raw(pg, pw, P, K, Cout, 0, 1)                                     # g . W^T
    outs, off = [], 0
    for i, c in enumerate(widths):
        outs.append(gcat[:, off:off + c] if needs[i] else None)
        off += c
    return outs


def _split_sources(srcs, w, planes=None):
    """sources -> column slices of one bf16 hi/lo operand.  `planes` (optional): a pre-allocated operand that some
    producers may already have filled (PlaneSinks.done); those sources are skipped."""
    P = srcs[0].shape[0]
    widths = [int(t.shape[1]) for t in srcs]
    K = sum(widths)
    if planes is None or tuple(planes.shape) != (_npl(), P, K) or _sinks is None:
        planes = torch.empty((_npl(), P, K), dtype=torch.bfloat16, device=w.device)
        done = {}
    else:
        done = _sinks.done
    off = 0
    for t, c in zip(srcs, widths):
        if done.get((t.data_ptr(), c, t.stride(0), planes.data_ptr())) != off:
            _split_into(t, planes, off)
        off += c
    return planes, widths, P, K


class _ConcatConvTC(torch.autograd.Function):
    """1x1 conv of the channel-concatenation of several [P, c_i] tensors with W [sum c_i, Cout], without building
    the concatenation (model.py:60-63,83-85 + slim.conv2d): each source is split straight into its column slice of
    one bf16 operand; forward, dX and dW are three tcgen05 GEMMs on the same planes."""

    @staticmethod
    def forward(ctx, w, *srcs):
        w = nv.require_cuda(w, "conv weights")
        srcs = [nv.require_cuda_rows(t, "conv input") for t in srcs]
        planes, widths, P, K = _split_sources(srcs, w)
        Cout = w.shape[1]
        pw = _split(w)
        ctx.save_for_backward(planes, pw)
        ctx.widths = widths
        return _tc_gemm_raw(planes, pw, P, Cout, K, 0, 0)

    @staticmethod
    def backward(ctx, g):
        planes, pw = ctx.saved_tensors
        _, P, K = planes.shape
        Cout = pw.shape[2]
        pg = _split(nv.require_cuda(g, "grad"), planes.shape[0])
        gw = _tc_gemm_raw(planes, pg, K, Cout, P, 1, 0) if ctx.needs_input_grad[0] else None   # X^T . g
        return tuple([gw] + _tc_dx_sources(pg, pw, P, K, Cout, ctx.widths, ctx.needs_input_grad[1:]))


class _ConvBnActTC(torch.autograd.Function):
    """One whole head layer (model.py:65-72, ops.py:151-160): 1x1 conv of a channel-concatenation + train-mode
    BatchNorm [+ per-cloud bias] [+ ReLU], fused around the tcgen05 GEMM:
      forward : sources -> bf16 planes -> GEMM whose epilogue also leaves the BN column statistics of every 128-row
                tile -> tiny finalisation -> one apply pass.  (No separate statistics pass over z.)
      backward: BN backward writes g_z directly as bf16 planes (the operand format of the dW / dX GEMMs).
    pool_rows > 0 (MergedEdgeConv, model.py:76-77): the layer also returns the max over every pool_rows consecutive rows
    (one cloud); the apply pass produces it on the way, the fp32 output is not even written when every consumer reads
    the operand planes, and the pool's gradient is added inside the BN-backward kernels.  -> (out, pooled or None)"""

    @staticmethod
    def forward(ctx, w, beta, gb, relu_flag, grows, scope, pool_rows, *srcs):
        w = nv.require_cuda(w, "conv weights")
        beta = nv.require_cuda(beta, "beta")
        gb = nv.require_cuda(gb, "group_bias") if gb is not None else None
        srcs = [nv.require_cuda_rows(t, "conv input") for t in srcs]
        sk = _sinks
        planes, widths, P, K = _split_sources(srcs, w, sk.planes.get(scope) if sk is not None else None)
        n_s, s_ptr, s_ld, s_pe, s_lst = sk.sink_args(("layer", scope)) if sk is not None else (0, None, None, None, [])
        Cout = w.shape[1]
        pw = _split(w, planes.shape[0])
        dev = w.device
        L = nv.lib()
        st = nv.stream_ptr(dev)
        grows = int(grows) if gb is not None else 0
        pool_rows = int(pool_rows)
        mean = torch.empty(Cout, dtype=torch.float32, device=dev)
        rstd = torch.empty(Cout, dtype=torch.float32, device=dev)
        out = torch.empty((P, Cout), dtype=torch.float32, device=dev)
        pmax = pcnt = None
        fused = L.dgcnn_tc_gemm_stats_supported(P, Cout, K) and (gb is None or grows % 128 == 0)
        if pool_rows and not (fused and gb is None and relu_flag and P % pool_rows == 0):
            raise ValueError("_ConvBnActTC: a fused pool needs the statistics GEMM path, ReLU and no per-cloud bias")
        if fused:
            z = torch.empty((P, Cout), dtype=torch.float32, device=dev)
            tiles = (P + 127) // 128
            cs = torch.empty((tiles, 2, Cout), dtype=torch.float32, device=dev)
            ev = None
            if _gemm_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            nv.check(L.dgcnn_tc_gemm_stats(planes.data_ptr(), pw.data_ptr(), z.data_ptr(), P, Cout, K, 0, 0, planes.shape[0],
                                           cs.data_ptr(), st), "tc_gemm_stats")
            if ev is not None:
                ev[1].record()
                _gemm_events.append((P, Cout, K, ev[0], ev[1]))
            nv.check(L.dgcnn_bn_stats_from_tiles(cs.data_ptr(), tiles, Cout, P, nv.ptr(gb), grows, mean.data_ptr(),
                                                 rstd.data_ptr(), st), "bn_stats_from_tiles")
            if pool_rows:
                G = P // pool_rows
                pmax = torch.empty((G, Cout), dtype=torch.float32, device=dev)
                pcnt = torch.empty((G, Cout), dtype=torch.float32, device=dev)
                pws = nv.workspace(dev, L.dgcnn_bn_pool_workspace_bytes(G, Cout), "pool")
                # when the planes take the output (n_s > 0) nobody reads the fp32 copy: `out` stays an unwritten handle
                nv.check(L.dgcnn_bn_apply_fwd_pool(z.data_ptr(), P, Cout, beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                   0 if n_s else out.data_ptr(), pool_rows, pmax.data_ptr(), pcnt.data_ptr(),
                                                   pws.data_ptr(), pws.numel(), n_s, s_ptr, s_ld, s_pe, st),
                         "bn_apply_fwd_pool")
            else:
                # a layer whose output goes to plane sinks has only tensor-core consumers (model._make_sinks): like the
                # pooled layer above, its fp32 copy is then never read and `out` stays an unwritten handle
                nv.check(L.dgcnn_bn_apply_fwd_sinks(z.data_ptr(), P, Cout, beta.data_ptr(), 0, nv.ptr(gb), grows,
                                                    int(bool(relu_flag)), mean.data_ptr(), rstd.data_ptr(),
                                                    0 if (n_s and _FUSE_POOL) else out.data_ptr(), n_s, s_ptr, s_ld, s_pe,
                                                    st), "bn_apply_fwd")
        else:
            z = _tc_gemm_raw(planes, pw, P, Cout, K, 0, 0)
            ws = nv.workspace(dev, L.dgcnn_bn_workspace_bytes(Cout), "stats")
            nv.check(L.dgcnn_bn_act_fwd_sinks(z.data_ptr(), P, Cout, beta.data_ptr(), 0, nv.ptr(gb), grows,
                                              int(bool(relu_flag)), out.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                              ws.data_ptr(), ws.numel(), n_s, s_ptr, s_ld, s_pe, st), "bn_act_fwd")
        if n_s:
            sk.mark(out.data_ptr(), Cout, Cout, s_lst)
        ctx.save_for_backward(planes, pw, z, beta, mean, rstd, gb, pmax, pcnt)   # the ReLU mask is re-evaluated from z
        ctx.widths, ctx.relu, ctx.grows, ctx.pool_rows = widths, bool(relu_flag), grows, pool_rows
        return out, pmax

    @staticmethod
    def backward(ctx, g, gpool):
        planes, pw, z, beta, mean, rstd, gb, pmax, pcnt = ctx.saved_tensors
        _, P, K = planes.shape
        Cout = pw.shape[2]
        dev = z.device
        g = nv.require_cuda(g, "grad") if g is not None else torch.zeros_like(z)
        if pmax is not None:
            gpool = nv.require_cuda(gpool, "grad pooled") if gpool is not None else torch.zeros_like(pmax)
        L = nv.lib()
        ws = nv.workspace(dev, L.dgcnn_bn_workspace_bytes(Cout), "stats")
        pg = torch.empty((planes.shape[0], P, Cout), dtype=torch.bfloat16, device=dev)
        gz = torch.empty_like(z) if gb is not None else None      # fp32 copy only for the per-cloud bias gradient
        gbeta = torch.empty(Cout, dtype=torch.float32, device=dev)
        nv.check(L.dgcnn_bn_act_bwd_planes(z.data_ptr(), 0, beta.data_ptr(), g.data_ptr(), P, Cout, mean.data_ptr(),
                                           rstd.data_ptr(), nv.ptr(gb), ctx.grows, int(ctx.relu), nv.ptr(gz),
                                           pg.data_ptr(), pg.shape[0], gbeta.data_ptr(), nv.ptr(pmax), nv.ptr(pcnt),
                                           nv.ptr(gpool) if pmax is not None else 0, ctx.pool_rows, ws.data_ptr(),
                                           ws.numel(), nv.stream_ptr(dev)), "bn_act_bwd_planes")
        gw = ggb = None
        with _SideStream(dev, planes, pg, gz):
            if ctx.needs_input_grad[0]:
                gw = _tc_gemm_raw(planes, pg, K, Cout, P, 1, 0)                                # X^T . g_z
        if gb is not None:
            ggb = gz.view(gb.shape[0], ctx.grows, Cout).sum(dim=1)
        return tuple([gw, gbeta, ggb, None, None, None, None] +
                     _tc_dx_sources(pg, pw, P, K, Cout, ctx.widths, ctx.needs_input_grad[7:]))


def conv1x1(srcs, w) -> torch.Tensor:
    """[P, sum c_i] x [sum c_i, Cout] on the tensor cores when the shape allows, else the SIMT kernel."""
    P = srcs[0].shape[0]
    if _tc_ok(P, w.shape[1], *[t.shape[1] for t in srcs]):
        return _ConcatConvTC.apply(w, *srcs)
    a = srcs[0] if len(srcs) == 1 else torch.cat(srcs, dim=1)
    return _Conv1x1.apply(a, w)


class _EdgeWeights(torch.autograd.Function):
    """conv0's weights W0 = [Wa ; Wb] ([2C, F], ops.py:47-54) -> the uv operand [Wa - Wb | Wb] ([C, 2F]) of
    [x_i, x_j - x_i].W0 = x_i.(Wa - Wb) + x_j.Wb.  Two small kernels each way (slicing / cat / neg through autograd
    took eleven)."""

    @staticmethod
    def forward(ctx, w0):
        C, F = w0.shape[0] // 2, w0.shape[1]
        wp = torch.empty((C, 2 * F), dtype=w0.dtype, device=w0.device)
        torch.sub(w0[:C], w0[C:], out=wp[:, :F])
        wp[:, F:].copy_(w0[C:])
        return wp

    @staticmethod
    def backward(ctx, g):
        C, F = g.shape[0], g.shape[1] // 2
        with _SideStream(g.device, g):               # g is a weight gradient: it may still be in flight on the side stream
            gw = torch.empty((2 * C, F), dtype=g.dtype, device=g.device)
            gw[:C].copy_(g[:, :F])                       # d/dWa
            torch.sub(g[:, F:], g[:, :F], out=gw[C:])    # d/dWb
        return gw


class _SplitWeightRows(torch.autograd.Function):
    """w [cin, cout] -> (w[:r], w[r:]) for a layer whose input is a concat handled in two parts (model.py:83-88: the tiled
    global feature and the per-point sources of FC0).  Exists for its backward: the two weight gradients are joined on the
    side stream they were produced on (plain slicing would route them through autograd kernels on the main stream)."""

    @staticmethod
    def forward(ctx, w, r):
        ctx.r, ctx.shape = int(r), tuple(w.shape)
        return w[:ctx.r], w[ctx.r:]

    @staticmethod
    def backward(ctx, g_top, g_rest):
        ref = g_top if g_top is not None else g_rest
        with _SideStream(ref.device, g_top, g_rest):
            gw = torch.empty(ctx.shape, dtype=ref.dtype, device=ref.device)
            if g_top is not None:
                gw[:ctx.r].copy_(g_top)
            else:
                gw[:ctx.r].zero_()
            if g_rest is not None:
                gw[ctx.r:].copy_(g_rest)
            else:
                gw[ctx.r:].zero_()
        return gw, None


class _EdgeConvGather(torch.autograd.Function):
    """ops.py:45-58 after the algebraic split z_ij = u_i + v_idx(i,j): BN(train)+ReLU+max_k/mean_k.
    -> (max [P,F], mean [P,F], both [P,2F]); max and mean are the two column halves of `both`, which therefore IS
    ops.py:58's concat(max, mean).  Two gather passes forward (statistics, apply), ONE backward: the backward BN sums
    follow from per-point quantities (csrc/edge.cu), and the gather pass sums the gradients of all three outputs."""

    @staticmethod
    def forward(ctx, uv, idx, beta, B, N, k, sink_key=None):
        uv = nv.require_cuda(uv, "uv")
        if _precision == "bf16" and _bf16_table is not None:
            uv = uv.to(_bf16_table)        # the gathered table in bf16: half the L2 gather bytes (fp32 arithmetic)
        idx = nv.require_cuda(idx, "idx", torch.int32)
        beta = nv.require_cuda(beta, "beta")
        P, F2 = uv.shape
        F = F2 // 2
        dev = uv.device
        L = nv.lib()
        st = nv.stream_ptr(dev)
        dt = {torch.bfloat16: nv.DT_BF16, torch.float16: nv.DT_F16}.get(uv.dtype, nv.DT_F32)
        ws = nv.workspace(dev, L.dgcnn_edgeconv_workspace_bytes(F), "stats")
        mean = torch.empty(F, dtype=torch.float32, device=dev)
        rstd = torch.empty(F, dtype=torch.float32, device=dev)
        need_bwd = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
        npos = torch.empty((P, F), dtype=torch.uint8, device=dev) if need_bwd else None
        zmax = torch.empty((P, F), dtype=torch.float32, device=dev) if need_bwd else None
        ev = None
        if _ec_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        nv.check(L.dgcnn_edgeconv_fwd_stats(uv.data_ptr(), dt, idx.data_ptr(), B, N, F, k, mean.data_ptr(),
                                            rstd.data_ptr(), ws.data_ptr(), ws.numel(), st), "edgeconv_fwd_stats")
        both = torch.empty((P, 2 * F), dtype=torch.float32, device=dev)
        sk = _sinks
        lst = sk.targets.get(sink_key, []) if (sk is not None and sink_key is not None) else []
        sink = None
        if len(lst) == 1 and tuple(lst[0][0].shape[1:2]) == (P,) and F % 4 == 0:
            sink = lst[0]
        sp, sld, spe = (0, 0, 0)
        if sink is not None:
            pl, col = sink
            sp, sld, spe = pl.data_ptr() + 2 * col, pl.shape[2], _pe(pl)
        nv.check(L.dgcnn_edgeconv_fwd_apply(uv.data_ptr(), dt, idx.data_ptr(), B, N, F, k, mean.data_ptr(),
                                            rstd.data_ptr(), beta.data_ptr(), both.data_ptr(), nv.ptr(zmax), nv.ptr(npos), sp,
                                            sld, spe, st), "edgeconv_fwd_apply")
        if sink is not None:
            pl, col = sink
            sk.mark(both.data_ptr(), F, 2 * F, [(pl, col)])                       # the max view
            sk.mark(both.data_ptr() + 4 * F, F, 2 * F, [(pl, col + F)])           # the mean view
            sk.mark(both.data_ptr(), 2 * F, 2 * F, [(pl, col)])                   # the concat itself (conv1's input)
        if ev is not None:
            ev[1].record()
            _ec_events.append((B, N, F, k, ev[0], ev[1]))
        ctx.save_for_backward(uv, idx, beta, mean, rstd, zmax, npos, both)
        ctx.dims = (B, N, F, k, dt)
        return both[:, :F], both[:, F:], both

    @staticmethod
    def backward(ctx, gmax, gmean, gboth):
        uv, idx, beta, mean, rstd, zmax, npos, both = ctx.saved_tensors
        B, N, F, k, dt = ctx.dims
        dev = uv.device
        L = nv.lib()
        st = nv.stream_ptr(dev)
        gmax = nv.require_cuda(gmax, "grad max") if gmax is not None else None
        gmean = nv.require_cuda(gmean, "grad mean") if gmean is not None else None
        gboth = nv.require_cuda(gboth, "grad concat") if gboth is not None else None
        ws = nv.workspace(dev, L.dgcnn_edgeconv_workspace_bytes(F), "stats")
        s1 = torch.empty(F, dtype=torch.float32, device=dev)
        s2 = torch.empty(F, dtype=torch.float32, device=dev)
        guv = torch.empty((uv.shape[0], 2 * F), dtype=torch.float32, device=dev)
        ev = None
        if _ec_bwd_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        nv.check(L.dgcnn_edgeconv_bwd_stats(both.data_ptr(), npos.data_ptr(), beta.data_ptr(), nv.ptr(gmax),
                                            nv.ptr(gmean), nv.ptr(gboth), B, N, F, k, s1.data_ptr(), s2.data_ptr(),
                                            guv.data_ptr(), ws.data_ptr(), ws.numel(), st), "edgeconv_bwd_stats")
        nv.check(L.dgcnn_edgeconv_bwd_apply(uv.data_ptr(), dt, idx.data_ptr(), B, N, F, k, mean.data_ptr(),
                                            rstd.data_ptr(), beta.data_ptr(), zmax.data_ptr(), nv.ptr(gmax), nv.ptr(gmean),
                                            nv.ptr(gboth), s1.data_ptr(), s2.data_ptr(), guv.data_ptr(), 1, st), "edgeconv_bwd_apply")
        if ev is not None:
            ev[1].record()
            _ec_bwd_events.append((B, N, F, k, ev[0], ev[1]))
        return guv, None, s1, None, None, None, None


class _BnAct(torch.autograd.Function):
    """slim.batch_norm (train mode, beta only, eps 1e-3) [+ residual] [+ ReLU] on a [P,C] tensor.
    group_bias [G,C] (optional) is added to the rows of group r // (P/G) before the statistics."""

    @staticmethod
    def forward(ctx, z, beta, residual, relu_flag, group_bias=None, sink_key=None):
        z = nv.require_cuda(z, "z")
        beta = nv.require_cuda(beta, "beta")
        res = nv.require_cuda(residual, "residual") if residual is not None else None
        gb = nv.require_cuda(group_bias, "group_bias") if group_bias is not None else None
        P, C = z.shape
        grows = P // gb.shape[0] if gb is not None else 0
        dev = z.device
        L = nv.lib()
        ws = nv.workspace(dev, L.dgcnn_bn_workspace_bytes(C), "stats")
        out = torch.empty_like(z)
        mean = torch.empty(C, dtype=torch.float32, device=dev)
        rstd = torch.empty(C, dtype=torch.float32, device=dev)
        sk = _sinks
        n_s, s_ptr, s_ld, s_pe, s_lst = (0, None, None, None, [])
        if sk is not None and sink_key is not None and C % 4 == 0:
            n_s, s_ptr, s_ld, s_pe, s_lst = sk.sink_args(sink_key)
            if any(pl.shape[1] != P for pl, _ in s_lst):
                n_s, s_lst = 0, []
        nv.check(L.dgcnn_bn_act_fwd_sinks(z.data_ptr(), P, C, beta.data_ptr(), nv.ptr(res), nv.ptr(gb), grows,
                                          int(bool(relu_flag)), out.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                          ws.data_ptr(), ws.numel(), n_s, s_ptr, s_ld, s_pe, nv.stream_ptr(dev)),
                 "bn_act_fwd")
        if n_s:
            sk.mark(out.data_ptr(), C, C, s_lst)
        ctx.save_for_backward(z, out, mean, rstd, gb)
        ctx.relu = bool(relu_flag)
        ctx.has_res = res is not None
        ctx.grows = grows
        return out

    @staticmethod
    def backward(ctx, g):
        z, out, mean, rstd, gb = ctx.saved_tensors
        g = nv.require_cuda(g, "grad")
        P, C = z.shape
        dev = z.device
        L = nv.lib()
        ws = nv.workspace(dev, L.dgcnn_bn_workspace_bytes(C), "stats")
        gz = torch.empty_like(z)
        gbeta = torch.empty(C, dtype=torch.float32, device=dev)
        gpre = torch.empty_like(z) if ctx.has_res else None
        nv.check(L.dgcnn_bn_act_bwd_gb(z.data_ptr(), out.data_ptr(), g.data_ptr(), P, C, mean.data_ptr(),
                                       rstd.data_ptr(), nv.ptr(gb), ctx.grows, int(ctx.relu), gz.data_ptr(),
                                       gbeta.data_ptr(), nv.ptr(gpre), ws.data_ptr(), ws.numel(), nv.stream_ptr(dev)),
                 "bn_act_bwd")
        ggb = gz.view(gb.shape[0], ctx.grows, C).sum(dim=1) if gb is not None else None
        return gz, gbeta, gpre, None, ggb, None


class _GroupMax(torch.autograd.Function):
    """max over the points of each cloud: x [G, rows, C] -> [G, C] (model.py:77)."""

    @staticmethod
    def forward(ctx, x):
        x = nv.require_cuda(x, "x")
        G, rows, C = x.shape
        out = torch.empty((G, C), dtype=torch.float32, device=x.device)
        cnt = torch.empty((G, C), dtype=torch.float32, device=x.device)
        nv.check(nv.lib().dgcnn_group_max_fwd(x.data_ptr(), G, rows, C, out.data_ptr(), cnt.data_ptr(),
                                              nv.stream_ptr(x.device)), "group_max_fwd")
        ctx.save_for_backward(x, out, cnt)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out, cnt = ctx.saved_tensors
        G, rows, C = x.shape
        g = nv.require_cuda(g, "grad")
        gx = torch.empty_like(x)
        nv.check(nv.lib().dgcnn_group_max_bwd(x.data_ptr(), out.data_ptr(), cnt.data_ptr(), g.data_ptr(), G, rows, C,
                                              gx.data_ptr(), nv.stream_ptr(x.device)), "group_max_bwd")
        return gx


class _PoolAndPass(torch.autograd.Function):
    """x [G, rows, C] -> (x itself, max over rows [G, C]) for a tensor that is BOTH max-pooled and consumed directly
    (model.py:76-85: the MergedEdgeConv output).  Backward adds the pooling gradient into the other consumer's
    gradient buffer at the arg-max positions only, instead of materialising a dense pooling gradient and summing."""

    @staticmethod
    def forward(ctx, x):
        x = nv.require_cuda(x, "x")
        G, rows, C = x.shape
        out = torch.empty((G, C), dtype=torch.float32, device=x.device)
        cnt = torch.empty((G, C), dtype=torch.float32, device=x.device)
        nv.check(nv.lib().dgcnn_group_max_fwd(x.data_ptr(), G, rows, C, out.data_ptr(), cnt.data_ptr(),
                                              nv.stream_ptr(x.device)), "group_max_fwd")
        ctx.save_for_backward(x, out, cnt)
        return x.view_as(x), out

    @staticmethod
    def backward(ctx, gx, gp):
        x, out, cnt = ctx.saved_tensors
        G, rows, C = x.shape
        if gx is None:
            gx = torch.zeros_like(x)
        else:
            gx = nv.require_cuda(gx, "grad")
        if gp is not None:
            gp = nv.require_cuda(gp, "grad pooled")
            nv.check(nv.lib().dgcnn_group_max_bwd_add(x.data_ptr(), out.data_ptr(), cnt.data_ptr(), gp.data_ptr(), G, rows,
                                                      C, gx.data_ptr(), nv.stream_ptr(x.device)), "group_max_bwd_add")
        return gx


def pool_and_pass(x: torch.Tensor):
    """-> (x, global max over dim 1) with the fused backward above (C % 4 == 0), else the separate ops."""
    if x.shape[-1] % 4 == 0:
        return _PoolAndPass.apply(x)
    return x, x.amax(dim=1)


class _SoftmaxXent(torch.autograd.Function):
    """trainval.py:39-52 fused: (logits [P,K], labels [P] int64, weights [P] or None) -> (mean weighted cross-entropy,
    accuracy); the gradient of the loss w.r.t. the logits is produced by the same kernel."""

    @staticmethod
    def forward(ctx, logits, labels, weights):
        logits = nv.require_cuda(logits, "logits")
        labels = nv.require_cuda(labels, "labels", torch.int64)
        weights = nv.require_cuda(weights, "weights") if weights is not None else None
        P, K = logits.shape
        dev = logits.device
        L = nv.lib()
        grad = torch.empty_like(logits)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        ws = torch.empty(L.dgcnn_softmax_xent_workspace_bytes(), dtype=torch.uint8, device=dev)
        nv.check(L.dgcnn_softmax_xent(logits.data_ptr(), labels.data_ptr(), nv.ptr(weights), P, K, grad.data_ptr(),
                                      out.data_ptr(), ws.data_ptr(), ws.numel(), nv.stream_ptr(dev)), "softmax_xent")
        ctx.save_for_backward(grad)
        loss, acc = out[0].clone(), out[1].clone()
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, gloss, gacc):
        (grad,) = ctx.saved_tensors
        return grad * gloss, None, None


def softmax_xent(logits2d, labels1d, weights1d=None):
    """-> (loss, accuracy) 0-d tensors; differentiable w.r.t. logits2d."""
    return _SoftmaxXent.apply(logits2d, labels1d, weights1d)


def global_max_pool(x: torch.Tensor) -> torch.Tensor:
    """x [B,N,C] -> [B,C]: hand-written kernels when C % 4 == 0, else torch.amax (same tie semantics)."""
    if x.shape[-1] % 4 == 0:
        return _GroupMax.apply(x)
    return x.amax(dim=1)


# =============================================================================== variables (slim.conv2d)
def _conv_bn_vars(scope: str, cin: int, cout: int, trainable: bool, device):
    """Variables slim.conv2d(normalizer_fn=slim.batch_norm) creates under `scope`: weights (xavier, no bias)
    and BatchNorm/beta (zeros; beta is always trainable, scale=False => no gamma)  [TF-default]."""
    st = default_store()
    with st.variable_scope(scope):
        w = st.get_variable("weights", (cin, cout), "xavier", trainable=trainable, device=device)
        with st.variable_scope("BatchNorm"):
            b = st.get_variable("beta", (cout,), "zeros", trainable=True, device=device)
    return w, b


def _conv_bn_act(net2d, scope, cout, trainable, activation, residual2d=None, sink_key=None):
    """1x1 conv + BN(train) [+ residual] + activation on a [P,Cin] tensor, all through the C ABI."""
    w, b = _conv_bn_vars(scope, net2d.shape[1], cout, trainable, net2d.device)
    z = _Conv1x1.apply(net2d, w)
    return _BnAct.apply(z, b, residual2d, activation is not None, None, sink_key)


def _dbg(debug, t, name):
    if debug:
        print("Shape %s ... Name %s" % (tuple(t.shape), name))


def _cur_scope():
    return "/".join(default_store()._scope)


# =============================================================================== reference API
def edge_conv(point_cloud, k, num_filters, trainable, activation=relu, debug=False, _residual=None,
              _knn_hint=None, _knn_out=None) -> List[torch.Tensor]:
    """ops.py:42-73 -> [net_max [B,N,1,F], net_mean [B,N,1,F], net [B,N,1,64]].

    edges() -> conv0 is evaluated as  [x_i, x_j-x_i].[Wa;Wb] = x_i.(Wa-Wb) + x_j.Wb : one per-point GEMM
    (uv) followed by two L2-resident gather passes; nothing of size B*N*k*{2C,F} touches HBM.
    `_residual` (private) fuses ops.py:134's relu(shortcut + net) into conv1's BN epilogue.
    """
    x = nv.require_cuda(point_cloud, "point_cloud")
    if x.dim() != 3:
        raise ValueError("edge_conv: point_cloud must be [B,N,C]")
    B, N, C = x.shape
    k = int(k)
    F = int(num_filters)
    idx = _layer_knn(x, k, _knn_hint)                                           # ops.py:23
    if _knn_out is not None:
        _knn_out.append(idx)
    if debug: _dbg(debug, torch.empty(B, N, k, 2 * C, device="meta"), _cur_scope() + "/edges (never materialised)")
    w0, b0 = _conv_bn_vars("conv0", 2 * C, F, trainable, x.device)              # ops.py:47-54
    wp = _EdgeWeights.apply(w0)                                                 # [C, 2F] = [Wa-Wb | Wb]
    # f32 mode: the uv GEMM stays on the exact fp32 SIMT kernel.  z_ij = u_i + v_j carries its information in the small
    # differences between neighbours' v rows (the reference subtracts x_j - x_i BEFORE the conv), so an operand error of
    # 2^-16 (bf16 hi/lo planes) would be amplified by the cancellation.
    uv = _Conv1x1.apply(x.reshape(B * N, C), wp, (not _bf16_uv_gemm) if _precision == "bf16" else True)
    li = _sinks.layer if _sinks is not None else -1
    net_max, net_mean, net = _EdgeConvGather.apply(uv, idx, b0, B, N, k, ("ec", li, "both"))   # ops.py:53-58
    if debug: _dbg(debug, torch.empty(B, N, k, F, device="meta"), _cur_scope() + "/conv0 (never materialised)")
    _dbg(debug, net_max.view(B, N, 1, F), _cur_scope() + "/Max")
    _dbg(debug, net_mean.view(B, N, 1, F), _cur_scope() + "/Mean")
    _dbg(debug, net.view(B, N, 1, 2 * F), _cur_scope() + "/concat")
    res2d = _residual.reshape(B * N, CONV1_WIDTH) if _residual is not None else None
    act = relu if (_residual is not None) else activation
    net = _conv_bn_act(net, "conv1", CONV1_WIDTH, trainable, act, res2d, ("ec", li, "net"))   # ops.py:62-70 (+134)
    _dbg(debug, net.view(B, N, 1, CONV1_WIDTH), _cur_scope() + "/conv1")
    return [net_max.view(B, N, 1, F), net_mean.view(B, N, 1, F), net.view(B, N, 1, CONV1_WIDTH)]


def _listify(v, repeat, what):
    if not type(v) == type(list()):
        return [int(v)] * repeat
    if not len(v) == repeat:
        print("Length of %s != repeat" % what)
        raise ValueError
    return v


def repeat_edge_conv(point_cloud, repeat, k, num_filters, trainable, debug=False):
    """ops.py:75-98."""
    repeat = int(repeat)
    k = _listify(k, repeat, "k")
    num_filters = _listify(num_filters, repeat, "num_filters")
    net = point_cloud
    tensors = []
    graph = []  # previous layer's neighbour lists warm-start the next layer's selection (same result)
    for i in range(repeat):
        if _sinks is not None:
            _sinks.layer = i
        with variable_scope("EdgeConv%d" % i):
            tensors += edge_conv(net, k[i], num_filters[i], trainable, debug=debug,
                                 _knn_hint=graph[-1] if graph else None, _knn_out=graph)
            net = tensors[-1].squeeze(-2)
    return tensors


def repeat_residual_edge_conv(point_cloud, repeat, k, num_filters, trainable, debug=False):
    """ops.py:100-140."""
    repeat = int(repeat)
    k = _listify(k, repeat, "k")
    num_filters = _listify(num_filters, repeat, "num_filters")
    net = point_cloud
    tensors = []
    shortcut = None
    graph = []
    for i in range(repeat):
        if _sinks is not None:
            _sinks.layer = i
        with variable_scope("EdgeConv%d" % i):
            if shortcut is None:
                tensors += edge_conv(net, k[i], num_filters[i], trainable, debug=debug, _knn_out=graph)
            else:
                if not num_filters[i] == num_filters[i - 1]:            # ops.py:124-133
                    B, N = shortcut.shape[0], shortcut.shape[1]
                    if int(num_filters[i]) != CONV1_WIDTH:
                        # the reference would fail in tf.add here: conv1 is 64 wide whatever num_filters says
                        raise ValueError("shortcut conv to %d channels cannot be added to the 64-channel conv1 "
                                         "output (ops.py:63,134)" % int(num_filters[i]))
                    sc = _conv_bn_act(shortcut.reshape(B * N, -1), "shortcut", int(num_filters[i]), trainable, None)
                    shortcut = sc.view(B, N, 1, -1)
                # conv1 with activation=None, then relu(shortcut + out)  (ops.py:121,134) fused in one epilogue
                tensors += edge_conv(net, k[i], num_filters[i], trainable, activation=None, debug=debug,
                                     _residual=shortcut, _knn_hint=graph[-1], _knn_out=graph)
            net = tensors[-1]
            shortcut = tensors[-1]
            net = net.squeeze(-2)
    return tensors


def fc(net, repeat, num_filters, trainable, debug=False):
    """ops.py:142-163: repeat x (1x1 conv + BN + ReLU) on [B,N,1,C]."""
    repeat = int(repeat)
    num_filters = _listify(num_filters, repeat, "num_filters")
    from .head import conv_bn_relu_dense
    for i in range(repeat):
        net = conv_bn_relu_dense(net, "FC%d" % i, int(num_filters[i]), trainable)
        _dbg(debug, net, _cur_scope() + "/FC%d" % i)
    return net
