"""Dense per-point layers of the model head (MergedEdgeConv / FC* / Final): SURVEY.md section 8(f) row N1.

/root/reference/dgcnn/model.py:60-104 and ops.py:151-160 -- tf.concat + slim.conv2d(kernel 1) + slim.batch_norm +
ReLU on [B,N,1,C].  Here:
  * the 1x1 convs run on the hand-written tcgen05 GEMM (fp32-faithful bf16 hi/lo split, csrc/tc_gemm.cu);
  * channel concatenations are never materialised: every source tensor is split straight into its column slice
    of the GEMM operand (ops._ConcatConvTC);
  * the global max-pooled feature (model.py:76-81 tiles it over all N points before the concat) enters FC0 as a
    per-cloud term  g_b . W[:1024]  added to the rows of cloud b -- algebraically the same product, 36 % fewer
    FLOPs and no [B,N,1024] broadcast;
  * BatchNorm(train)+ReLU are the hot path's own kernels (ops._BnAct), so BN semantics are identical everywhere.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops as _ops


def conv_bn_act(srcs: List[torch.Tensor], scope: str, cout: int, trainable: bool, activation=_ops.relu,
                cloud_feature: Optional[torch.Tensor] = None, points_per_cloud: int = 0, pool_rows: int = 0):
    """srcs: [P, c_i] tensors whose channel concat is the layer input (after the optional per-cloud feature
    [B, cg] that the reference tiles in front of it).  -> [P, cout]; with pool_rows > 0 -> ([P, cout], [P/pool_rows, cout]):
    the layer output and its maximum over every pool_rows consecutive rows (model.py:76-77, fused when possible)."""
    cg = cloud_feature.shape[1] if cloud_feature is not None else 0
    cin = cg + sum(int(t.shape[1]) for t in srcs)
    w, b = _ops._conv_bn_vars(scope, cin, cout, trainable, srcs[0].device)
    w_g, w_rest = _ops._SplitWeightRows.apply(w, cg) if cg else (None, w)
    gb = _ops._Conv1x1.apply(cloud_feature, w_g, True) if cg else None   # [B, cout]: tiny, exact SIMT; added inside BN
    P = srcs[0].shape[0]
    if _ops._tc_ok(P, cout, *[t.shape[1] for t in srcs]):
        # the whole layer around one tcgen05 GEMM: BN statistics from its epilogue, gradients as bf16 planes
        kin = sum(int(t.shape[1]) for t in srcs)
        fuse_pool = bool(pool_rows) and gb is None and activation is not None and P % pool_rows == 0 and \
            _ops.nv.lib().dgcnn_tc_gemm_stats_supported(P, cout, kin) and _ops._FUSE_POOL
        out, pooled = _ops._ConvBnActTC.apply(w_rest, b, gb, activation is not None, points_per_cloud, scope,
                                              pool_rows if fuse_pool else 0, *srcs)
        if not pool_rows:
            return out
        if fuse_pool:
            return out, pooled
    else:
        z = _ops.conv1x1(srcs, w_rest)
        out = _ops._BnAct.apply(z, b, None, activation is not None, gb, None)
        if not pool_rows:
            return out
    out3, pooled = _ops.pool_and_pass(out.view(P // pool_rows, pool_rows, cout))      # separate pooling kernels
    return out3.view(P, cout), pooled


def conv_bn_relu_dense(net: torch.Tensor, scope: str, cout: int, trainable: bool, activation=_ops.relu) -> torch.Tensor:
    """net [B,N,1,Cin] -> [B,N,1,cout] (the reference's tensor shapes)."""
    B, N, one, cin = net.shape
    y = conv_bn_act([net.reshape(B * N * one, cin)], scope, cout, trainable, activation)
    return y.view(B, N, one, cout)
