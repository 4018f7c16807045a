"""Dense per-point layers of the model head (MergedEdgeConv / FC* / Final): SURVEY.md section 8(f) row N1.

/root/reference/dgcnn/model.py:65-72,94-101 and ops.py:151-160 -- slim.conv2d(kernel 1) + slim.batch_norm +
ReLU on [B,N,1,C].  The head is outside the EdgeConv hot path; its GEMM is a plain library GEMM
(torch.matmul -> cuBLAS, true fp32: TF32 stays disabled), while BatchNorm+ReLU reuse the hand-written
train-mode BN kernels of the hot path so that BN semantics (batch statistics, biased variance, eps=1e-3,
beta only) are identical everywhere.
"""
from __future__ import annotations

import torch

from . import ops as _ops


def conv_bn_relu_dense(net: torch.Tensor, scope: str, cout: int, trainable: bool, activation=_ops.relu) -> torch.Tensor:
    """net [B,N,1,Cin] -> [B,N,1,cout]."""
    B, N, one, cin = net.shape
    w, b = _ops._conv_bn_vars(scope, cin, cout, trainable, net.device)
    z = torch.matmul(net.reshape(B * N * one, cin), w)
    y = _ops._BnAct.apply(z, b, None, activation is not None)
    return y.view(B, N, one, cout)
