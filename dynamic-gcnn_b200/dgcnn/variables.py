"""Variable store with TF1-style scopes.

The reference creates its weights implicitly: `slim.conv2d(..., scope='conv0')` inside
`tf.variable_scope('EdgeConv%d')` inside `tf.variable_scope("dgcnn", reuse=tf.AUTO_REUSE)`
(/root/reference/dgcnn/ops.py:47-54,92 ; trainval.py:29).  To keep the signatures of dgcnn.ops /
dgcnn.model.build unchanged, this module provides the same create-or-reuse semantics on top of
torch tensors: names equal the TF variable names (`dgcnn/EdgeConv0/conv0/weights`,
`.../BatchNorm/beta`), so checkpoints are keyed like the reference's Saver files.

Weights are stored 2-D [Cin, Cout] (the TF [1,1,Cin,Cout] HWIO kernel with the unit axes dropped).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from contextlib import contextmanager
from typing import Dict, Iterable, List, Optional

import torch


class VariableStore:
    def __init__(self, device=None, seed: int = 0):
        self.device = torch.device(device) if device is not None else None
        self.vars: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.trainable: Dict[str, bool] = {}
        self._scope: List[str] = []
        self._gen = torch.Generator().manual_seed(int(seed) & 0x7FFFFFFF)
        self.flat_param: Optional[torch.Tensor] = None
        self.flat_grad: Optional[torch.Tensor] = None

    # -- scopes ------------------------------------------------------------------------------
    @contextmanager
    def variable_scope(self, name: str):
        self._scope.append(name)
        try:
            yield
        finally:
            self._scope.pop()

    def full_name(self, name: str) -> str:
        return "/".join(self._scope + [name])

    # -- create-or-reuse (tf.AUTO_REUSE) ---------------------------------------------------------
    def get_variable(self, name: str, shape, init: str, trainable: bool = True, device=None) -> torch.Tensor:
        full = self.full_name(name)
        v = self.vars.get(full)
        if v is not None:
            if tuple(v.shape) != tuple(shape):
                raise ValueError("variable %s exists with shape %s, requested %s" % (full, tuple(v.shape), tuple(shape)))
            return v
        dev = self.device or device
        if dev is None:
            raise RuntimeError("VariableStore has no device yet")
        if self.device is None:
            self.device = torch.device(dev)
        if init == "xavier":  # tf.contrib.layers.xavier_initializer(): uniform +-sqrt(6/(fan_in+fan_out))
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(tuple(shape), generator=self._gen, dtype=torch.float64) * 2 - 1) * lim
        elif init == "zeros":
            t = torch.zeros(tuple(shape), dtype=torch.float64)
        else:
            raise ValueError(init)
        v = t.to(torch.float32).to(self.device)
        v.requires_grad_(bool(trainable))
        self.vars[full] = v
        self.trainable[full] = bool(trainable)
        if self.flat_param is not None:
            raise RuntimeError("variable %s created after the store was flattened" % full)
        return v

    # -- flat views: one buffer for Adam and for the single gradient all-reduce ---------------------
    def trainable_names(self) -> List[str]:
        return [n for n in self.vars if self.trainable[n]]

    def flatten(self, extra: int = 0) -> None:
        """Re-home every trainable variable (and its .grad) as a view into one flat fp32 buffer.
        `extra` trailing floats ride along in the gradient buffer (loss/accuracy for logging)."""
        names = self.trainable_names()
        total = sum(self.vars[n].numel() for n in names)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(total + extra, dtype=torch.float32, device=self.device)
        off = 0
        for n in names:
            v = self.vars[n]
            k = v.numel()
            self.flat_param[off:off + k].copy_(v.detach().reshape(-1))
            v.data = self.flat_param[off:off + k].view(v.shape)
            v.grad = self.flat_grad[off:off + k].view(v.shape)
            off += k
        self.num_trainable = total

    def num_params(self) -> int:
        return sum(self.vars[n].numel() for n in self.trainable_names())

    # -- checkpoint I/O (names = TF variable names) ------------------------------------------------
    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((n, v.detach().cpu().clone()) for n, v in self.vars.items())

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        for n, v in self.vars.items():
            if n not in sd:
                if strict:
                    raise KeyError("checkpoint lacks variable %s" % n)
                continue
            src = sd[n]
            if src.dim() == 4 and v.dim() == 2:  # TF [1,1,Cin,Cout] kernels
                src = src.reshape(src.shape[2], src.shape[3])
            with torch.no_grad():
                v.copy_(src.to(v.device, v.dtype))


_default: Optional[VariableStore] = None


def default_store() -> VariableStore:
    global _default
    if _default is None:
        _default = VariableStore()
    return _default


def set_default_store(store: Optional[VariableStore]) -> Optional[VariableStore]:
    global _default
    old, _default = _default, store
    return old


def reset_default_store(device=None, seed: int = 0) -> VariableStore:
    """tf.reset_default_graph() analogue."""
    global _default
    _default = VariableStore(device=device, seed=seed)
    return _default


@contextmanager
def variable_scope(name: str):
    with default_store().variable_scope(name):
        yield
