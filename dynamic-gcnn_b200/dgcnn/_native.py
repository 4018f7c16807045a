"""ctypes binding of libdgcnn_b200.so (include/dgcnn_b200.h) for torch CUDA tensors.

This is the only place where Python touches the native library.  There is NO CPU fallback: if the
shared library is missing, or a tensor is not a CUDA tensor, the call raises.  PyTorch is used for
device memory (caching allocator), the current stream and autograd bookkeeping only.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGCNN_LIB_PATH") or os.path.normpath(os.path.join(_HERE, "..", "lib", "libdgcnn_b200.so"))

_lib = None

_vp, _i, _i64, _sz, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes); mirrors include/dgcnn_b200.h one to one
SIGNATURES = {
    "dgcnn_abi_version": (_i, []),
    "dgcnn_last_error": (ctypes.c_char_p, []),
    "dgcnn_launch_count": (ctypes.c_uint64, []),
    "dgcnn_knn_workspace_bytes": (_sz, [_i, _i, _i]),
    "dgcnn_pairwise_distance": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "dgcnn_knn": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dgcnn_knn_hinted": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dgcnn_knn_mode": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dgcnn_topk_rows": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "dgcnn_edge_feature": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dgcnn_edge_feature_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dgcnn_gemm_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dgcnn_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dgcnn_split_bf16": (_i, [_vp, _i64, _i, _i64, _vp, _i64, _i64, _vp]),
    "dgcnn_tc_gemm_workspace_bytes": (_sz, [_i, _i, _i]),
    "dgcnn_tc_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dgcnn_tc_gemm_a_slice": (_i, [_vp, _i64, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dgcnn_tc_gemm_stats_supported": (_i, [_i, _i, _i]),
    "dgcnn_tc_gemm_stats": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "dgcnn_tc_gemm_grouped": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dgcnn_edgeconv_workspace_bytes": (_sz, [_i]),
    "dgcnn_edgeconv_fwd_stats": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dgcnn_edgeconv_fwd_apply": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "dgcnn_edgeconv_bwd_stats": (_i, [_vp] * 6 + [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dgcnn_edgeconv_bwd_apply": (_i, [_vp, _i, _vp, _i, _i, _i, _i] + [_vp] * 10 + [_i, _vp]),
    "dgcnn_bn_act_fwd_sinks": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp, _vp, _vp, _vp]),
    "dgcnn_bn_apply_fwd_sinks": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "dgcnn_bn_workspace_bytes": (_sz, [_i]),
    "dgcnn_bn_act_fwd": (_i, [_vp, _i64, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dgcnn_bn_act_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dgcnn_bn_act_fwd_gb": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dgcnn_bn_act_bwd_gb": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dgcnn_bn_stats_from_tiles": (_i, [_vp, _i, _i, _i64, _vp, _i, _vp, _vp, _vp]),
    "dgcnn_bn_apply_fwd": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "dgcnn_bn_act_bwd_planes": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "dgcnn_bn_pool_workspace_bytes": (_sz, [_i, _i]),
    "dgcnn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "dgcnn_bn_apply_fwd_pool": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _i, _vp, _vp, _vp, _vp]),
    "dgcnn_group_max_fwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "dgcnn_group_max_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "dgcnn_group_max_bwd_add": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "dgcnn_softmax_xent_workspace_bytes": (_sz, []),
    "dgcnn_softmax_xent": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _sz, _vp]),
    "dgcnn_adam_tf_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _vp]),
}

ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = -1, -2, -3, -4
DT_F32, DT_BF16, DT_F16 = 0, 1, 2   # DGCNN_F32 / DGCNN_BF16 / DGCNN_F16


def lib():
    """Load the C-ABI library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "dgcnn: native library %s is missing (build it with `make -C dynamic-gcnn_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU fallback." % LIB_PATH
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().dgcnn_last_error().decode("utf-8", "replace")


_replayed_launches = 0


def add_replayed_launches(n: int) -> None:
    """Kernels of this library launched by replaying a captured CUDA graph (the library's own counter only sees
    the capture)."""
    global _replayed_launches
    _replayed_launches += int(n)


def launch_count() -> int:
    return int(lib().dgcnn_launch_count()) + _replayed_launches


def check(rc: int, what: str) -> None:
    """Error codes -> exceptions: bad shapes are ValueError like the reference's ops.py:80-87."""
    if rc == 0:
        return
    msg = "%s: %s" % (what, last_error())
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg + " (code %d)" % rc)


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("dgcnn: %s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("dgcnn: %s must live on a CUDA device (no CPU fallback exists)" % name)
    if t.dtype != dtype:
        raise TypeError("dgcnn: %s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def require_cuda_rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """fp32 CUDA [rows, cols] whose rows may be strided (a column slice of a wider buffer): no copy if the elements of
    a row are contiguous and rows / base are 16-byte aligned."""
    if not isinstance(t, torch.Tensor):
        raise TypeError("dgcnn: %s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("dgcnn: %s must live on a CUDA device (no CPU fallback exists)" % name)
    if t.dtype != torch.float32:
        raise TypeError("dgcnn: %s must be float32, got %s" % (name, t.dtype))
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1] and t.stride(0) % 4 == 0 and \
            t.data_ptr() % 16 == 0:
        return t
    return t.contiguous()


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


_ws_cache: Dict[Tuple[int, str], torch.Tensor] = {}


def workspace(dev: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    """Caller-owned scratch (the library never allocates): a grow-only uint8 buffer per (device, tag).
    All calls are ordered on the current stream, so reuse between consecutive ops is safe."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    return buf


def workspaces_snapshot():
    """References to every scratch buffer currently in the cache (see trainval._tower_step: a captured CUDA graph must
    keep the buffers it was recorded with alive)."""
    return list(_ws_cache.values())


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
