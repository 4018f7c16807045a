"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/dgcnn_b200.h declares.
No compute is launched here: only argument validation paths that return before touching CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dgcnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgcnn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(dg):
    from dgcnn import _native
    lib = _native.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_native.SIGNATURES) == names  # the ctypes table covers the whole header, nothing else
    assert lib.dgcnn_abi_version() == 2


def test_workspace_queries_are_pure(dg):
    from dgcnn import _native
    lib = _native.lib()
    # xT [B,C,Npad] + s [B,Npad], Npad = N rounded up to 128
    base = (24 * 64 * 2048 + 24 * 2048) * 4
    got = lib.dgcnn_knn_workspace_bytes(24, 2048, 64)   # + tensor-core filter scratch: norms, fp16 operand, norm
    # k-slices, candidate lists (4 x 48 x u16 per row), counts, fallback queue  (knn_tc.cu knn_tc_bytes)
    per_row = 8 + 2 * 64 * 2 + 64 + 4 * 48 * 2 + 4 + 16
    assert base + 24 * 2048 * per_row <= got <= base + 24 * 2048 * (per_row + 8)   # + per-cloud range partials
    assert lib.dgcnn_knn_workspace_bytes(24, 2048, 64) == got                 # pure function of the shape
    assert lib.dgcnn_knn_workspace_bytes(2, 100, 3) == (2 * 3 * 128 + 2 * 128) * 4
    assert lib.dgcnn_knn_workspace_bytes(0, 5, 5) == 0
    assert lib.dgcnn_gemm_workspace_bytes(49152, 128, 64, 0, 0) == 0          # enough tiles: no split-K
    assert lib.dgcnn_gemm_workspace_bytes(64, 128, 49152, 1, 0) > 0           # weight gradient: split over points
    assert lib.dgcnn_edgeconv_workspace_bytes(64) > 0 and lib.dgcnn_bn_workspace_bytes(64) > 0
    # the single query of SURVEY.md 8b dispatches to the per-op ones (op, B, N, C, k, F)
    OP_KNN, OP_EDGECONV, OP_CONV_FWD, OP_CONV_DW, OP_BN, OP_XENT = range(6)
    assert lib.dgcnn_workspace_bytes(OP_KNN, 24, 2048, 64, 20, 64) == got
    assert lib.dgcnn_workspace_bytes(OP_EDGECONV, 24, 2048, 64, 20, 64) == lib.dgcnn_edgeconv_workspace_bytes(64)
    assert lib.dgcnn_workspace_bytes(OP_CONV_FWD, 24, 2048, 64, 20, 128) == lib.dgcnn_gemm_workspace_bytes(49152, 128, 64, 0, 0)
    assert lib.dgcnn_workspace_bytes(OP_CONV_DW, 24, 2048, 64, 20, 128) == lib.dgcnn_gemm_workspace_bytes(64, 128, 49152, 1, 0)
    assert lib.dgcnn_workspace_bytes(OP_BN, 24, 2048, 64, 20, 512) == lib.dgcnn_bn_workspace_bytes(512)
    assert lib.dgcnn_workspace_bytes(OP_XENT, 24, 2048, 64, 20, 2) == lib.dgcnn_softmax_xent_workspace_bytes()
    assert lib.dgcnn_workspace_bytes(99, 24, 2048, 64, 20, 64) == 0 and lib.dgcnn_workspace_bytes(OP_KNN, 0, 8, 3, 2, 4) == 0


def test_invalid_arguments_return_codes_not_aborts(dg):
    from dgcnn import _native
    lib = _native.lib()
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.addressof(buf)
    p = (p + 15) & ~15
    # null pointers
    assert lib.dgcnn_knn(None, None, 1, 8, 3, 2, None, 0, None) == _native.ERR_INVALID
    assert "null" in _native.last_error()
    # k > N
    assert lib.dgcnn_knn(p, p, 1, 8, 3, 9, p, 1 << 20, None) == _native.ERR_INVALID
    assert "k" in _native.last_error()
    # k above the compiled envelope
    assert lib.dgcnn_knn(p, p, 1, 200, 3, 65, p, 1 << 20, None) == _native.ERR_UNSUPPORTED
    # workspace too small
    assert lib.dgcnn_knn(p, p, 1, 8, 3, 2, p, 16, None) == _native.ERR_WORKSPACE
    assert lib.dgcnn_topk_rows(p, p, 4, 8, 0, None) == _native.ERR_INVALID
    assert lib.dgcnn_gemm(p, p, p, 0, 4, 4, 0, 0, None, 0, None) == _native.ERR_INVALID
    assert lib.dgcnn_edgeconv_fwd_stats(p, 0, p, 1, 8, 64, 9, p, p, p, 1 << 20, None) == _native.ERR_INVALID      # k > N
    assert lib.dgcnn_edgeconv_fwd_stats(p, 7, p, 1, 8, 64, 4, p, p, p, 1 << 20, None) == _native.ERR_INVALID      # dtype
    assert lib.dgcnn_edgeconv_fwd_stats(p, 0, p, 1, 8, 64, 4, p, p, p, 8, None) == _native.ERR_WORKSPACE
    assert lib.dgcnn_bn_act_fwd(p, 0, 4, p, None, 1, p, p, p, p, 1 << 20, None) == _native.ERR_INVALID
    assert lib.dgcnn_adam_tf_step(p, p, p, p, 0, 0.1, 0.9, 0.999, 1e-8, 1.0, None) == _native.ERR_INVALID
    with pytest.raises(ValueError):
        _native.check(_native.ERR_INVALID, "x")
    with pytest.raises(NotImplementedError):
        _native.check(_native.ERR_UNSUPPORTED, "x")


def test_product_path_refuses_cpu_tensors(dg):
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dg.ops.k_nn(torch.zeros(1, 8, 3), 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dg.ops.pairwise_distance(torch.zeros(1, 8, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dynamic-gcnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
