"""tcgen05 GEMM (dgcnn_tc_gemm, bf16 hi/lo split, fp32 accumulate) against an fp64 matmul, all three operand
layouts (forward, dX, dW) and ragged tile edges.  Tolerance: 2^-17-class relative error per product."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(dg, x):
    from dgcnn import _native as nv
    x = x.contiguous()
    planes = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    x2 = x.reshape(-1, x.shape[-1])
    nv.check(nv.lib().dgcnn_split_bf16(x2.data_ptr(), x2.shape[0], x2.shape[1], x2.shape[1], planes.data_ptr(),
                                       x2.shape[1], x.numel(), nv.stream_ptr(x.device)), "split")
    return planes


def _tc(dg, A, B, M, N, K, tA, tB):
    from dgcnn import _native as nv
    L = nv.lib()
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    need = L.dgcnn_tc_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=A.device)
    pa, pb = _split(dg, A), _split(dg, B)
    nv.check(L.dgcnn_tc_gemm(pa.data_ptr(), pb.data_ptr(), out.data_ptr(), M, N, K, tA, tB, 2, ws.data_ptr(), ws.numel(),
                             nv.stream_ptr(A.device)), "tc_gemm")
    return out


def test_split_planes_reconstruct(dg, cuda):
    x = torch.randn(64, 64, device=cuda) * 3
    p = _split(dg, x)
    rec = p[0].float() + p[1].float()
    assert (rec - x).abs().max() <= x.abs().max() * 2.0 ** -16
    assert torch.equal(p[0], x.to(torch.bfloat16))


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 256), (1000, 264, 520), (4096, 512, 1792), (136, 8, 72)])
@pytest.mark.parametrize("mode", ["fwd", "dx", "dw"])
def test_tc_gemm_matches_fp64(dg, cuda, M, N, K, mode):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    if mode == "fwd":      # C[M,N] = A[M,K] . B[K,N]
        A = torch.randn((M, K), generator=g).to(cuda)
        B = torch.randn((K, N), generator=g).to(cuda)
        ref = A.double() @ B.double()
        out = _tc(dg, A, B, M, N, K, 0, 0)
    elif mode == "dx":     # C[M,N] = A[M,K] . B[N,K]^T
        A = torch.randn((M, K), generator=g).to(cuda)
        B = torch.randn((N, K), generator=g).to(cuda)
        ref = A.double() @ B.double().t()
        out = _tc(dg, A, B, M, N, K, 0, 1)
    else:                  # C[M,N] = A[K,M]^T . B[K,N]
        A = torch.randn((K, M), generator=g).to(cuda)
        B = torch.randn((K, N), generator=g).to(cuda)
        ref = A.double().t() @ B.double()
        out = _tc(dg, A, B, M, N, K, 1, 0)
    err = (out.double() - ref).abs().max().item()
    scale = (A.double().abs().max() * B.double().abs().max()).item() * np.sqrt(K)
    assert err <= 4e-5 * scale, (err, scale)
    fp32 = (A @ B if mode == "fwd" else (A @ B.t() if mode == "dx" else A.t() @ B)).double()
    # not worse than 8x the error of a plain fp32 GEMM on the same data (+ floor)
    assert err <= 8 * (fp32 - ref).abs().max().item() + 1e-5 * scale


def test_tc_gemm_weight_gradient_split_k(dg, cuda):
    """dW shape of the head: small M,N and the 49152 points as contraction dimension (split over the grid)."""
    g = torch.Generator(device="cpu").manual_seed(0)
    P, Cin, Cout = 24 * 2048, 256, 512
    X = torch.randn((P, Cin), generator=g).to(cuda)
    G = torch.randn((P, Cout), generator=g).to(cuda) * 0.01
    out = _tc(dg, X, G, Cin, Cout, P, 1, 0)
    ref = X.double().t() @ G.double()
    assert (out.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_tc_gemm_rejects_bad_shapes(dg, cuda):
    from dgcnn import _native as nv
    x = torch.zeros(64, device=cuda)
    assert nv.lib().dgcnn_tc_gemm(x.data_ptr(), x.data_ptr(), x.data_ptr(), 12, 8, 8, 0, 0, 2, None, 0, None) == nv.ERR_UNSUPPORTED
    assert nv.lib().dgcnn_tc_gemm(None, x.data_ptr(), x.data_ptr(), 8, 8, 8, 0, 0, 2, None, 0, None) == nv.ERR_INVALID


@pytest.mark.parametrize("M,N,K", [(130, 70, 33), (256, 128, 64), (49, 8, 1030), (64, 128, 4096), (1000, 64, 128)])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_simt_gemm_exact_fp32(dg, cuda, M, N, K, tA, tB):
    """dgcnn_gemm (exact fp32 SIMT, vectorised and scalar paths, split-K) against fp64."""
    from dgcnn import ops
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(cuda)
    B = torch.randn((N, K) if tB else (K, N), generator=g).to(cuda)
    out = ops._gemm_raw(A, B, M, N, K, tA, tB)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    assert (out.double() - ref).abs().max().item() <= 2e-6 * np.sqrt(K) * max(1.0, ref.abs().max().item())


def test_concat_conv_grouped_gradients(dg, cuda):
    """_ConcatConvTC: forward == conv of the concatenation; every source gets its own dense gradient."""
    from dgcnn import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    P = 2048
    srcs = [torch.randn((P, c), generator=g).to(cuda).requires_grad_(True) for c in (64, 64, 128, 1024, 32)]
    w = (torch.randn((sum(t.shape[1] for t in srcs), 256), generator=g) * 0.05).to(cuda).requires_grad_(True)
    out = ops.conv1x1(srcs, w)
    cat = torch.cat([t.detach() for t in srcs], 1).double().requires_grad_(True)
    wd = w.detach().double().requires_grad_(True)
    ref = cat @ wd
    assert (out.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    go = torch.randn((P, 256), generator=g).to(cuda)
    out.backward(go)
    ref.backward(go.double())
    off = 0
    for t in srcs:
        c = t.shape[1]
        assert t.grad.is_contiguous()
        assert (t.grad.double() - cat.grad[:, off:off + c]).abs().max().item() <= 1e-4 * cat.grad.abs().max().item()
        off += c
    assert (w.grad.double() - wd.grad).abs().max().item() <= 1e-4 * wd.grad.abs().max().item()


def test_global_max_pool_matches_amax(dg, cuda):
    from dgcnn import ops
    x = torch.randn(3, 257, 64, device=cuda)
    x[:, 5] = x[:, 9]                      # exact ties across points
    x[0, :, 3] = 0.0                       # a whole channel tied
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    ya, yb = ops.global_max_pool(a), b.amax(dim=1)
    assert torch.equal(ya, yb)
    w = torch.randn_like(ya)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    assert torch.allclose(a.grad, b.grad, atol=1e-6)


@pytest.mark.parametrize("M,N,K,tA,tB", [(1024, 128, 64, 0, 0), (4096, 512, 320, 0, 0), (2048, 64, 128, 0, 1),
                                          (128, 64, 8192, 1, 0), (2816, 512, 4096, 1, 0)])
def test_tc_gemm_single_plane_bf16(dg, cuda, M, N, K, tA, tB):
    """planes = 1 (BASELINE.json configs[2]'s bf16 arithmetic): one MMA per product on the hi plane only.  Against fp64 on
    the bf16-ROUNDED operands the result is exact up to fp32 accumulation."""
    from dgcnn import _native as nv
    L = nv.lib()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(cuda)
    B = torch.randn((N, K) if tB else (K, N), generator=g).to(cuda)
    pa = torch.empty((1,) + tuple(A.shape), dtype=torch.bfloat16, device=cuda)
    pb = torch.empty((1,) + tuple(B.shape), dtype=torch.bfloat16, device=cuda)
    for x, p in ((A, pa), (B, pb)):
        nv.check(L.dgcnn_split_bf16(x.data_ptr(), x.shape[0], x.shape[1], x.shape[1], p.data_ptr(), x.shape[1], 0,
                                    nv.stream_ptr(cuda)), "split")
        assert torch.equal(p[0], x.to(torch.bfloat16))
    out = torch.empty((M, N), device=cuda)
    need = L.dgcnn_tc_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=cuda)
    nv.check(L.dgcnn_tc_gemm(pa.data_ptr(), pb.data_ptr(), out.data_ptr(), M, N, K, tA, tB, 1, ws.data_ptr(), ws.numel(),
                             nv.stream_ptr(cuda)), "tc_gemm")
    a64 = pa[0].double().t() if tA else pa[0].double()
    b64 = pb[0].double().t() if tB else pb[0].double()
    ref = a64 @ b64
    assert (out.double() - ref).abs().max().item() <= 2e-6 * np.sqrt(K) * 9.0 + 1e-6 * ref.abs().max().item()
    # mixing modes is refused
    assert L.dgcnn_tc_gemm(pa.data_ptr(), pb.data_ptr(), out.data_ptr(), M, N, K, tA, tB, 3, ws.data_ptr(), ws.numel(),
                           nv.stream_ptr(cuda)) == nv.ERR_INVALID


@pytest.mark.parametrize("M,N,K,tA,tB", [(4096, 384, 256, 0, 1), (2048 + 40, 2176, 512, 0, 1), (1024, 288, 64, 0, 0),
                                          (2176, 512, 4096, 1, 0), (384, 1024, 8192, 1, 0)])
def test_wide_gemm_partial_last_column_tile(dg, cuda, M, N, K, tA, tB):
    """Widths that are not multiples of the 256-column tile (FC0 / MergedEdgeConv gradients with 6 EdgeConv layers:
    2176 and 384 columns) take the persistent wide kernel too: the missing columns of the last tile are zero-filled by
    TMA and never stored.  Guard columns after the output must stay untouched."""
    from dgcnn import _native as nv
    L = nv.lib()
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(cuda)
    B = torch.randn((N, K) if tB else (K, N), generator=g).to(cuda)
    buf = torch.full((M * N + 4096,), 7.0, device=cuda)            # output + guard
    need = L.dgcnn_tc_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=cuda)
    pa, pb = _split(dg, A), _split(dg, B)
    nv.check(L.dgcnn_tc_gemm(pa.data_ptr(), pb.data_ptr(), buf.data_ptr(), M, N, K, tA, tB, 2, ws.data_ptr(), ws.numel(),
                             nv.stream_ptr(cuda)), "tc_gemm")
    out = buf[:M * N].view(M, N)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    assert (out.double() - ref).abs().max().item() <= 4e-5 * float(np.sqrt(K)) * 9.0
    assert bool((buf[M * N:] == 7.0).all())
