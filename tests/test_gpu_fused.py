"""GPU parity of the round-1 fused pieces: wide tcgen05 GEMM + BN statistics from its epilogue, BN backward as bf16
planes, in-place pooling gradient, skinny streaming GEMMs, packed EdgeConv gradients, CUDA-graph replay of the
trainer micro-step, and the tensor-core k_nn at the shapes of BASELINE.json configs[2] / configs[4].
References are fp64 torch (floating point, tolerance stated per test) or the oracle (indices: bit-exact)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nv():
    from dgcnn import _native as nv
    return nv, nv.lib()


def _split(x):
    nv, L = _nv()
    x = x.contiguous()
    p = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    nv.check(L.dgcnn_split_bf16(x.data_ptr(), x.shape[0], x.shape[1], x.shape[1], p.data_ptr(), x.shape[1], x.numel(),
                                nv.stream_ptr(x.device)), "split")
    return p


@pytest.mark.parametrize("M,N,K", [(1024, 256, 64), (2048 + 72, 512, 328), (4096, 1024, 256)])
def test_wide_gemm_column_statistics(dg, cuda, M, N, K):
    """dgcnn_tc_gemm_stats: C matches fp64, colstats [tiles][2][N] are the per-128-row column sum / sum of squares."""
    nv, L = _nv()
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn((M, K), generator=g).to(cuda)
    B = torch.randn((K, N), generator=g).to(cuda)
    assert L.dgcnn_tc_gemm_stats_supported(M, N, K) == 1
    assert L.dgcnn_tc_gemm_stats_supported(M, N + 8, K) == 0
    tiles = (M + 127) // 128
    C = torch.empty((M, N), device=cuda)
    cs = torch.empty((tiles, 2, N), device=cuda)
    pa, pb = _split(A), _split(B)
    nv.check(L.dgcnn_tc_gemm_stats(pa.data_ptr(), pb.data_ptr(), C.data_ptr(), M, N, K, 0, 0, 2, cs.data_ptr(),
                                   nv.stream_ptr(cuda)), "tc_gemm_stats")
    ref = A.double() @ B.double()
    scale = float(np.sqrt(K)) * 9.0
    assert (C.double() - ref).abs().max().item() <= 4e-5 * scale
    pad = torch.zeros((tiles * 128, N), dtype=torch.float64, device=cuda)
    pad[:M] = ref
    t = pad.view(tiles, 128, N)
    assert torch.allclose(cs[:, 0].double(), t.sum(1), rtol=1e-4, atol=1e-3 * scale)
    assert torch.allclose(cs[:, 1].double(), (t * t).sum(1), rtol=1e-4, atol=1e-3 * scale * scale)


@pytest.mark.parametrize("grouped", [False, True])
def test_bn_statistics_from_tiles_match_direct(dg, cuda, grouped):
    """mean / rstd from tile partials (+ analytic per-group bias) == the two-pass kernel on z + bias."""
    nv, L = _nv()
    G, rows_per, C = 3, 256, 512
    P = G * rows_per
    g = torch.Generator().manual_seed(5)
    z = (torch.randn((P, C), generator=g) * 2 + 0.5).to(cuda)
    gb = torch.randn((G, C), generator=g).to(cuda) if grouped else None
    tiles = P // 128
    t = z.view(tiles, 128, C)
    cs = torch.stack([t.sum(1), (t * t).sum(1)], dim=1).contiguous()
    mean = torch.empty(C, device=cuda)
    rstd = torch.empty(C, device=cuda)
    nv.check(L.dgcnn_bn_stats_from_tiles(cs.data_ptr(), tiles, C, P, nv.ptr(gb), rows_per if grouped else 0,
                                         mean.data_ptr(), rstd.data_ptr(), nv.stream_ptr(cuda)), "tiles")
    zz = z.double() + (gb.double().repeat_interleave(rows_per, 0) if grouped else 0.0)
    m = zz.mean(0)
    v = zz.var(0, unbiased=False)
    assert torch.allclose(mean.double(), m, atol=1e-5)
    assert torch.allclose(rstd.double(), 1.0 / torch.sqrt(v + 1e-3), rtol=1e-5)


def test_bn_backward_planes_equal_fp32_path(dg, cuda):
    """dgcnn_bn_act_bwd_planes (mask re-evaluated from z, g_z as bf16 hi/lo planes) vs dgcnn_bn_act_bwd_gb."""
    nv, L = _nv()
    P, C = 1024, 256
    g = torch.Generator().manual_seed(11)
    z = torch.randn((P, C), generator=g).to(cuda)
    beta = (0.3 * torch.randn(C, generator=g)).to(cuda)
    go = torch.randn((P, C), generator=g).to(cuda)
    ws = torch.empty(L.dgcnn_bn_workspace_bytes(C), dtype=torch.uint8, device=cuda)
    out = torch.empty_like(z)
    mean, rstd = torch.empty(C, device=cuda), torch.empty(C, device=cuda)
    st = nv.stream_ptr(cuda)
    nv.check(L.dgcnn_bn_act_fwd_gb(z.data_ptr(), P, C, beta.data_ptr(), 0, 0, 0, 1, out.data_ptr(), mean.data_ptr(),
                                   rstd.data_ptr(), ws.data_ptr(), ws.numel(), st), "fwd")
    gz_ref, gb_ref = torch.empty_like(z), torch.empty(C, device=cuda)
    nv.check(L.dgcnn_bn_act_bwd_gb(z.data_ptr(), out.data_ptr(), go.data_ptr(), P, C, mean.data_ptr(), rstd.data_ptr(), 0,
                                   0, 1, gz_ref.data_ptr(), gb_ref.data_ptr(), 0, ws.data_ptr(), ws.numel(), st), "bwd")
    planes = torch.empty((2, P, C), dtype=torch.bfloat16, device=cuda)
    gz, gb = torch.empty_like(z), torch.empty(C, device=cuda)
    nv.check(L.dgcnn_bn_act_bwd_planes(z.data_ptr(), 0, beta.data_ptr(), go.data_ptr(), P, C, mean.data_ptr(),
                                       rstd.data_ptr(), 0, 0, 1, gz.data_ptr(), planes.data_ptr(), 2, gb.data_ptr(), 0, 0, 0, 0,
                                       ws.data_ptr(), ws.numel(), st), "bwd planes")
    assert torch.equal(gz, gz_ref) and torch.equal(gb, gb_ref)
    rec = planes[0].float() + planes[1].float()
    assert (rec - gz_ref).abs().max().item() <= gz_ref.abs().max().item() * 2.0 ** -16
    assert torch.equal(planes[0], gz_ref.to(torch.bfloat16))


def test_pool_and_pass_gradient(dg, cuda):
    """x used both directly and through the global max pool: fused in-place backward == autograd on amax (with ties)."""
    from dgcnn import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randint(-3, 4, (3, 50, 64), generator=g).float().to(cuda)     # integers: many exact ties
    w1 = torch.randn((3, 50, 64), generator=g).to(cuda)
    w2 = torch.randn((3, 64), generator=g).to(cuda)
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    xa, pa = ops.pool_and_pass(a)
    ((xa * w1).sum() + (pa * w2).sum()).backward()
    ((b * w1).sum() + (b.amax(dim=1) * w2).sum()).backward()
    assert torch.equal(pa, b.amax(dim=1))
    assert torch.allclose(a.grad, b.grad, atol=1e-6)


@pytest.mark.parametrize("M,N,K,tA", [(8192, 2, 256, 0), (5000, 3, 100, 0), (256, 2, 24576, 1), (3, 128, 16384, 1),
                                      (4, 64, 9000, 1)])
def test_skinny_gemm_shapes(dg, cuda, M, N, K, tA):
    """dgcnn_gemm on the class-score layer's shapes (streaming kernels) against fp64."""
    nv, L = _nv()
    g = torch.Generator().manual_seed(M + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(cuda)
    B = torch.randn((K, N), generator=g).to(cuda)
    out = torch.empty((M, N), device=cuda)
    need = L.dgcnn_gemm_workspace_bytes(M, N, K, tA, 0)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=cuda)
    nv.check(L.dgcnn_gemm(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, tA, 0, ws.data_ptr(), ws.numel(),
                          nv.stream_ptr(cuda)), "gemm")
    ref = (A.double().t() if tA else A.double()) @ B.double()
    assert (out.double() - ref).abs().max().item() <= 2e-6 * np.sqrt(K) * 9.0 + 1e-6 * ref.abs().max().item()


def test_edgeconv_packed_gradient_sources(dg, cuda):
    """ops._EdgeConvGather: gradients reaching max / mean / their concat through different consumers are summed inside
    the gather kernels; compare with a pure-torch restatement of the same math in fp64."""
    from dgcnn import ops
    B, N, F, k = 2, 96, 64, 9
    P = B * N
    g = torch.Generator().manual_seed(3)
    uv0 = torch.randn((P, 2 * F), generator=g)
    idx = torch.stack([torch.stack([torch.randperm(N, generator=g)[:k] for _ in range(N)]) for _ in range(B)]).int()
    beta0 = 0.1 * torch.randn(F, generator=g)
    wm, wa, wb = (torch.randn((P, F), generator=g), torch.randn((P, F), generator=g), torch.randn((P, 2 * F), generator=g))
    uv = uv0.to(cuda).requires_grad_(True)
    beta = beta0.to(cuda).requires_grad_(True)
    mx, mn, both = ops._EdgeConvGather.apply(uv, idx.to(cuda), beta, B, N, k, None)
    assert torch.equal(both[:, :F], mx) and torch.equal(both[:, F:], mn)
    ((mx * wm.to(cuda)).sum() + (mn * wa.to(cuda)).sum() + (both * wb.to(cuda)).sum()).backward()
    # fp64 reference
    uvr = uv0.double().requires_grad_(True)
    br = beta0.double().requires_grad_(True)
    u, v = uvr[:, :F], uvr[:, F:]
    flat = (idx.long() + (torch.arange(B) * N).view(B, 1, 1)).view(P, k)
    z = u[:, None, :] + v[flat]                                           # [P,k,F]
    zh = (z - z.mean((0, 1))) / torch.sqrt(z.var((0, 1), unbiased=False) + 1e-3)
    y = torch.relu(zh + br)
    rmx, rmn = y.amax(1), y.mean(1)
    rboth = torch.cat([rmx, rmn], 1)
    ((rmx * wm.double()).sum() + (rmn * wa.double()).sum() + (rboth * wb.double()).sum()).backward()
    assert torch.allclose(mx.detach().cpu().double(), rmx.detach(), atol=1e-5)
    assert torch.allclose(mn.detach().cpu().double(), rmn.detach(), atol=1e-5)
    assert torch.allclose(uv.grad.cpu().double(), uvr.grad, atol=2e-4, rtol=1e-3)
    assert torch.allclose(beta.grad.cpu().double(), br.grad, atol=2e-3, rtol=1e-3)


def _train_flags(B, N):
    from types import SimpleNamespace
    return SimpleNamespace(NUM_CLASS=2, MODEL_NAME="dgcnn", TRAIN=True, KVALUE=8, DEBUG=False, EDGE_CONV_LAYERS=2,
                           EDGE_CONV_FILTERS=64, FC_LAYERS=2, FC_FILTERS=[256, 256], LEARNING_RATE=1e-3, GPUS=[0],
                           MINIBATCH_SIZE=B, NUM_CHANNEL=3, WEIGHT_KEY="", SEED=0, BATCH_SIZE=B, NUM_POINT=N)


def test_cuda_graph_replay_equals_eager(dg, cuda, monkeypatch):
    """trainval captures the micro-step after two eager runs; gradients of a replayed step == the eager ones.
    (Dropout draws differ between runs, so it is disabled through an all-ones mask: TRAIN stays True.)"""
    from dgcnn import model as M
    B, N = 2, 512
    orig = M.build
    mask = torch.ones((B, N, 1, 256), device=cuda)
    monkeypatch.setattr(M, "build", lambda pc, fl, dropout_mask=None: orig(pc, fl, dropout_mask=mask * M.DROPOUT_KEEP))
    g = torch.Generator().manual_seed(4)
    x = torch.rand((B, N, 3), generator=g)
    y = torch.randint(0, 2, (B, N), generator=g)

    def grads(use_graph, reps):
        monkeypatch.setenv("DGCNN_CUDA_GRAPH", "1" if use_graph else "0")
        tr = dg.trainval(_train_flags(B, N))
        tr.initialize()
        out = None
        for _ in range(reps):
            tr.zero_gradients(None)
            res = tr.accum_gradient(None, [x], [y])
            out = (tr.variables.flat_grad.clone(), res[2])
        return out, tr

    (g_eager, l_eager), _ = grads(False, 1)
    (g_graph, l_graph), tr = grads(True, 4)            # 2 eager + capture/replay + replay
    assert len(tr._graphs) == 1
    assert abs(l_eager - l_graph) < 1e-6
    scale = g_eager.abs().max().item()
    assert (g_eager - g_graph).abs().max().item() <= 2e-4 * scale     # fp32 atomics reorder sums between runs


@pytest.mark.parametrize("B,N,C,k", [(2, 4096, 64, 40), (1, 16384, 64, 20), (2, 1000, 3, 20), (1, 2500, 16, 33),
                                     (3, 300, 7, 24)])
def test_knn_tensor_core_other_configs(dg, oracle, cuda, B, N, C, k):
    """Tensor-core k_nn at the N / k of configs[2] and configs[4], ragged N and odd channel counts: bit-exact."""
    rng = np.random.RandomState(N + k)
    x = rng.rand(B, N, C).astype(np.float32)
    if C >= 16:
        x = np.maximum(rng.randn(B, N, C).astype(np.float32), 0.0)         # post-ReLU-like features
    got = dg.ops.k_nn(torch.from_numpy(x).to(cuda), k).cpu().numpy()
    ref = oracle.k_nn(torch.from_numpy(x), k).numpy()
    assert got.dtype == np.int32 and np.array_equal(got, ref)


def test_knn_tensor_core_heavy_ties_and_offsets(dg, oracle, cuda):
    """Voxel lattices far from the origin (ties everywhere, large norms) and duplicated points: still bit-exact
    (overflowing rows go through the exact fallback queue)."""
    rng = np.random.RandomState(0)
    lat = rng.randint(0, 768, size=(2, 1024, 3)).astype(np.float32) + 5000.0
    dup = rng.rand(1, 512, 64).astype(np.float32)
    dup[:, 256:] = dup[:, :256]                                             # every point twice
    small = (rng.rand(1, 700, 3).astype(np.float32) - 0.5) * 1e-3            # tiny extents: fp16 scaling path
    for x, k in ((lat, 20), (dup, 20), (small, 16)):
        got = dg.ops.k_nn(torch.from_numpy(x).to(cuda), k).cpu().numpy()
        ref = oracle.k_nn(torch.from_numpy(x), k).numpy()
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("K,weighted", [(2, False), (5, True)])
def test_fused_loss_head_matches_torch(dg, cuda, K, weighted):
    """ops.softmax_xent (trainval.py:39-52 in one kernel): loss, accuracy and d loss / d logits vs torch in fp64."""
    from dgcnn import ops
    g = torch.Generator().manual_seed(K)
    P = 3000
    logits = torch.randn((P, K), generator=g) * 2
    logits[:50] = 0.0                                   # exact ties: argmax takes the first maximum
    labels = torch.randint(0, K, (P,), generator=g)
    w = (torch.rand(P, generator=g) + 0.5) if weighted else None
    a = logits.to(cuda).requires_grad_(True)
    loss, acc = ops.softmax_xent(a, labels.to(cuda), w.to(cuda) if weighted else None)
    (loss * 0.5).backward()
    b = logits.double().requires_grad_(True)
    xent = torch.nn.functional.cross_entropy(b, labels, reduction="none")
    ref = (xent * w.double()).mean() if weighted else xent.mean()
    (ref * 0.5).backward()
    racc = (b.argmax(1) == labels).double().mean()
    assert abs(loss.item() - ref.item()) < 2e-6 * max(1.0, abs(ref.item()))
    assert abs(acc.item() - racc.item()) < 1e-6
    assert torch.allclose(a.grad.cpu().double(), b.grad, atol=1e-9, rtol=1e-5)


def test_plane_sinks_do_not_change_the_model(dg, cuda, monkeypatch):
    """model.build with producer-filled operand planes (ops.PlaneSinks) == the same build with separate split passes:
    identical logits and gradients (the sinks only move where the bf16 hi/lo planes are written)."""
    from dgcnn import model as M
    B, N = 2, 512
    mask = torch.ones((B, N, 1, 256), device=cuda) * M.DROPOUT_KEEP
    g = torch.Generator().manual_seed(7)
    x = torch.rand((B, N, 3), generator=g)
    y = torch.randint(0, 2, (B, N), generator=g)
    orig = M.build
    monkeypatch.setattr(M, "build", lambda pc, fl, dropout_mask=None: orig(pc, fl, dropout_mask=mask))
    monkeypatch.setenv("DGCNN_CUDA_GRAPH", "0")
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DGCNN_PLANE_SINKS", mode)
        tr = dg.trainval(_train_flags(B, N))
        tr.initialize()
        tr.zero_gradients(None)
        r = tr.accum_gradient(None, [x], [y])
        sm = tr.inference(None, [x], [y])
        res[mode] = (tr.variables.flat_grad.clone(), r[2], torch.from_numpy(sm[0]))
    assert torch.equal(res["1"][2], res["0"][2])                       # forward: bit-identical softmax
    assert abs(res["1"][1] - res["0"][1]) < 1e-7
    scale = res["0"][0].abs().max().item()
    assert (res["1"][0] - res["0"][0]).abs().max().item() <= 2e-4 * scale    # backward: fp32 atomics reorder sums


def test_cuda_graph_cache_is_bounded(dg, cuda, monkeypatch):
    """Ragged data (a different N per batch, reference production mode -np -1) must not accumulate captured graphs:
    the trainer keeps at most DGCNN_CUDA_GRAPH_MAX of them (LRU) and keeps producing finite losses."""
    monkeypatch.setenv("DGCNN_CUDA_GRAPH", "1")
    monkeypatch.setenv("DGCNN_CUDA_GRAPH_MAX", "2")
    tr = dg.trainval(_train_flags(2, 512))
    tr.initialize()
    g = torch.Generator().manual_seed(9)
    for N in (512, 640, 768, 512, 640):
        x = torch.rand((2, N, 3), generator=g)
        y = torch.randint(0, 2, (2, N), generator=g)
        for _ in range(4):                                   # 2 eager + capture + replay per shape
            tr.zero_gradients(None)
            r = tr.accum_gradient(None, [x], [y])
            tr.apply_gradient(None)
            assert np.isfinite(r[2])
        assert len(tr._graphs) <= 2
    assert list(tr._graphs.keys())[-1][0] == (2, 640, 3)


def test_cached_graph_survives_workspace_growth(dg, cuda, monkeypatch):
    """A captured micro-step keeps the scratch buffers it was recorded with: when a later, larger batch (or an eager
    inference call) replaces the cached workspaces, replaying the older graph must still give the eager result --
    its scratch may not have been handed back to the allocator and reused by live tensors."""
    monkeypatch.setenv("DGCNN_CUDA_GRAPH", "1")
    monkeypatch.setenv("DGCNN_CUDA_GRAPH_MAX", "4")
    from dgcnn import _native
    _native._ws_cache.clear()                                # earlier tests of this process may have grown the scratch
    fl = _train_flags(2, 512)
    fl.LEARNING_RATE = 0.0                                   # parameters stay put: the same input gives the same loss
    tr = dg.trainval(fl)
    tr.initialize()
    g = torch.Generator().manual_seed(21)

    def run(x, y):
        tr.zero_gradients(None)
        r = tr.accum_gradient(None, [x], [y])
        grad = tr.variables.flat_grad.clone()
        tr.apply_gradient(None)
        return r[2], grad

    xs = torch.rand((2, 640, 3), generator=g)
    ys = torch.randint(0, 2, (2, 640), generator=g)
    monkeypatch.setenv("DGCNN_CUDA_GRAPH", "0")
    loss_eager, grad_eager = run(xs, ys)                     # dropout: eval the eager reference with a fixed seed
    monkeypatch.setenv("DGCNN_CUDA_GRAPH", "1")
    for _ in range(4):                                       # 2 eager + capture + replay
        run(xs, ys)
    assert (2, 640, 3) in [k[0] for k in tr._graphs]
    before = {id(t) for t in _native.workspaces_snapshot()}
    # grow every workspace: a much larger cloud, eagerly (inference) and as a training step
    xl = torch.rand((2, 2304, 3), generator=g)
    yl = torch.randint(0, 2, (2, 2304), generator=g)
    for _ in range(4):
        run(xl, yl)
    assert {id(t) for t in _native.workspaces_snapshot()} != before        # the cache really was replaced
    junk = [torch.full((1 << 20,), float("nan"), device=cuda) for _ in range(64)]   # reuse whatever was freed
    loss_a, grad_a = run(xs, ys)                             # replay of the OLD graph
    loss_b, grad_b = run(xs, ys)
    del junk
    assert np.isfinite(loss_a) and torch.isfinite(grad_a).all()
    # dropout draws differ between runs (philox offset advances), so compare the dropout-free part: the two replays
    # and the eager step agree on the loss to within the dropout noise, and the gradients are finite and similar
    assert abs(loss_a - loss_b) < 0.05 and abs(loss_a - loss_eager) < 0.05
    assert float((grad_a - grad_b).abs().max()) < 0.5 * float(grad_eager.abs().max()) + 1e-3   # dropout draws differ


def test_bn_statistics_survive_large_means(dg, cuda):
    """Channels whose mean is huge against their spread (mean 1000, std 0.01): the statistics kernels shift by a pivot
    row before squaring, so E[z^2] - E[z]^2 does not cancel.  Both the per-point BN (dgcnn_bn_act_fwd) and the EdgeConv
    gather statistics (dgcnn_edgeconv_fwd_stats) must return the fp64 mean / rstd."""
    nv, L = _nv()
    g = torch.Generator().manual_seed(8)
    P, C = 8192, 64
    z = (1000.0 + 0.01 * torch.randn((P, C), generator=g, dtype=torch.float64)).float().to(cuda)
    beta = torch.zeros(C, device=cuda)
    out = torch.empty_like(z)
    mean, rstd = torch.empty(C, device=cuda), torch.empty(C, device=cuda)
    ws = torch.empty(L.dgcnn_bn_workspace_bytes(C), dtype=torch.uint8, device=cuda)
    nv.check(L.dgcnn_bn_act_fwd(z.data_ptr(), P, C, beta.data_ptr(), 0, 0, out.data_ptr(), mean.data_ptr(),
                                rstd.data_ptr(), ws.data_ptr(), ws.numel(), nv.stream_ptr(cuda)), "bn")
    zd = z.double()
    ref_r = 1.0 / torch.sqrt(zd.var(0, unbiased=False) + 1e-3)
    assert torch.allclose(mean.double(), zd.mean(0), rtol=0, atol=1e-4)
    assert torch.allclose(rstd.double(), ref_r, rtol=1e-4)
    # EdgeConv statistics: z_ij = u_i + v_j with both halves far from zero
    B, N, F, k = 2, 256, 64, 20
    uv = torch.cat([500.0 + 0.01 * torch.randn((B * N, F), generator=g, dtype=torch.float64),
                    -300.0 + 0.01 * torch.randn((B * N, F), generator=g, dtype=torch.float64)], 1).float().to(cuda)
    idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).to(cuda)
    ws2 = torch.empty(L.dgcnn_edgeconv_workspace_bytes(F), dtype=torch.uint8, device=cuda)
    m2, r2 = torch.empty(F, device=cuda), torch.empty(F, device=cuda)
    nv.check(L.dgcnn_edgeconv_fwd_stats(uv.data_ptr(), nv.DT_F32, idx.data_ptr(), B, N, F, k, m2.data_ptr(), r2.data_ptr(),
                                        ws2.data_ptr(), ws2.numel(), nv.stream_ptr(cuda)), "ec stats")
    flat = (idx.long() + (torch.arange(B, device=cuda) * N).view(B, 1, 1)).view(B * N, k)
    zz = uv[:, :F].double()[:, None, :] + uv[:, F:].double()[flat]
    assert torch.allclose(m2.double(), zz.mean((0, 1)), rtol=0, atol=1e-4)
    assert torch.allclose(r2.double(), 1.0 / torch.sqrt(zz.var((0, 1), unbiased=False) + 1e-3), rtol=1e-4)
