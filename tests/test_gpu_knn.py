"""GPU parity of the k_nn hot path (dgcnn_knn / dgcnn_pairwise_distance / dgcnn_topk_rows through the C ABI)
against the oracle: BIT-EXACT indices and distances, including tie order, ragged N, k edge cases."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cmp_knn(dg, oracle, x, k):
    got = dg.ops.k_nn(torch.from_numpy(x).cuda(), k).cpu().numpy()
    ref = oracle.k_nn(x, k).numpy()
    assert got.dtype == np.int32 and got.shape == ref.shape
    bad = (got != ref).sum()
    assert bad == 0, "%d / %d indices differ" % (bad, ref.size)


@pytest.mark.parametrize("B,N,C,k", [
    (2, 512, 3, 20),      # BASELINE configs[0]
    (3, 100, 3, 20),      # ragged: N not a multiple of the 64-row / 128-column tiles
    (2, 129, 4, 16),
    (1, 1, 3, 1),         # single point
    (2, 64, 64, 64),      # k == N == tile edge, two list registers per lane
    (2, 300, 64, 40),     # production k (scripts/lsf/train_dgcnn.sh:28)
    (1, 257, 17, 33),     # odd channel count across the 16-channel stage boundary
    (2, 200, 130, 7),     # many channel stages
    (1, 60, 5, 60),       # k == N, ragged
])
def test_knn_random_bit_exact(dg, oracle, cuda, B, N, C, k):
    rng = np.random.RandomState(B * 1000 + N + C + k)
    _cmp_knn(dg, oracle, rng.rand(B, N, C).astype(np.float32), k)
    _cmp_knn(dg, oracle, rng.randn(B, N, C).astype(np.float32) * 3.0, k)


def test_knn_lattice_ties_bit_exact(dg, oracle, cuda):
    rng = np.random.RandomState(0)
    for hi, N, k in ((6, 400, 20), (768, 1024, 20), (3, 333, 40)):
        x = rng.randint(0, hi, size=(2, N, 3)).astype(np.float32)   # voxel lattice: ties everywhere
        _cmp_knn(dg, oracle, x, k)
        got = dg.ops.k_nn(torch.from_numpy(x).cuda(), k).cpu().numpy()
        d = ((x[:, :, None, :].astype(np.int64) - x[:, None, :, :].astype(np.int64)) ** 2).sum(-1)
        assert np.array_equal(got, np.argsort(d, axis=-1, kind="stable")[:, :, :k])   # analytic answer


def test_knn_duplicates_and_hand_cases(dg, oracle, cuda):
    x = np.zeros((1, 6, 3), np.float32)
    x[0, 3:] = 1.0
    assert dg.ops.k_nn(torch.from_numpy(x).cuda(), 3).cpu().tolist() == [[[0, 1, 2]] * 3 + [[3, 4, 5]] * 3]
    line = torch.arange(5, dtype=torch.float32).reshape(1, 5, 1).cuda()
    assert dg.ops.k_nn(line, 3).cpu().tolist() == [[[0, 1, 2], [1, 0, 2], [2, 1, 3], [3, 2, 4], [4, 3, 2]]]
    _cmp_knn(dg, oracle, np.zeros((2, 150, 8), np.float32), 20)   # all points identical: pure index order


@pytest.mark.parametrize("B,N,C", [(2, 512, 3), (2, 130, 64), (1, 77, 9), (3, 256, 33)])
def test_pairwise_distance_bit_exact(dg, oracle, cuda, B, N, C):
    rng = np.random.RandomState(N + C)
    x = (rng.randn(B, N, C) * 2).astype(np.float32)
    got = dg.ops.pairwise_distance(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = oracle.pairwise_distance(x).numpy()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))   # bit for bit
    assert np.array_equal(got, np.swapaxes(got, 1, 2))                # symmetric by construction


@pytest.mark.parametrize("rows,N,k", [(64, 2048, 20), (10, 64, 64), (33, 129, 40), (7, 5, 1)])
def test_topk_rows_bit_exact(dg, oracle, cuda, rows, N, k):
    rng = np.random.RandomState(rows + N)
    D = rng.rand(rows, N).astype(np.float32)
    D[:, ::3] = np.round(D[:, ::3] * 8) / 8   # inject exact ties
    got = dg.ops.knn(torch.from_numpy(D).cuda(), k).cpu().numpy()
    assert np.array_equal(got, oracle.topk_rows(D, k).numpy())
    assert np.array_equal(got, np.argsort(D, axis=-1, kind="stable")[:, :k])


def test_fused_equals_unfused(dg, cuda):
    x = torch.rand(2, 777, 16, device=cuda)
    assert torch.equal(dg.ops.k_nn(x, 24), dg.ops.knn(dg.ops.pairwise_distance(x), 24))


def test_knn_error_behaviour(dg, cuda):
    x = torch.rand(1, 8, 3, device=cuda)
    with pytest.raises(ValueError):
        dg.ops.k_nn(x, 9)          # k > N
    with pytest.raises(ValueError):
        dg.ops.k_nn(x, 0)
    with pytest.raises(NotImplementedError):
        dg.ops.k_nn(torch.rand(1, 128, 3, device=cuda), 65)
    with pytest.raises(ValueError):
        dg.ops.k_nn(torch.rand(8, 3, device=cuda), 2)
    with pytest.raises(TypeError):
        dg.ops.k_nn(x.double(), 2)


def test_knn_full_size_config2(dg, oracle, cuda):
    """BASELINE configs[1] shape (B=24, N=2048, k=20), both channel counts of the model: whole-tensor equality
    with the oracle (it finishes in ~1 s at this size) plus size-independent properties."""
    g = torch.Generator().manual_seed(1234)
    for C in (3, 64):
        x = torch.rand((24, 2048, C), generator=g)
        xc = x.cuda()
        idx = dg.ops.k_nn(xc, 20)
        assert torch.equal(idx.cpu(), oracle.k_nn(x, 20))
        # sortedness of the selected distances, every index in range, no duplicates per row
        D = dg.ops.pairwise_distance(xc[:2])
        sel = torch.gather(D, 2, idx[:2].long())
        assert (sel[:, :, 1:] >= sel[:, :, :-1]).all()
        assert int(idx.min()) >= 0 and int(idx.max()) < 2048
        assert (torch.sort(idx, dim=-1).values.diff(dim=-1) > 0).all()
        # k-th selected distance bounds everything not selected
        kth = sel[:, :, -1:]
        mask = torch.ones_like(D, dtype=torch.bool).scatter_(2, idx[:2].long(), False)
        assert (D[mask].reshape(2, 2048, -1) >= kth).all()
        # determinism / idempotence
        assert torch.equal(idx, dg.ops.k_nn(xc, 20))


def test_knn_hint_never_changes_the_result(dg, oracle, cuda):
    """dgcnn_knn_hinted: any k distinct in-range indices per row are a valid warm start (good, bad or adversarial)."""
    rng = np.random.RandomState(5)
    for B, N, C, k in ((2, 700, 64, 20), (1, 130, 3, 40), (2, 256, 16, 64)):
        x = torch.from_numpy(rng.randn(B, N, C).astype(np.float32)).cuda()
        ref = dg.ops.k_nn(x, k)
        assert torch.equal(ref.cpu(), oracle.k_nn(x.cpu(), k))
        good = ref                                                           # the answer itself
        near = dg.ops.k_nn(x + 0.05 * torch.randn_like(x), k)                # a similar graph (previous layer)
        rnd = torch.stack([torch.stack([torch.randperm(N)[:k] for _ in range(N)]) for _ in range(B)]).int().cuda()
        far = torch.flip(dg.ops.k_nn(-x, k), dims=[-1])                      # unrelated graph
        for hint in (good, near, rnd, far):
            assert torch.equal(dg.ops.k_nn(x, k, hint=hint), ref)


@pytest.mark.parametrize("B,N,C,k", [(2, 700, 64, 20), (3, 2048, 64, 20), (2, 384, 64, 40), (1, 130, 8, 5),
                                      (2, 1000, 32, 24), (2, 256, 64, 56), (1, 128, 16, 1)])
def test_knn_tensor_core_path_bit_exact(dg, oracle, cuda, B, N, C, k):
    """The tcgen05 filter + exact refinement (taken for clouds of >= 256 points with C <= 64 and k <= 48; smaller clouds
    and larger k take the SIMT kernel, the only path that uses a hint) must reproduce the oracle bit for bit whatever
    hint the caller passes, including useless ones."""
    rng = np.random.RandomState(N + C + k)
    x = torch.from_numpy((rng.randn(B, N, C) * rng.uniform(0.2, 3.0, size=(1, 1, C))).astype(np.float32)).cuda()
    ref = oracle.k_nn(x.cpu(), k).cuda()
    near = dg.ops.k_nn(x + 0.05 * torch.randn_like(x), k)
    rnd = torch.stack([torch.stack([torch.randperm(N)[:k] for _ in range(N)]) for _ in range(B)]).int().cuda()
    for name, hint in (("exact", ref), ("near", near), ("random", rnd)):
        got = dg.ops.k_nn(x, k, hint=hint)
        bad = int((got != ref).sum())
        assert bad == 0, "%s hint: %d / %d indices differ" % (name, bad, ref.numel())


def test_knn_tensor_core_path_ties_fall_back_exactly(dg, oracle, cuda):
    """Voxel-lattice features and duplicated points: distance ties far beyond the filter's resolution.  Rows that
    cannot be certified are recomputed by the exact SIMT kernel; the result is still the oracle's."""
    rng = np.random.RandomState(3)
    for x in (rng.randint(0, 3, size=(2, 512, 64)).astype(np.float32),          # lattice in 64-d
              np.repeat(rng.rand(2, 64, 64).astype(np.float32), 8, axis=1),      # every point 8 times
              np.zeros((1, 300, 64), np.float32)):                              # all identical
        xc = torch.from_numpy(x).cuda()
        ref = oracle.k_nn(x, 20).cuda()
        hint = torch.stack([torch.stack([torch.randperm(x.shape[1])[:20] for _ in range(x.shape[1])])
                            for _ in range(x.shape[0])]).int().cuda()
        assert torch.equal(dg.ops.k_nn(xc, 20, hint=ref), ref)
        assert torch.equal(dg.ops.k_nn(xc, 20, hint=hint), ref)


def test_knn_filter_modes_agree(dg, oracle, cuda):
    """dgcnn_knn_mode: the coarse and the fine precision mode of the tensor-core filter (and the library's own rule) give
    the same, oracle-exact result; the switch is an argument, the library holds no state."""
    from dgcnn import _native as nv
    L = nv.lib()
    g = torch.Generator().manual_seed(31)
    x = torch.rand((2, 640, 64), generator=g)
    ref = oracle.k_nn(x, 20)
    xc = x.cuda()
    ws = torch.empty(L.dgcnn_knn_workspace_bytes(2, 640, 64), dtype=torch.uint8, device=cuda)
    for mode in (-1, 0, 1):
        idx = torch.empty((2, 640, 20), dtype=torch.int32, device=cuda)
        nv.check(L.dgcnn_knn_mode(xc.data_ptr(), 0, idx.data_ptr(), 2, 640, 64, 20, mode, ws.data_ptr(), ws.numel(),
                                  nv.stream_ptr(cuda)), "knn_mode")
        assert torch.equal(idx.cpu(), ref), mode
    assert L.dgcnn_knn_mode(xc.data_ptr(), 0, idx.data_ptr(), 2, 640, 64, 20, 5, ws.data_ptr(), ws.numel(),
                            nv.stream_ptr(cuda)) == nv.ERR_INVALID


def test_knn_and_edges_equal_the_reference_code_bit_for_bit(dg, cuda):
    """tests/golden/ref_knn_edges.npz: produced by the REFERENCE'S OWN k_nn / edges (ops.py:8-40, executed unmodified through
    oracle/tf1_shim by tests/golden/make_reference_golden.py) on clouds whose distances are exact in fp32 -- duplicates and
    lattice ties included; tf.nn.top_k's rule (lower index first) is the only freedom.  The CUDA path must equal them."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_knn_edges.npz"))
    for name in ("dyadic3", "lattice3", "feat8"):
        x = torch.from_numpy(z[name + ":x"]).cuda()
        for k in (1, 7, 20):
            got = dg.ops.k_nn(x, k).cpu().numpy()
            assert np.array_equal(got, z["%s:k%d:idx" % (name, k)]), (name, k)
        e = dg.ops.edges(x, k=5).cpu().numpy()
        assert np.array_equal(e, z[name + ":edges"]), name
