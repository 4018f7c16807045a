"""CPU check of the error budget that makes the tensor-core k_nn filter exact (dynamic-gcnn_b200/csrc/knn_tc.cu header).

The kernel keeps a column j for row i iff a LOWER bound of the oracle distance D_ij clears an UPPER bound of the row's
k-th smallest distance.  Both bounds come from  D~_ij = n_i + n_j - 2 g_ij  (g = products of the fp16-rounded, centred,
power-of-two scaled points, accumulated in fp32) widened by sig_i + sig_j.  The filter is exact as long as
    |D~_ij - 4^e * D_oracle_ij| <= sig_i + sig_j      for every pair,
which is what this test verifies with numpy restatements of knn_tc_prep_kernel (centring, scaling, fp16 rounding,
sig) and of the oracle's arithmetic, on the cloud families the GPU tests use.  The accumulation order of the tensor
core is not reproducible on a CPU; the fp32 matmul used here has the same operand roundings (the dominant term) and an
accumulation error of the same class, and the measured slack shows how far the budget is from being tight."""
import numpy as np
import pytest

EPS = {"coarse": 1.125 / 1024.0, "fine": 1.0 / 8192.0}
EPS_ORACLE = 1.0 / 65536.0
EPS_ABS = 1.0 / 2097152.0


def _oracle_d(x):
    """oracle/knn_oracle.c restated: s sequential sum of rounded squares, p sequential fmaf chain (float64 product
    rounded once to fp32 per step is exactly fmaf for fp32 inputs), D = fl(fl(s_i+s_j) - 2p)."""
    n, c = x.shape
    s = np.zeros(n, np.float32)
    for k in range(c):
        s = (s + (x[:, k] * x[:, k]).astype(np.float32)).astype(np.float32)
    p = np.zeros((n, n), np.float32)
    for k in range(c):
        p = (x[:, k].astype(np.float64)[:, None] * x[:, k].astype(np.float64)[None, :] + p.astype(np.float64)).astype(np.float32)
    d = ((s[:, None] + s[None, :]).astype(np.float32) - (2.0 * p).astype(np.float32)).astype(np.float32)
    return s, d


def _prep(x, s, mode):
    """knn_tc_prep_kernel restated for one cloud."""
    mn, mx = x.min(0), x.max(0)
    mid = (np.float32(0.5) * mn + np.float32(0.5) * mx).astype(np.float32)
    ext = float(np.maximum(mx - mid, mid - mn).max())
    e = 0
    if ext > 0:
        _, xe = np.frexp(np.float32(ext))
        e = int(np.clip(13 - xe, -60, 60))
    sc = np.float32(2.0 ** e)
    y = ((x - mid).astype(np.float32) * sc).astype(np.float32)
    n = np.zeros(len(x), np.float32)
    for k in range(x.shape[1]):
        n = (y[:, k].astype(np.float64) * y[:, k].astype(np.float64) + n.astype(np.float64)).astype(np.float32)
    hi = y.astype(np.float16)
    lo = (y - hi.astype(np.float32)).astype(np.float16)
    sig = (np.float32(EPS[mode]) * n + np.float32(EPS_ORACLE) * (s * np.float32(4.0 ** e)) +
           np.float32(EPS_ABS) * np.sqrt(n)).astype(np.float32)
    return hi, lo, n, sig, e


def _clouds():
    rng = np.random.RandomState(0)
    yield "uniform xyz", rng.rand(300, 3).astype(np.float32)
    yield "gaussian x3", (rng.randn(300, 3) * 3).astype(np.float32)
    yield "lattice + offset", (rng.randint(0, 768, (300, 3)) + 5000).astype(np.float32)
    yield "post-ReLU 64ch", np.maximum(rng.randn(300, 64), 0).astype(np.float32)
    yield "tiny extent", ((rng.rand(300, 3) - 0.5) * 1e-3).astype(np.float32)
    yield "wide dynamic range", (rng.randn(300, 16) * np.exp(rng.randn(300, 1) * 3)).astype(np.float32)
    d = rng.rand(150, 8).astype(np.float32)
    yield "duplicates", np.concatenate([d, d]).astype(np.float32)


@pytest.mark.parametrize("mode", ["coarse", "fine"])
def test_filter_intervals_contain_the_oracle_distance(mode):
    worst = 0.0
    for name, x in _clouds():
        s, d = _oracle_d(x)
        hi, lo, n, sig, e = _prep(x, s, mode)
        h32, l32 = hi.astype(np.float32), lo.astype(np.float32)
        g = h32 @ h32.T
        if mode == "fine":
            g = g + h32 @ l32.T + l32 @ h32.T
        dt = (n[:, None].astype(np.float64) + n[None, :] - 2.0 * g.astype(np.float64))
        err = np.abs(dt - d.astype(np.float64) * 4.0 ** e)
        bud = sig[:, None].astype(np.float64) + sig[None, :]
        ok = err <= bud
        assert ok.all(), "%s (%s): interval misses the oracle distance, worst ratio %.3f" % (
            name, mode, float((err / np.maximum(bud, 1e-300)).max()))
        worst = max(worst, float((err / np.maximum(bud, 1e-300)).max()))
    assert worst < 0.9        # the budget is not tight: largest observed error / budget over all families


def test_candidate_rule_keeps_every_true_neighbour():
    """End to end on the CPU: threshold from the k-th largest group maximum (sweep 1), admission by the lower bound
    (sweep 2) -> the candidate set contains the oracle's k nearest (ties included) for every row."""
    rng = np.random.RandomState(3)
    k = 20
    for x in (rng.rand(512, 3).astype(np.float32), np.maximum(rng.randn(512, 64), 0).astype(np.float32)):
        s, d = _oracle_d(x)
        hi, lo, n, sig, e = _prep(x, s, "coarse")
        h32 = hi.astype(np.float32)
        g = (h32 @ h32.T).astype(np.float32)
        a1 = g - np.float32(0.5) * (n + sig)[None, :]          # sweep 1 accumulator (upper bounds)
        a2 = g - np.float32(0.5) * (n - sig)[None, :]          # sweep 2 accumulator (lower bounds)
        gm = a1.reshape(len(x), -1, 16).max(2)                 # group maxima, 16 columns per group
        ak = -np.sort(-gm, axis=1)[:, k - 1]                   # k-th largest group maximum
        cand = a2 >= (ak - sig)[:, None]
        kth = np.sort(d, axis=1)[:, k - 1]
        need = d <= kth[:, None]                               # every column at or below the k-th distance
        assert (cand | ~need).all()
        assert cand.sum(1).mean() < 2.0 * k                    # and the filter is selective
