"""Model-level parity AT THE BENCHMARKED CONFIGURATION (BASELINE.json configs[1]: B=24, N=2048, k=20, 4 EdgeConv layers,
FC 512/256, fp32) -- the path bench.py times: P = 49152 rows, FC0 on dgcnn_tc_gemm_stats with the per-cloud global
feature folded into the tile statistics, producer-filled operand planes, 3-pass EdgeConv gather kernels.
/root/reference/dgcnn/model.py:60-104, ops.py:91-96.

kNN in feature space is a discontinuous function of activations that differ from the oracle's by fp32 reassociation,
so the checks are split the only way that is meaningful:
  1. every layer's indices are BIT-EXACT against the C oracle run on the GPU's OWN layer input (no teacher forcing:
     the GPU's activations go through oracle.k_nn);
  2. with the GPU's indices handed to the oracle (same graph on both sides), logits <= 1e-3 everywhere (north_star),
     loss <= 1e-4, and every parameter gradient within a stated relative L2 error;
  3. free-running on both sides (each side its own graph), the fraction of logits beyond 1e-3 is measured and bounded.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, N, K, L, C0 = 24, 2048, 20, 4, 3


def _setup(dg, oracle, train):
    fl = oracle.make_flags(EDGE_CONV_LAYERS=L, EDGE_CONV_FILTERS=64, KVALUE=K, FC_LAYERS=2, FC_FILTERS=[512, 256],
                           NUM_CLASS=2, MODEL_NAME="dgcnn", TRAIN=train, NUM_CHANNEL=C0, MINIBATCH_SIZE=B)
    P = oracle.init_params(fl, C0, seed=0)
    g = torch.Generator().manual_seed(77)
    for n, t in P.items():
        if n.endswith("beta"):                       # non-trivial BN offsets (zeros-init would hide beta handling)
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
    x = torch.rand((B, N, C0), generator=g)
    y = torch.randint(0, 2, (B, N), generator=g)
    tr = dg.trainval(fl)
    tr.initialize()
    tr.variables.load_state_dict({"dgcnn/" + n: t for n, t in P.items()})
    return fl, P, x, y, tr, g


def _gpu_build(dg, tr, fl, x, mask):
    from dgcnn.variables import set_default_store
    old = set_default_store(tr.variables)
    try:
        with tr.variables.variable_scope("dgcnn"):
            return dg.build(x, fl, dropout_mask=mask)
    finally:
        set_default_store(old)


def test_config2_train_step_same_graph(dg, oracle, cuda):
    fl, P, x, y, tr, g = _setup(dg, oracle, True)
    mask = (torch.rand((B, N, 1, 256), generator=g) < 0.7).float()
    dg.ops._knn_trace, dg.ops._knn_input_trace = [], []
    try:
        tr.zero_gradients(None)
        logits = _gpu_build(dg, tr, fl, x.cuda(), mask.cuda())
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, 2), y.cuda().reshape(-1))
        loss.backward()
    finally:
        trace, dg.ops._knn_trace = dg.ops._knn_trace, None
        inputs, dg.ops._knn_input_trace = dg.ops._knn_input_trace, None
    assert len(trace) == L and len(inputs) == L
    # 1. bit-exact indices on the GPU's own activations, every layer
    for i in range(L):
        ref_idx = oracle.k_nn(inputs[i].cpu(), K)
        assert torch.equal(trace[i].cpu(), ref_idx), "layer %d: %d of %d indices differ" % (
            i, int((trace[i].cpu() != ref_idx).sum()), ref_idx.numel())
    # 2. same graph on both sides: the fp32 oracle (the reference's arithmetic) and, for the gradients, the fp64 oracle
    idx_list = [t.cpu() for t in trace]
    res = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        Pd = {n: t.detach().to(dt).requires_grad_(True) for n, t in P.items()}
        lg = oracle.build(x.to(dt), fl, Pd, idx_list=idx_list, dropout_mask=mask.to(dt))
        _, _, ls = oracle.softmax_loss_accuracy(lg, y)
        ls.backward()
        res[name] = ({n: t.grad.double() for n, t in Pd.items()}, lg.detach(), float(ls.detach()))
    g32, ref, ref_loss = res["f32"]
    g64 = res["f64"][0]
    err = (logits.detach().cpu() - ref).abs()
    print("configs[1] same-graph: max |logit diff| %.3g, mean %.3g, loss %.6f vs %.6f" % (
        err.max().item(), err.mean().item(), loss.item(), ref_loss))
    assert err.max().item() <= 1e-3                                          # north_star bound, every point
    assert abs(loss.item() - ref_loss) <= 1e-4
    # Gradients.  At this size fp32 itself is the limit: the Final layer's BN backward removes the mean and zhat
    # components of d loss / d logits, which dominate it for a random-init network, so the fp32 ORACLE differs from the
    # fp64 oracle by ~3e-3 (relative L2) on every tensor below Final (profiles/scripts/grad_error_probe.py).  The bound
    # is therefore relative to that: the GPU may not be more than 2.5x (+1e-4) further from fp64 than fp32 arithmetic is
    # (both numbers are rounding-noise amplitudes; they move by tens of percent with the summation order alone).
    worst = 0.0
    for n in P:
        a = tr.variables.vars["dgcnn/" + n].grad.cpu().double()
        den = max(float(g64[n].norm()), 1e-30)
        e_gpu, e_f32 = float((a - g64[n]).norm()) / den, float((g32[n] - g64[n]).norm()) / den
        print("   %-36s rel L2 error vs fp64 oracle: gpu %.3g, fp32 oracle %.3g" % (n, e_gpu, e_f32))
        assert e_gpu <= 2.5 * e_f32 + 1e-4, (n, e_gpu, e_f32)
        assert e_gpu <= 1e-2, (n, e_gpu)
        worst = max(worst, e_gpu)
    print("configs[1] same-graph: worst relative L2 gradient error vs fp64 %.3g" % worst)


def test_config2_free_running_violation_rate(dg, oracle, cuda):
    """Both sides compute their own neighbour graphs.  Where a near-tie resolves differently the two networks see
    different neighbourhoods and the logits of the affected points legitimately differ; the rate is measured here."""
    fl, P, x, y, tr, g = _setup(dg, oracle, False)
    dg.ops._knn_trace = []
    try:
        with torch.no_grad():
            logits = _gpu_build(dg, tr, fl, x.cuda(), None)
    finally:
        trace, dg.ops._knn_trace = dg.ops._knn_trace, None
    idx_ref = []
    with torch.no_grad():
        ref = oracle.build(x, fl, P, idx_out=idx_ref)
    assert torch.equal(trace[0].cpu(), idx_ref[0])                           # layer 0 input is the raw cloud
    agree = [float((trace[i].cpu() == idx_ref[i]).float().mean()) for i in range(L)]
    err = (logits.cpu() - ref).abs().amax(-1)                                # [B,N] worst class per point
    beyond = float((err > 1e-3).float().mean())
    # points whose L-hop receptive field contains no edge that resolved differently ("clean")
    dirty = torch.zeros((B, N), dtype=torch.bool)
    base = (torch.arange(B) * N).view(B, 1, 1)
    for i in range(L):
        a, b = trace[i].cpu().long(), idx_ref[i].long()
        flat = dirty.reshape(-1)
        nb = flat[(a + base).reshape(-1)].view(B, N, K).any(-1) | flat[(b + base).reshape(-1)].view(B, N, K).any(-1)
        dirty = dirty | nb | (a != b).any(-1)
    clean = ~dirty
    print("configs[1] free-running: index agreement per layer %s; points beyond 1e-3: %.3f %% (max diff %.3g); "
          "clean points %.2f %%, max diff on them %.3g, median diff overall %.3g" % (
              ["%.5f" % v for v in agree], 100 * beyond, err.max().item(), 100 * float(clean.float().mean()),
              err[clean].max().item() if clean.any() else 0.0, err.median().item()))
    q = torch.quantile(err.reshape(-1), torch.tensor([0.5, 0.9, 0.99]))
    print("configs[1] free-running: |logit diff| quantiles 50/90/99 %%: %.3g / %.3g / %.3g; beyond 1e-2: %.3f %%" % (
        q[0], q[1], q[2], 100 * float((err > 1e-2).float().mean())))
    assert min(agree) >= 0.995
    # Measured on a B200: 99.84-100 % of the indices agree per layer, yet ~22 % of the points differ
    # by more than 1e-3 and the median difference is 5e-4: a flipped edge changes its point, through the next graph layers
    # that point's neighbourhood, and through the global max-pooled feature (model.py:76-85) and the batch statistics every
    # point of the cloud a little.  Same-graph parity (the test above) is the meaningful 1e-3 criterion; here only the
    # size of the effect is bounded.
    assert err.median().item() <= 1e-3
    assert beyond <= 0.35
    assert float((err > 1e-2).float().mean()) <= 0.05
