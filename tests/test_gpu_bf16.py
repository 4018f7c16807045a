"""The reduced-precision variant (BASELINE.json configs[2]: residual-dgcnn, bf16): flags.DTYPE = "bf16".
Every 1x1 convolution (uv, conv1, shortcut, MergedEdgeConv, FC*, and all their gradients) runs as ONE bf16 tcgen05 MMA per
product (operands rounded to 8 mantissa bits, fp32 accumulation) and the EdgeConv gather passes read an fp16 uv table
(half the gather bytes); BatchNorm statistics, activations in HBM and k_nn stay fp32.

What is held against the fp32 oracle (/root/reference/dgcnn/ops.py:100-140, model.py:9-106), on the SAME neighbour graph:
  * k_nn stays bit-exact on the GPU's own (now bf16-perturbed) activations -- it is the same fp32 kernel;
  * the yardstick is the reference graph itself executed with bf16-rounded matmul operands (oracle.bf16_matmul()): on
    the 6-layer residual model that restatement is 0.24 (max) / 0.022 (mean) away from the fp32 logits and its parameter
    gradients differ by 44 % (median relative L2) -- train-mode BatchNorm on a random-initialised net with random labels
    makes the gradients ill-conditioned (already 3e-3 between fp32 and fp64, tests/test_gpu_config2_parity.py).
    The GPU variant must be NO FURTHER from fp32 than 1.25x that restatement (+ small absolute slack), and within the
    absolute tolerances BF16_LOGIT_TOL = 0.5 (max) / 0.05 (mean), loss 5e-3.
The fp32 path's 1e-3 bound does NOT apply to this variant; these are the stated tolerances of the bf16 variant."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_LOGIT_TOL = 0.5
BF16_LOGIT_MEAN_TOL = 0.05


def _run(dg, oracle, cuda, model, B, N, k, L, dtype):
    fl = oracle.make_flags(EDGE_CONV_LAYERS=L, EDGE_CONV_FILTERS=64, KVALUE=k, FC_LAYERS=2, FC_FILTERS=[512, 256],
                           NUM_CLASS=2, MODEL_NAME=model, TRAIN=True, NUM_CHANNEL=3, MINIBATCH_SIZE=B)
    fl.DTYPE = dtype
    P = oracle.init_params(fl, 3, seed=4)
    g = torch.Generator().manual_seed(5)
    for n, t in P.items():
        if n.endswith("beta"):
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
    x = torch.rand((B, N, 3), generator=g)
    y = torch.randint(0, 2, (B, N), generator=g)
    mask = (torch.rand((B, N, 1, 256), generator=g) < 0.7).float()
    tr = dg.trainval(fl)
    tr.initialize()
    tr.variables.load_state_dict({"dgcnn/" + n: t for n, t in P.items()})
    from dgcnn.variables import set_default_store
    dg.ops._knn_trace, dg.ops._knn_input_trace = [], []
    try:
        tr.zero_gradients(None)
        old = set_default_store(tr.variables)
        with tr.variables.variable_scope("dgcnn"):
            logits = dg.build(x.cuda(), fl, dropout_mask=mask.cuda())
        set_default_store(old)
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, 2), y.cuda().reshape(-1))
        loss.backward()
    finally:
        trace, dg.ops._knn_trace = dg.ops._knn_trace, None
        inputs, dg.ops._knn_input_trace = dg.ops._knn_input_trace, None
    for i in range(L):
        assert torch.equal(trace[i].cpu(), oracle.k_nn(inputs[i].cpu(), k)), "layer %d kNN not bit-exact" % i
    for t in P.values():
        t.requires_grad_(True)
    idx_list = [t.cpu() for t in trace]
    ref = oracle.build(x, fl, P, idx_list=idx_list, dropout_mask=mask)
    _, _, ref_loss = oracle.softmax_loss_accuracy(ref, y)
    ref_loss.backward()
    err = (logits.detach().cpu() - ref.detach()).abs()
    rel = {n: float((tr.variables.vars["dgcnn/" + n].grad.cpu() - t.grad).norm() / max(float(t.grad.norm()), 1e-12))
           for n, t in P.items()}
    yard = None
    if dtype == "bf16":     # the reference graph with bf16-rounded matmul operands, against the same fp32 oracle
        Pb = {n: t.detach().clone().requires_grad_(True) for n, t in P.items()}
        with oracle.bf16_matmul():
            lb = oracle.build(x, fl, Pb, idx_list=idx_list, dropout_mask=mask)
            oracle.softmax_loss_accuracy(lb, y)[2].backward()
        eb = (lb.detach() - ref.detach()).abs()
        yard = (eb.max().item(), eb.mean().item(),
                float(np.median([float((Pb[n].grad - P[n].grad).norm() / max(float(P[n].grad.norm()), 1e-12)) for n in P])))
    return err, float(loss), float(ref_loss), rel, yard


@pytest.mark.parametrize("model,B,N,k,L", [("residual-dgcnn", 2, 1024, 40, 6), ("dgcnn", 4, 512, 20, 4)])
def test_bf16_variant_against_fp32_oracle(dg, oracle, cuda, model, B, N, k, L):
    err, loss, ref_loss, rel, yard = _run(dg, oracle, cuda, model, B, N, k, L, "bf16")
    med = float(np.median(list(rel.values())))
    print("bf16 %s B=%d N=%d k=%d L=%d: logits max |diff| %.3g mean %.3g (bf16 restatement of the reference: %.3g / %.3g); "
          "loss %.5f vs %.5f; median grad rel L2 %.3g (restatement %.3g)" % (
              model, B, N, k, L, err.max().item(), err.mean().item(), yard[0], yard[1], loss, ref_loss, med, yard[2]))
    assert err.max().item() <= BF16_LOGIT_TOL and err.max().item() <= 1.25 * yard[0] + 0.05
    assert err.mean().item() <= BF16_LOGIT_MEAN_TOL and err.mean().item() <= 1.25 * yard[1] + 2e-3
    assert abs(loss - ref_loss) <= 5e-3
    assert med <= 1.25 * yard[2] + 0.02, (med, yard[2])


def test_fp32_path_unchanged_by_the_mode_switch(dg, oracle, cuda):
    """Same shapes on DTYPE=f32 right after a bf16 run: the mode is scoped to build(), nothing leaks."""
    _run(dg, oracle, cuda, "dgcnn", 2, 512, 20, 2, "bf16")
    err, loss, ref_loss, rel, _ = _run(dg, oracle, cuda, "dgcnn", 2, 512, 20, 2, "f32")
    assert err.max().item() <= 1e-3 and abs(loss - ref_loss) <= 1e-4
    assert dg.ops._precision == "f32"


@pytest.mark.parametrize("tdt", [torch.float16, torch.bfloat16])
def test_edgeconv_gather_reduced_table(dg, cuda, tdt):
    """_EdgeConvGather with the uv table in fp16 / bf16 == the fp64 restatement evaluated on the rounded table."""
    from dgcnn import ops
    B, N, F, k = 2, 160, 64, 20
    P = B * N
    g = torch.Generator().manual_seed(13)
    uv0 = torch.randn((P, 2 * F), generator=g)
    idx = torch.stack([torch.stack([torch.randperm(N, generator=g)[:k] for _ in range(N)]) for _ in range(B)]).int()
    beta0 = 0.1 * torch.randn(F, generator=g)
    wb = torch.randn((P, 2 * F), generator=g)
    uv = uv0.to(cuda).requires_grad_(True)
    beta = beta0.to(cuda).requires_grad_(True)
    old = ops._precision, ops._bf16_table
    ops._precision, ops._bf16_table = "bf16", tdt
    try:
        mx, mn, both = ops._EdgeConvGather.apply(uv, idx.to(cuda), beta, B, N, k, None)
        (both * wb.to(cuda)).sum().backward()
    finally:
        ops._precision, ops._bf16_table = old
    uvr = uv0.to(tdt).double().requires_grad_(True)
    br = beta0.double().requires_grad_(True)
    u, v = uvr[:, :F], uvr[:, F:]
    flat = (idx.long() + (torch.arange(B) * N).view(B, 1, 1)).view(P, k)
    z = u[:, None, :] + v[flat]
    zh = (z - z.mean((0, 1))) / torch.sqrt(z.var((0, 1), unbiased=False) + 1e-3)
    yv = torch.relu(zh + br)
    rboth = torch.cat([yv.amax(1), yv.mean(1)], 1)
    (rboth * wb.double()).sum().backward()
    assert torch.allclose(both.detach().cpu().double(), rboth.detach(), atol=1e-5)
    assert torch.allclose(uv.grad.cpu().double(), uvr.grad, atol=2e-4, rtol=1e-3)
    assert torch.allclose(beta.grad.cpu().double(), br.grad, atol=2e-3, rtol=1e-3)
