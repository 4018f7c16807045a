"""Small-shape pass over every hand-written kernel family, meant to run under compute-sanitizer
(profiles/scripts/sanitize.sh: memcheck, racecheck, synccheck).  Shapes are the smallest that still take the production
code paths: tensor-core k_nn (coarse and fine filter), tcgen05 GEMMs (narrow, wide + BN statistics, split-K, grouped),
EdgeConv gather passes (exact k = 20 and generic k, fp32 / fp16 table), BN kernels, one trainer micro-step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
os.environ["DGCNN_CUDA_GRAPH"] = "0"
import torch  # noqa: E402
import dgcnn  # noqa: E402
from dgcnn import _native as nv, ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
L = nv.lib()

# k_nn: tensor-core path, both filter modes, 64 and 3 channels
for C in (64, 3):
    x = torch.rand((2, 384, C), generator=g).to(dev)
    ref = None
    for mode in (0, 1):
        ops._KNN_FILTER_MODE = mode
        idx = ops.k_nn(x, 20)
        ref = idx if ref is None else ref
        assert torch.equal(idx, ref)
ops._KNN_FILTER_MODE = -1
torch.cuda.synchronize()
print("k_nn ok")

# EdgeConv gather passes
for k, F, prec in ((20, 64, "f32"), (7, 32, "f32"), (20, 64, "bf16"), (40, 64, "f32")):
    B, N = 2, 128
    uv = torch.randn((B * N, 2 * F), generator=g).to(dev).requires_grad_(True)
    idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).to(dev)
    beta = torch.zeros(F, device=dev, requires_grad=True)
    ops._precision = prec
    mx, mn, both = ops._EdgeConvGather.apply(uv, idx, beta, B, N, k, None)
    (both.sum() + mx.sum()).backward()
    ops._precision = "f32"
torch.cuda.synchronize()
print("edgeconv ok")

# tcgen05 GEMMs
for (M, N, K, tA, tB) in ((1024, 64, 128, 0, 0), (1024, 256, 64, 0, 0), (128, 64, 2048, 1, 0), (512, 256, 4096, 1, 0),
                          (1024, 512, 256, 0, 1)):
    for npl in (2, 1):
        A = torch.randn((K, M) if tA else (M, K), generator=g).to(dev)
        Bm = torch.randn((N, K) if tB else (K, N), generator=g).to(dev)
        out = ops._tc_gemm_raw(ops._split(A, npl), ops._split(Bm, npl), M, N, K, tA, tB)
        refm = (A.t() if tA else A) @ (Bm.t() if tB else Bm)
        assert (out - refm).abs().max() < (0.5 if npl == 1 else 1e-2) * max(1.0, float(refm.abs().max()) * 0.02 + 1)
torch.cuda.synchronize()
print("tc gemm ok")

# one training micro-step of each model family (head: wide GEMM + stats, grouped dX, BN, pool, loss, Adam)
from types import SimpleNamespace
for model, dt in (("dgcnn", "f32"), ("residual-dgcnn", "bf16")):
    fl = SimpleNamespace(NUM_CLASS=2, MODEL_NAME=model, TRAIN=True, KVALUE=20, DEBUG=False, EDGE_CONV_LAYERS=2,
                         EDGE_CONV_FILTERS=64, FC_LAYERS=2, FC_FILTERS=[256, 256], LEARNING_RATE=1e-3, GPUS=[0],
                         MINIBATCH_SIZE=2, NUM_CHANNEL=3, WEIGHT_KEY="", SEED=0, BATCH_SIZE=2, NUM_POINT=512, DTYPE=dt)
    tr = dgcnn.trainval(fl)
    tr.initialize()
    x = torch.rand((2, 512, 3), generator=g)
    y = torch.randint(0, 2, (2, 512), generator=g)
    tr.zero_gradients(None)
    r = tr.accum_gradient(None, [x], [y])
    tr.apply_gradient(None)
    torch.cuda.synchronize()
    print("train step ok", model, dt, r[2])
print("sanitize_smoke done")
