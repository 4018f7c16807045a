"""Times the EdgeConv gather passes alone through the C ABI (CUDA events, warm, back to back over a ring of buffers so
that the uv table of the call is NOT already in L1; it is L2 resident like in the model where the uv GEMM just wrote it).
usage: python profiles/scripts/prof_edge.py [B N F k [bf16]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
import torch  # noqa: E402
from dgcnn import _native as nv  # noqa: E402

B, N, F, k = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (24, 2048, 64, 20)
bf16 = len(sys.argv) > 5 and sys.argv[5] == "bf16"
dev = torch.device("cuda", 0)
P = B * N
L = nv.lib()
st = nv.stream_ptr(dev)
g = torch.Generator(device="cpu").manual_seed(0)
RING = 3
uvs = [torch.randn((P, 2 * F), generator=g).to(dev) for _ in range(RING)]
if bf16:
    uvs = [u.to(torch.bfloat16) for u in uvs]
dt = nv.DT_BF16 if bf16 else nv.DT_F32
# neighbours: random within the cloud (worst case for locality, like feature-space graphs of a random-init net)
idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).to(dev)
beta = (0.1 * torch.randn(F, generator=g)).to(dev)
mean = torch.empty(F, device=dev)
rstd = torch.empty(F, device=dev)
both = torch.empty((P, 2 * F), device=dev)
npos = torch.empty((P, F), dtype=torch.uint8, device=dev)
zmax = torch.empty((P, F), device=dev)
gboth = torch.randn((P, 2 * F), generator=g).to(dev)
gmax = torch.randn((P, F), generator=g).to(dev)
gmean = torch.randn((P, F), generator=g).to(dev)
s1 = torch.empty(F, device=dev)
s2 = torch.empty(F, device=dev)
guv = torch.empty((P, 2 * F), device=dev)
ws = torch.empty(L.dgcnn_edgeconv_workspace_bytes(F), dtype=torch.uint8, device=dev)


def fwd_stats(i):
    nv.check(L.dgcnn_edgeconv_fwd_stats(uvs[i % RING].data_ptr(), dt, idx.data_ptr(), B, N, F, k, mean.data_ptr(),
                                        rstd.data_ptr(), ws.data_ptr(), ws.numel(), st), "fs")


def fwd_apply(i):
    nv.check(L.dgcnn_edgeconv_fwd_apply(uvs[i % RING].data_ptr(), dt, idx.data_ptr(), B, N, F, k, mean.data_ptr(),
                                        rstd.data_ptr(), beta.data_ptr(), both.data_ptr(), zmax.data_ptr(), npos.data_ptr(), 0, 0, 0,
                                        st),
             "fa")


def bwd_stats(i):
    nv.check(L.dgcnn_edgeconv_bwd_stats(both.data_ptr(), npos.data_ptr(), beta.data_ptr(), gmax.data_ptr(),
                                        gmean.data_ptr(), gboth.data_ptr(), B, N, F, k, s1.data_ptr(), s2.data_ptr(),
                                        guv.data_ptr(), ws.data_ptr(), ws.numel(), st), "bs")


def bwd_apply(i):
    nv.check(L.dgcnn_edgeconv_bwd_apply(uvs[i % RING].data_ptr(), dt, idx.data_ptr(), B, N, F, k, mean.data_ptr(),
                                        rstd.data_ptr(), beta.data_ptr(), zmax.data_ptr(), gmax.data_ptr(),
                                        gmean.data_ptr(), gboth.data_ptr(), s1.data_ptr(), s2.data_ptr(), guv.data_ptr(), 1, st), "ba")


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


fwd_stats(0); fwd_apply(0); bwd_stats(0)
E = P * k
esz = 2 if bf16 else 4
for name, fn in (("fwd_stats", fwd_stats), ("fwd_apply", fwd_apply), ("bwd_stats", bwd_stats), ("bwd_apply", bwd_apply)):
    us = timeit(fn)
    print("%-10s %7.1f us   L2 gather %.2f TB/s" % (name, us, (E * F * esz / (us * 1e-6) / 1e12) if name != "bwd_stats" else 0.0))
print("shape B=%d N=%d F=%d k=%d dtype=%s  (times include the tiny memset/finalize launches of each entry)" %
      (B, N, F, k, "bf16" if bf16 else "fp32"))
