"""k_nn timing at the N / k of the other BASELINE configs (random post-ReLU-like 64-channel features and xyz)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
import dgcnn
from dgcnn import _native as nv
def al(v): return (v + 255) // 256 * 256
g = torch.Generator().manual_seed(1)
for (B, N, C, k) in [(24, 2048, 64, 20), (24, 4096, 64, 40), (24, 4096, 3, 40), (8, 16384, 64, 20), (8, 16384, 3, 20)]:
    x = torch.rand((B, N, C), generator=g).cuda() if C == 3 else torch.relu(torch.randn((B, N, C), generator=g)).cuda()
    for _ in range(2): idx = dgcnn.ops.k_nn(x, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): idx = dgcnn.ops.k_nn(x, k)
    e1.record(); torch.cuda.synchronize()
    ws = nv._ws_cache[(0, "knn")]
    Npad = (N + 127) // 128 * 128
    base = al((B * C * Npad + B * Npad) * 4)
    P, Pp, Cp, cap = B * N, B * Npad, (C + 7) // 8 * 8, (32 if k <= 24 else 48)
    off = base + 2 * al(Pp * 4) + al(B * 16 * 2 * C * 4) + al(2 * Pp * Cp * 2) + al(2 * Pp * 32) + al(P * 4 * cap * 2)
    cc = ws[off:off + P * 4].view(P, 4).int()
    nq = int(ws[off + al(P * 4):off + al(P * 4) + 4].view(torch.int32)[0])
    ok = (cc < 255).all(1)
    print("B=%d N=%d C=%d k=%d: %.3f ms; fallback rows %d; candidates/row mean %.1f max %d" % (
        B, N, C, k, e0.elapsed_time(e1) / 5, int((~ok).sum()), cc.sum(1).float()[ok].mean().item(), int(cc.sum(1)[ok].max())))
