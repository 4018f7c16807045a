"""BN-backward (statistics pass + apply pass) timed alone at the head's shapes of configs[1] (P = 49152 rows), L2 flushed
by rotating buffer sets:  python profiles/scripts/bn_bwd_bench.py
`sweep` (profiles/r02_bn_bwd_sweep.txt) additionally needed an experiment build whose debug hook `dgcnn_debug_bn_tune(blocks
per SM, rows in flight)` set the statistics kernel's grid; the hook was removed once the configuration was chosen, so with
the shipped library `sweep` only repeats the default line."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
import torch
from dgcnn import _native as nv

dev = torch.device("cuda", 0)
L = nv.lib()
P, B, N = 49152, 24, 2048
tune = getattr(ctypes.CDLL(nv.LIB_PATH), "dgcnn_debug_bn_tune", None) if "sweep" in sys.argv else None


def make(C, pool, gbias, sets=3):
    out = []
    for _ in range(sets):
        d = dict(z=torch.randn(P, C, device=dev), g=torch.randn(P, C, device=dev) * 1e-3,
                 planes=torch.empty(2, P, C, dtype=torch.bfloat16, device=dev))
        out.append(d)
    com = dict(beta=torch.zeros(C, device=dev), mean=torch.zeros(C, device=dev), rstd=torch.ones(C, device=dev),
               gbeta=torch.empty(C, device=dev), ws=torch.empty(L.dgcnn_bn_workspace_bytes(C) + 64, dtype=torch.uint8, device=dev))
    if pool:
        zz = out[0]["z"].view(B, N, C)
        com.update(pmax=torch.relu(zz).amax(1).contiguous(), pcnt=torch.ones(B, C, device=dev), pgrad=torch.randn(B, C, device=dev))
    if gbias:
        com["gb"] = torch.randn(B, C, device=dev) * 0.1
    return out, com


def call(d, com, C):
    p = nv.ptr
    rc = L.dgcnn_bn_act_bwd_planes(p(d["z"]), None, p(com["beta"]), p(d["g"]), P, C, p(com["mean"]), p(com["rstd"]),
                                   p(com["gb"]) if "gb" in com else None, N if "gb" in com else 0, 1, None, p(d["planes"]), 2,
                                   p(com["gbeta"]), p(com["pmax"]) if "pmax" in com else None,
                                   p(com["pcnt"]) if "pmax" in com else None, p(com["pgrad"]) if "pmax" in com else None,
                                   N if "pmax" in com else 0, p(com["ws"]), com["ws"].numel(), nv.stream_ptr(dev))
    nv.check(rc, "bn_act_bwd_planes")


def bench(C, pool, gbias, reps=12):
    sets, com = make(C, pool, gbias)
    for i in range(3):
        call(sets[i % len(sets)], com, C)
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(sets[i % len(sets)], com, C); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    ref = com["gbeta"].clone()
    return ts[len(ts) // 2], ref


shapes = [(1024, True, False), (512, False, True), (256, False, False), (64, False, False)]
configs = [(0, 2)] + ([(b, u) for u in (2, 4) for b in (2, 3, 4, 6)] if tune else [])   # (0, 2) = the library's own choice
base = {}
for bps, u in configs:
    if tune:
        tune(bps, u)
    row = []
    for C, pool, gb in shapes:
        torch.manual_seed(C)
        t, ref = bench(C, pool, gb)
        if (bps, u) == (0, 2):
            base[C] = ref
        err = float((ref - base[C]).abs().max() / base[C].abs().max())
        row.append("C=%4d %6.1f us (dgbeta %.1e)" % (C, t, err))
    print("blocks/SM %d rows-in-flight %d | %s" % (bps, u, " | ".join(row)), flush=True)
