"""Runs the fused k_nn a few times at the BASELINE configs[1] shapes (for ncu captures)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
import dgcnn
g = torch.Generator().manual_seed(1234)
torch.manual_seed(0)
for C in (3, 64):
    x = torch.rand((24, 2048, C), generator=g).cuda()
    for _ in range(3):
        idx = dgcnn.ops.k_nn(x, 20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        idx = dgcnn.ops.k_nn(x, 20)
    e1.record(); torch.cuda.synchronize()
    print("C=%d k_nn %.3f ms" % (C, e0.elapsed_time(e1) / 10))
    hint = dgcnn.ops.k_nn(x + 0.02 * torch.randn_like(x), 20)
    e0.record()
    for _ in range(10):
        idx2 = dgcnn.ops.k_nn(x, 20, hint=hint)
    e1.record(); torch.cuda.synchronize()
    assert torch.equal(idx, idx2)
    print("C=%d k_nn hinted (perturbed graph) %.3f ms" % (C, e0.elapsed_time(e1) / 10))
