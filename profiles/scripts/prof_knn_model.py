"""kNN on the real model's features (layer inputs of the random-init DGCNN at configs[1]): per-layer time of the
hinted k_nn and the number of rows the tensor-core filter could not certify."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200")); sys.path.insert(0, ROOT)
import bench, dgcnn
from dgcnn import ops, _native as nv
fl = bench.make_flags(1); fl.TRAIN = False
tr = dgcnn.trainval(fl); tr.initialize()
g = torch.Generator().manual_seed(1234)
x = torch.rand((24, 2048, 3), generator=g).cuda()
feats = []
orig = ops._layer_knn
def spy(xx, k, hint=None):
    feats.append((xx.detach().clone(), None if hint is None else hint.clone()))
    return orig(xx, k, hint)
ops._layer_knn = spy
with torch.no_grad():
    tr.inference(None, [x])
ops._layer_knn = orig
L = nv.lib()
for li, (f, h) in enumerate(feats):
    B, N, C = f.shape
    for _ in range(2): idx = ops.k_nn(f, 20, hint=h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): idx = ops.k_nn(f, 20, hint=h)
    e1.record(); torch.cuda.synchronize()
    msg = "layer %d C=%d hinted=%s: %.3f ms" % (li, C, h is not None, e0.elapsed_time(e1) / 10)
    if h is not None and C % 8 == 0:
        ws = nv._ws_cache[(0, "knn")]
        Npad = (N + 127) // 128 * 128
        base = (((B * C * Npad + B * Npad) * 4) + 255) // 256 * 256
        P = B * N
        off = base + ((2 * P * C * 2 + 255) // 256 * 256) + ((P * 64 * 4 + 255) // 256 * 256)
        flags = ws[off:off + P * 4].view(torch.int32)
        msg += "  uncertified rows: %d of %d" % (int(flags.sum()), P)
        assert torch.equal(idx, ops.k_nn(f, 20))
    print(msg)
