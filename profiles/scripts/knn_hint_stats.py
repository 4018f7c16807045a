"""How tight is the previous layer's graph as a warm start?  For each 64-channel layer input of the random-init
DGCNN at configs[1]: per-row count of columns whose distance is <= U = max distance to the k hinted neighbours,
(and <= the true k-th distance, = k + ties), as quantiles."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200")); sys.path.insert(0, ROOT)
import bench, dgcnn
from dgcnn import ops
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
qs = torch.tensor([0.1, 0.5, 0.9, 0.99, 0.999, 1.0], device="cuda")
for li, (f, h) in enumerate(feats):
    if h is None:
        continue
    B, N, C = f.shape
    s = (f * f).sum(-1)
    D = s[:, :, None] + s[:, None, :] - 2 * torch.bmm(f, f.transpose(1, 2))
    dh = torch.gather(D, 2, h.long())
    U = dh.max(dim=2, keepdim=True).values
    cnt = (D <= U).sum(-1).float().flatten()
    eps = (s[:, :, None] + s[:, None, :]) / 2048.0
    cnt_lo = ((D - eps) <= U).sum(-1).float().flatten()
    kth = D.kthvalue(20, dim=2, keepdim=True).values
    cnt_k = ((D - eps) <= kth).sum(-1).float().flatten()
    print("layer %d: s mean %.3f  kth-dist mean %.4f  U mean %.4f  eps mean %.5f" % (li, s.mean(), kth.mean(), U.mean(), eps.mean()))
    print("   #(D<=U)        q", torch.quantile(cnt, qs).tolist(), "mean", cnt.mean().item())
    print("   #(D-eps<=U)    q", torch.quantile(cnt_lo, qs).tolist(), "mean", cnt_lo.mean().item())
    print("   #(D-eps<=kth)  q", torch.quantile(cnt_k, qs).tolist(), "mean", cnt_k.mean().item())
    # overlap of the new graph with the hint
    new = ops.k_nn(f, 20)
    ov = (new[:, :, :, None] == h[:, :, None, :]).any(-1).float().sum(-1).mean().item()
    print("   overlap with hint: %.2f of 20" % ov)
