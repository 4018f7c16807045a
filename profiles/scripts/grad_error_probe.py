"""Who is wrong on the gradients at configs[1]?  GPU (fp32 + bf16x3 tensor cores) and the fp32 oracle are both compared
with the fp64 oracle on the same neighbour graph.  usage: python profiles/scripts/grad_error_probe.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
import torch
import dgcnn as dg
from oracle import dgcnn_oracle as O
from dgcnn.variables import set_default_store

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
N, K, L = 2048, 20, 4
fl = O.make_flags(EDGE_CONV_LAYERS=L, EDGE_CONV_FILTERS=64, KVALUE=K, FC_LAYERS=2, FC_FILTERS=[512, 256], NUM_CLASS=2,
                  MODEL_NAME="dgcnn", TRAIN=True, NUM_CHANNEL=3, MINIBATCH_SIZE=B)
P = O.init_params(fl, 3, seed=0)
g = torch.Generator().manual_seed(77)
for n, t in P.items():
    if n.endswith("beta"):
        t.copy_(0.1 * torch.randn(t.shape, generator=g))
x = torch.rand((B, N, 3), generator=g); y = torch.randint(0, 2, (B, N), generator=g)
mask = (torch.rand((B, N, 1, 256), generator=g) < 0.7).float()
tr = dg.trainval(fl); tr.initialize()
tr.variables.load_state_dict({"dgcnn/" + n: t for n, t in P.items()})
dg.ops._knn_trace = []
tr.zero_gradients(None)
old = set_default_store(tr.variables)
with tr.variables.variable_scope("dgcnn"):
    logits = dg.build(x.cuda(), fl, dropout_mask=mask.cuda())
set_default_store(old)
loss = torch.nn.functional.cross_entropy(logits.reshape(-1, 2), y.cuda().reshape(-1)); loss.backward()
trace = [t.cpu() for t in dg.ops._knn_trace]; dg.ops._knn_trace = None
res = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    Pd = {n: t.detach().to(dt).requires_grad_(True) for n, t in P.items()}
    lg = O.build(x.to(dt), fl, Pd, idx_list=trace, dropout_mask=mask.to(dt))
    _, _, ls = O.softmax_loss_accuracy(lg, y); ls.backward()
    res[name] = ({n: t.grad.double() for n, t in Pd.items()}, lg.detach().double(), float(ls))
g64, l64, loss64 = res["f64"]; g32, l32, loss32 = res["f32"]
print("logits: |gpu-f64| max %.3g  |f32-f64| max %.3g ; loss gpu %.7f f32 %.7f f64 %.7f" % (
    (logits.detach().cpu().double() - l64).abs().max(), (l32 - l64).abs().max(), loss.item(), loss32, loss64))
for n in P:
    a = tr.variables.vars["dgcnn/" + n].grad.cpu().double()
    den = float(g64[n].norm())
    print("%-36s rel L2 vs fp64: gpu %.3g   oracle-fp32 %.3g" % (n, float((a - g64[n]).norm()) / den, float((g32[n] - g64[n]).norm()) / den))
