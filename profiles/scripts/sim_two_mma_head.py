"""CPU experiment (torch-CPU oracle, no GPU): what would 2 tensor-core MMAs per product instead of 3 cost in parity?
The head GEMMs (MergedEdgeConv / FC0 / FC1) run hi.hi + hi.lo + lo.hi on bf16 pairs (operands exact to ~2^-16).  Two MMAs
means one operand as a SINGLE 16-bit value: `w16` = activations as fp16 hi + lo, weights as one fp16 (2^-12 relative);
`a16` = the reverse; `bf3` = the shipped three-product scheme.  Same neighbour graph on all sides, configs[1] shape with B
clouds (usage: python profiles/scripts/sim_two_mma_head.py [B]).  Result at B = 4 (profiles/r02_two_mma_simulation.txt):
max |logit - fp32 oracle| 2.9e-3 (w16) / 5.8e-3 (a16) against north_star's 1e-3, gradients 10-40x further from fp64 than
fp32 arithmetic is -- the two-MMA head fails the parity gate and was not built."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import dgcnn_oracle as O
torch.set_num_threads(16)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
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
HEAD = ("MergedEdgeConv", "FC0", "FC1")

def split2(a, fmt):
    hi = a.to(fmt).to(a.dtype); lo = (a - hi).to(fmt).to(a.dtype)
    return hi + lo

class MM(torch.autograd.Function):
    mode = None
    @staticmethod
    def forward(ctx, a, w):
        m = MM.mode
        if m == "w16":      # a: fp16 hi+lo, w: single fp16
            a2 = split2(a.float(), torch.float16).to(a.dtype); w2 = w.float().half().to(w.dtype)
        elif m == "a16":
            a2 = a.float().half().to(a.dtype); w2 = split2(w.float(), torch.float16).to(w.dtype)
        elif m == "bf3":
            a2 = split2(a.float(), torch.bfloat16).to(a.dtype); w2 = split2(w.float(), torch.bfloat16).to(w.dtype)
        ctx.save_for_backward(a, w2)
        return a2 @ w2
    @staticmethod
    def backward(ctx, gr):
        a, w2 = ctx.saved_tensors
        return gr @ w2.t(), a.reshape(-1, a.shape[-1]).t() @ gr.reshape(-1, gr.shape[-1])

orig = O.conv_bn
def conv_bn(t, P_, scope, relu=True):
    if MM.mode and scope in HEAD:
        z = MM.apply(t, P_[scope + "/weights"])
        yv = O.bn_train(z, P_[scope + "/BatchNorm/beta"])
        return torch.relu(yv) if relu else yv
    return orig(t, P_, scope, relu)
O.conv_bn = conv_bn

idx = []
with torch.no_grad():
    O.build(x, fl, P, idx_out=idx, dropout_mask=mask)
res = {}
for name, dt, mode in (("f64", torch.float64, None), ("f32", torch.float32, None), ("bf3", torch.float32, "bf3"),
                       ("w16", torch.float32, "w16"), ("a16", torch.float32, "a16")):
    t0 = time.time()
    MM.mode = mode
    Pd = {n: t.detach().to(dt).requires_grad_(True) for n, t in P.items()}
    lg = O.build(x.to(dt), fl, Pd, idx_list=idx, dropout_mask=mask.to(dt))
    _, _, ls = O.softmax_loss_accuracy(lg, y); ls.backward()
    res[name] = ({n: t.grad.double() for n, t in Pd.items()}, lg.detach().double(), float(ls))
    print(name, "done %.1fs" % (time.time() - t0), flush=True)
g64, l64, _ = res["f64"]; g32, l32, _ = res["f32"]
for name in ("bf3", "w16", "a16"):
    gg, ll, _ = res[name]
    print("%s: logits |x-f32| max %.3g mean %.3g ; |x-f64| max %.3g ; |f32-f64| max %.3g" % (
        name, (ll - l32).abs().max(), (ll - l32).abs().mean(), (ll - l64).abs().max(), (l32 - l64).abs().max()))
    for n in P:
        den = float(g64[n].norm())
        print("   %-36s rel L2 vs fp64: %s %.3g   f32 %.3g" % (n, name, float((gg[n] - g64[n]).norm()) / den, float((g32[n] - g64[n]).norm()) / den))
