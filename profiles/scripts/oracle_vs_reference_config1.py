"""The oracle against the reference's own code AT THE BENCHMARKED ARCHITECTURE (BASELINE.json configs[1]: 4 EdgeConv layers,
N = 2048, k = 20, FC 512 / 256; B clouds, default 4 -- BN statistics are per micro-batch, so B only sets the run time).
Runs /root/reference/dgcnn/ops.py + model.py unmodified through oracle/tf1_shim (tests/golden/make_reference_golden.py) in
fp32 and fp64, then the oracle on the reference's neighbour graphs.  CPU only, build container only (needs /root/reference).
usage: python profiles/scripts/oracle_vs_reference_config1.py [B]   ->  profiles/r02_oracle_vs_reference_config1.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from tests.golden import make_reference_golden as g
from oracle import dgcnn_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N, K, L = 2048, 20, 4
torch.set_num_threads(os.cpu_count())
rng = np.random.RandomState(77)
x = rng.random_sample((B, N, 3)).astype(np.float32)
y = rng.randint(0, 2, (B, N)).astype(np.int64)
flags = g._flags(EDGE_CONV_LAYERS=L, KVALUE=K, FC_FILTERS=[512, 256], EDGE_CONV_FILTERS=64)
t0 = time.time()
z = g.run_model_both(flags, x, y, seed=9)
print("reference code through the shim (fp32 + fp64): %.1f s, B=%d N=%d k=%d L=%d FC 512/256" % (time.time() - t0, B, N, K, L))
fl = O.make_flags(EDGE_CONV_LAYERS=L, EDGE_CONV_FILTERS=64, KVALUE=K, FC_LAYERS=2, FC_FILTERS=[512, 256], NUM_CLASS=2,
                  MODEL_NAME="dgcnn", TRAIN=True, NUM_CHANNEL=3)
P = {k[len("param:"):]: torch.from_numpy(v) for k, v in z.items() if k.startswith("param:")}
xt = torch.from_numpy(x)
inputs = [xt] + [torch.from_numpy(z["tensor%d" % (3 * i + 2)]).squeeze(-2) for i in range(L - 1)]
for i in range(L):
    mine = O.k_nn(inputs[i], K).numpy()
    diff = mine != z["idx%d" % i]
    print("layer %d: oracle k_nn on the reference's layer input: %d of %d indices differ (%.4f %%)" % (
        i, int(diff.sum()), diff.size, 100.0 * diff.mean()))
mask = torch.from_numpy(z["dropout_mask"])
for name, dt, lg_key, gr_key, ix_key in (("fp32", torch.float32, "logits", "grad:", "idx%d"), ("fp64", torch.float64, "logits64", "grad64:", "knn64_%d")):
    Pd = {n: t.detach().to(dt).requires_grad_(True) for n, t in P.items()}
    tensors = []
    lg = O.build(xt.to(dt), fl, Pd, idx_list=[torch.from_numpy(z[ix_key % i]) for i in range(L)], dropout_mask=mask.to(dt),
                 tensors_out=tensors)
    _, acc, loss = O.softmax_loss_accuracy(lg, torch.from_numpy(y))
    loss.backward()
    ref_lg = torch.from_numpy(z[lg_key]).to(dt)
    print("%s: logits max |oracle - reference| %.3g; loss %.9f vs %.9f" % (
        name, float((lg.detach() - ref_lg).abs().max()), float(loss.detach()), float(z["loss" if name == "fp32" else "loss64"])))
    if name == "fp32":
        worst = max(float((t.detach() - torch.from_numpy(z["tensor%d" % i])).abs().max()) for i, t in enumerate(tensors))
        print("fp32: worst EdgeConv tensor max |diff| %.3g over %d tensors" % (worst, len(tensors)))
    rels = []
    for n, t in Pd.items():
        ref = torch.from_numpy(z[gr_key + n]).to(dt)
        rels.append((float((t.grad - ref).norm()) / max(float(ref.norm()), 1e-30), n))
    print("%s: parameter gradients, relative L2 |oracle - reference|: worst %.3g (%s), median %.3g" % (
        name, max(rels)[0], max(rels)[1], float(np.median([r for r, _ in rels]))))
