"""Phase timeline of knn_tc_filter_kernel (library built with EXTRA=-DK2_DEBUG): per-CTA globaltimer stamps."""
import os, sys, ctypes, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
import dgcnn
from dgcnn import _native as nv
g = torch.Generator().manual_seed(1234)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.rand((24, 2048, C), generator=g).cuda()
for _ in range(3): dgcnn.ops.k_nn(x, 20)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (512 * 16))()
assert nv.lib().dgcnn_knn_debug_stamps(buf) == 0
a = np.array(buf, dtype=np.int64).reshape(512, 16)[:384]
t0 = a[:, 0].min()
names = ["start", "setup done", "scan: first acc", "sweep1 done", "bisect done", "sweep2 done", "emit done", "exit",
         "prod: sweep1 issued", "prod: done", "mma: A ready", "mma: sweep2 start", "mma: done", "mma: step1"]
rel = a - a[:, :1]
print("kernel span %.1f us; CTA duration mean %.1f us (min %.1f max %.1f)" % ((a[:, 7].max() - t0) / 1e3, rel[:, 7].mean() / 1e3, rel[:, 7].min() / 1e3, rel[:, 7].max() / 1e3))
for i, n in enumerate(names):
    print("%-22s mean %8.2f us   min %8.2f  max %8.2f" % (n, rel[:, i].mean() / 1e3, rel[:, i].min() / 1e3, rel[:, i].max() / 1e3))
starts = np.sort((a[:, 0] - t0) / 1e3)
print("CTA start times (us): first wave max %.1f; quantiles" % starts[147], np.quantile(starts, [0.4, 0.5, 0.75, 0.9, 1.0]))
