"""One eagerly issued training step at BASELINE.json configs[1] between cudaProfilerStart / Stop, for
    ncu --profile-from-start off --set full --clock-control none -k regex:... -s S -c C python profiles/scripts/prof_step.py
(weight gradients on the main stream, no CUDA graph: every kernel is its own ncu range, in program order)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
os.environ["DGCNN_CUDA_GRAPH"] = "0"
os.environ["DGCNN_ASYNC_DW"] = "0"
import torch  # noqa: E402
import bench  # noqa: E402
import dgcnn  # noqa: E402

cfg = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 1]
fl = bench.make_flags(1, cfg, sys.argv[2] if len(sys.argv) > 2 else None)
tr = dgcnn.trainval(fl)
tr.initialize()
g = torch.Generator().manual_seed(1234)
x = torch.rand((cfg["B"], cfg["N"], 3), generator=g).cuda()
y = torch.randint(0, 2, (cfg["B"], cfg["N"]), generator=g).cuda()


def step():
    tr.zero_gradients(None)
    tr.accum_gradient(None, [x], [y], sync=False)
    tr.apply_gradient(None)


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("prof_step done")
