import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
import bench, dgcnn
fl = bench.make_flags(1)
tr = dgcnn.trainval(fl); tr.initialize()
g = torch.Generator().manual_seed(1234)
x = [torch.rand((24, 2048, 3), generator=g).cuda() for _ in range(4)]
y = [torch.randint(0, 2, (24, 2048), generator=g).cuda() for _ in range(4)]
def step(i, sync):
    tr.zero_gradients(None)
    r = tr.accum_gradient(None, [x[i % 4]], [y[i % 4]], sync=False)
    tr.apply_gradient(None)
    if sync: torch.cuda.synchronize()
for i in range(10): step(i, False)
for mode in (False, True, False, True):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): step(i, mode)
    e1.record(); torch.cuda.synchronize()
    print("sync each step" if mode else "free running ", e0.elapsed_time(e1) / 100, "ms/step")
