"""Runs one training step at the shapes of BASELINE.json configs[2] (residual, N=4096, k=40, L=6) and configs[4]
(N=16384, k=20, B=8) -- fp32 path -- and reports ms/step and peak memory."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200")); sys.path.insert(0, ROOT)
import bench, dgcnn
SEL = sys.argv[1] if len(sys.argv) > 1 else ""
for name, kw, B, N in (("configs[2] residual-dgcnn N=4096 k=40 L=6 (fp32)", dict(MODEL_NAME="residual-dgcnn", KVALUE=40, EDGE_CONV_LAYERS=6), 24, 4096),
                       ("configs[4] dgcnn N=16384 k=20 L=4", dict(), 8, 16384)):
    if SEL and SEL not in name:
        continue
    fl = bench.make_flags(1)
    for k, v in kw.items(): setattr(fl, k, v)
    fl.MINIBATCH_SIZE = fl.BATCH_SIZE = B; fl.NUM_POINT = N
    tr = dgcnn.trainval(fl); tr.initialize()
    g = torch.Generator().manual_seed(1)
    x = torch.rand((B, N, 3), generator=g).cuda(); y = torch.randint(0, 2, (B, N), generator=g).cuda()
    torch.cuda.reset_peak_memory_stats()
    for i in range(9):   # 2 eager micro-steps, capture on the 3rd, one replay, then 5 timed replays
        if i == 4:
            torch.cuda.synchronize(); t0 = time.time()
        tr.zero_gradients(None); r = tr.accum_gradient(None, [x], [y], sync=False); tr.apply_gradient(None)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 5
    print("%s: %.2f ms/step, %.2f M points/s, loss %.4f, peak mem %.1f GB" % (name, dt * 1e3, B * N / dt / 1e6, float(r[2]), torch.cuda.max_memory_allocated() / 2**30))
    del tr
