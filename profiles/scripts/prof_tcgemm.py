"""Times dgcnn_tc_gemm on the head's GEMM shapes against torch (cuBLAS) fp32 / tf32 / bf16."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))
from dgcnn import _native as nv
L = nv.lib()
dev = torch.device("cuda")

def split(x):
    p = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=dev)
    nv.check(L.dgcnn_split_bf16(x.data_ptr(), x.shape[0], x.shape[1], x.shape[1], p.data_ptr(), x.shape[1], x.numel(), nv.stream_ptr(dev)), "split")
    return p

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

P = 24 * 2048
for name, (M, N, K, tA, tB) in {"FC0 fwd": (P, 512, 1792, 0, 0), "FC0 dX": (P, 1792, 512, 0, 1), "FC0 dW": (1792, 512, P, 1, 0),
                                "Merged fwd": (P, 1024, 256, 0, 0), "Merged dX": (P, 256, 1024, 0, 1), "Merged dW": (256, 1024, P, 1, 0),
                                "FC1 fwd": (P, 256, 512, 0, 0), "FC1 dX": (P, 512, 256, 0, 1), "FC1 dW": (512, 256, P, 1, 0),
                                "uv fwd": (P, 128, 64, 0, 0), "uv dW": (64, 128, P, 1, 0)}.items():
    A = torch.randn((K, M) if tA else (M, K), device=dev)
    B = torch.randn((N, K) if tB else (K, N), device=dev)
    pa, pb = split(A), split(B)
    out = torch.empty((M, N), device=dev)
    ws = torch.empty(max(L.dgcnn_tc_gemm_workspace_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)
    f = lambda: nv.check(L.dgcnn_tc_gemm(pa.data_ptr(), pb.data_ptr(), out.data_ptr(), M, N, K, tA, tB, 2, ws.data_ptr(), ws.numel(), nv.stream_ptr(dev)), "g")
    t = timeit(f)
    ts = timeit(lambda: split(A))
    opA = A.t() if tA else A
    opB = B.t() if tB else B
    torch.backends.cuda.matmul.allow_tf32 = False
    t32 = timeit(lambda: opA @ opB)
    torch.backends.cuda.matmul.allow_tf32 = True
    ttf = timeit(lambda: opA @ opB)
    torch.backends.cuda.matmul.allow_tf32 = False
    a16, b16 = opA.bfloat16(), opB.bfloat16()
    t16 = timeit(lambda: a16 @ b16)
    fl = 2.0 * M * N * K
    print("%-11s M=%6d N=%5d K=%6d | tc_gemm %.3f ms (%.0f TF/s fp32-equiv, %.0f bf16 TF/s) split(A) %.3f ms | cublas fp32 %.3f tf32 %.3f bf16 %.3f ms"
          % (name, M, N, K, t, fl / t / 1e9, 3 * fl / t / 1e9, ts, t32, ttf, t16))
