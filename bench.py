#!/usr/bin/env python
"""bench.py -- points/sec, forward+backward+Adam, of the full DGCNN segmentation model on synthetic clouds.

Workload (BASELINE.json configs[1]): 4 EdgeConv layers + FC head, N=2048, k=20, C=3, 24 clouds per GPU, fp32.
With --gpus N (launched under torchrun) every rank keeps 24 clouds (weak scaling: global batch 24*N, which at
N=8 is configs[3]) and the step ends with ONE NCCL all-reduce of the flat gradient buffer.

A "step" = zero_gradients + accum_gradient (fwd+bwd of the rank's 24 clouds) + apply_gradient through the
reference-facing trainer API (dgcnn.trainval); the trainer replays the micro-step from a CUDA graph.  Two timings:
  value : inputs already resident in HBM, no host read inside the loop; CUDA events, max over ranks.
  e2e   : the same call with pinned HOST inputs (H2D copy every step) and a device->host read of the loss
          every step.
roofline    : the largest single launch of the step, tc_gemm_wide_kernel on the FC0 layer (tensor bound), timed live
              with CUDA events on the launching stream in three extra eagerly-issued steps (a captured graph cannot
              carry timing events); the fused k_nn and the EdgeConv gather passes are reported beside it.
cpu_baseline: the oracle (reference-equivalent CPU restatement; TF1 cannot be installed) on a bounded sample.
--impl reference : that CPU arm alone, with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dynamic-gcnn_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B_PER_GPU, NPTS, KNN, CH, LAYERS, FILT, FCF, NCLS = 24, 2048, 20, 3, 4, 64, [512, 256], 2
METRIC = "points/sec fwd+bwd, B=24 N=2048 k=20, at 1/2/4/8 B200; EdgeConv HBM GB/s"
CPU_SAMPLE_B = 4  # clouds per CPU-baseline step (BN statistics are per micro-batch, cost is linear in clouds)


def make_flags(n_gpus):
    from types import SimpleNamespace
    return SimpleNamespace(NUM_CLASS=NCLS, MODEL_NAME="dgcnn", TRAIN=True, KVALUE=KNN, DEBUG=False,
                           EDGE_CONV_LAYERS=LAYERS, EDGE_CONV_FILTERS=FILT, FC_LAYERS=len(FCF), FC_FILTERS=list(FCF),
                           LEARNING_RATE=1e-3, GPUS=list(range(n_gpus)), MINIBATCH_SIZE=B_PER_GPU,
                           NUM_CHANNEL=CH, WEIGHT_KEY="", SEED=0, BATCH_SIZE=B_PER_GPU * n_gpus, NUM_POINT=NPTS)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons DURING the timed region (pynvml; nvidia-smi's clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline_run(steps, warmup):
    """Reference-equivalent CPU restatement (oracle) timed on the host cores: TF-literal graph (materialised
    [B,N,N] distances, top_k, gathered edge tensor, 1x1 convs as matmuls, train-mode BN, autograd backward)."""
    from oracle import dgcnn_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fl = O.make_flags(EDGE_CONV_LAYERS=LAYERS, EDGE_CONV_FILTERS=FILT, KVALUE=KNN, FC_FILTERS=list(FCF), NUM_CLASS=NCLS,
                      TRAIN=True)
    P = O.init_params(fl, CH, seed=0)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand((CPU_SAMPLE_B, NPTS, CH), generator=g)
    y = torch.randint(0, NCLS, (CPU_SAMPLE_B, NPTS), generator=g)
    for _ in range(warmup):
        O.train_step_reference(x, y, fl, P)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step_reference(x, y, fl, P)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": CPU_SAMPLE_B * NPTS / dt, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "%d of %d clouds per step (N=%d k=%d L=%d, fwd+bwd, no optimizer), %d steps after %d warm-up; "
                      "torch-CPU restatement of the TF1 graph (TF1 not installable)" % (CPU_SAMPLE_B, B_PER_GPU, NPTS,
                                                                                       KNN, LAYERS, steps, warmup),
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline_run(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "configs[1]: full DGCNN seg, 4 EdgeConv + FC(512,256) head, N=2048 k=20 C=3, 24 clouds per GPU, "
                        "fwd+bwd+Adam" + ("" if n == 1 else "; %d GPUs = global batch %d, one NCCL grad all-reduce" %
                                          (n, n * B_PER_GPU)),
            "clouds_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * n, "points": NPTS, "k": KNN, "channels": CH,
            "edge_conv_layers": LAYERS, "parallelism": "dp%d" % n,
            "l2": "no explicit flush: one step streams >2 GB of activations per GPU, far beyond the 126 MB L2",
            "launch": "micro-step (fwd+bwd) replayed from a CUDA graph captured by dgcnn.trainval after 2 eager runs"}


def edgeconv_entry(eev):
    """EdgeConv forward after the uv GEMM (ops.py:45-58: gather, conv0 as u_i + v_j, BN, ReLU, max_k / mean_k, concat) =
    ec_fwd_stats_kernel + finalize + ec_fwd_apply_kernel, on the 64-channel layers."""
    ts = [a.elapsed_time(b) for (_, _, f, _, a, b) in eev]
    if not ts:
        return None
    t = float(np.mean(ts)) * 1e-3
    P_, E_ = B_PER_GPU * NPTS, B_PER_GPU * NPTS * KNN
    hbm = (P_ * 128 + E_ + 3 * P_ * 64 + P_ * 128) * 4.0        # uv, idx, zmax/cnt, (max | mean): compulsory bytes
    l2 = 2.0 * E_ * 64 * 4                                       # two gather passes over the L2-resident v rows
    ref = (E_ * 128 * 4.0) * 2 + (E_ * 64 * 4.0) * 5             # the reference's edge tensor w+r, conv0 out w + BN r/w + max/mean r
    return {"ms_per_call": t * 1e3, "calls_timed": len(ts), "bound": "L2 gather (E*F*4 bytes per pass) + fp32 ALU",
            "compulsory_hbm_gbs": hbm / t / 1e9, "l2_gather_gbs": l2 / t / 1e9,
            "effective_unfused_hbm_gbs": ref / t / 1e9,
            "note": "effective = bytes the reference's op-by-op graph moves for the same layer (edge tensor, conv0 output, "
                    "BN, max/mean passes) divided by this time; none of those tensors exists here"}


def run_ours(args):
    import torch.distributed as dist
    import dgcnn
    from dgcnn import _native, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "--gpus %d but WORLD_SIZE=%d (launch with torchrun)" % (args.gpus, world)

    fl = make_flags(world)
    tr = dgcnn.trainval(fl)
    tr.initialize()
    tower = tr._towers[0]

    # synthetic inputs: a small ring of distinct batches, pinned on the host and mirrored on the device
    g = torch.Generator().manual_seed(1234 + rank)
    ring = 4
    h_pts = [torch.rand((B_PER_GPU, NPTS, CH), generator=g).pin_memory() for _ in range(ring)]
    h_lab = [torch.randint(0, NCLS, (B_PER_GPU, NPTS), generator=g, dtype=torch.int64).pin_memory() for _ in range(ring)]
    d_pts = [t.to(dev) for t in h_pts]
    d_lab = [t.to(dev) for t in h_lab]

    def feed(lst, i):
        out = [None] * world
        out[tower] = lst[i % ring]
        return out

    def step(i, host):
        tr.zero_gradients(None)
        res = tr.accum_gradient(None, feed(h_pts if host else d_pts, i), feed(h_lab if host else d_lab, i), sync=host)
        tr.apply_gradient(None)
        return res[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        n0 = _native.launch_count()
        prof = (not host) and os.environ.get("DGCNN_PROFILE") == "1"   # ncu --profile-from-start off
        if prof:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(steps):
            step(i, host)
        e1.record()
        barrier()
        if prof:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _native.launch_count() - n0

    # warm-up: the trainer runs its first micro-steps eagerly, then captures the micro-step into a CUDA graph
    for i in range(max(args.warmup, 3)):
        step(i, False)
    for i in range(min(args.warmup, 3)):
        step(i, True)

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches = timed(False, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_e2e, _ = timed(True, args.steps)

    # per-kernel timing for the roofline entry: the same steps issued eagerly (a captured graph cannot carry timing
    # events), the kernels of interest bracketed by CUDA events on the launching stream
    os.environ["DGCNN_CUDA_GRAPH"] = "0"
    ops._knn_events, ops._gemm_events, ops._ec_events = [], [], []
    barrier()
    for i in range(3):
        step(i, False)
    barrier()
    ev, ops._knn_events = ops._knn_events, None
    gev, ops._gemm_events = ops._gemm_events, None
    eev, ops._ec_events = ops._ec_events, None
    os.environ.pop("DGCNN_CUDA_GRAPH", None)

    pts_per_step = B_PER_GPU * NPTS * world
    value = pts_per_step * args.steps / (ms_dev * 1e-3)
    e2e = pts_per_step * args.steps / (ms_e2e * 1e-3)

    # Roofline of the dominant kernel: tc_gemm_wide_kernel on the FC0 layer (ops.py:151-160: the 1x1 conv over the
    # 1792 non-broadcast channels of model.py:83-85's concat -> 512), the largest single launch of the step.
    pk, pk_kind = peaks()
    P_ = B_PER_GPU * NPTS
    fc0 = [a.elapsed_time(b) for (m, n, k, a, b) in gev if n == FCF[0] and m == P_]
    kfc0 = max([k for (m, n, k, a, b) in gev if n == FCF[0] and m == P_], default=0)
    knn64 = [a.elapsed_time(b) for (_, _, c, _, a, b) in ev if c == 64]
    knn3 = [a.elapsed_time(b) for (_, _, c, _, a, b) in ev if c == CH]
    roof = None
    if fc0:
        t_ms = float(np.mean(fc0))
        flops = 2.0 * P_ * FCF[0] * kfc0                     # algorithmic: one fp32 multiply-add per (row, col, k)
        ach = flops / (t_ms * 1e-3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        roof = {"kernel": "tc_gemm_wide_kernel<A K-major, B MN-major> on FC0 forward: [%d x %d] . [%d x %d], persistent "
                          "128x256 tiles, TMA -> 2-stage smem ring -> tcgen05.mma (3 bf16 MMAs per k-slice: hi.hi + hi.lo + "
                          "lo.hi, fp32 accumulate in TMEM, double-buffered) -> epilogue with BN column statistics"
                          % (P_, kfc0, kfc0, FCF[0]),
                "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": 438.2e6, "traffic_source": "profiles/r01e_top_kernels_ncu_full.csv: ncu --set full of this launch, "
                                                       "dram read 362.8 MB + write 75.4 MB (algorithmic: bf16 hi/lo operand "
                                                       "planes 352 MB + weights 3.7 MB + fp32 output 100.7 MB)",
                "peak_source": pk_kind + " bf16 sustained (kernel timed inside the step)",
                "ms_per_launch": t_ms, "launches_timed": len(fc0),
                "algorithmic_flops": flops, "executed_tensor_flops": 3 * flops,
                "executed_frac_of_peak": 3 * ach / peak,
                "note": "achieved counts ONE multiply-add per product (the fp32 GEMM the reference runs); the kernel "
                        "executes three bf16 MMAs per product to stay within the 1e-3 logits bound, so the tensor pipe "
                        "itself runs at executed_frac_of_peak of the measured cuBLAS bf16 rate",
                "other_kernels": {
                    "k_nn_fused_64ch": {
                        "what": "dgcnn_knn on [24,2048,64]: range + prep + knn_tc_filter_kernel (fp16 tcgen05 pass, two "
                                "sweeps over TMEM accumulators) + exact refine; the [B,N,N] matrix never leaves the SM",
                        "ms_per_call": float(np.mean(knn64)) if knn64 else None,
                        "bound": "TMEM read port (64 B/clk/SM): 2 sweeps x 402.7 MB of fp32 accumulators",
                        "effective_unfused_hbm_gbs": (821.9e6 / (float(np.mean(knn64)) * 1e-3) / 1e9) if knn64 else None,
                        "algorithmic_tflops": (2.0 * B_PER_GPU * NPTS * NPTS * 64 / (float(np.mean(knn64)) * 1e-3) / 1e12)
                        if knn64 else None},
                    "k_nn_fused_xyz": {"ms_per_call": float(np.mean(knn3)) if knn3 else None},
                    "edgeconv_fwd_gather": edgeconv_entry(eev),
                    "knn_share_of_step": (sum(knn64) + sum(knn3)) / 3.0 / (ms_dev / args.steps)}}

    if rank == 0:
        cb = cpu_baseline_run(3, 1) if (world == 1 and not args.no_cpu_baseline) else None
        line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": e2e, "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": int(h_pts[0].numel() * 4 + h_lab[0].numel() * 8),
                        "d2h_bytes_per_step": 8},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof}
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
