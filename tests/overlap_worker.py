"""Worker of tests/test_gpu_overlap.py (not collected by pytest): the trainer's two-bucket gradient all-reduce that runs
inside the last micro-step (parallel.GradBuckets, captured in the CUDA graph) against the plain single all-reduce in
apply_gradient().  Runs as ONE process on a 1-rank NCCL group (DGCNN_OVERLAP_AR=force) or under torchrun on N GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/overlap_worker.py

Every rank gets its own micro-batch; the all-reduced flat gradient buffer (head + tail + loss / accuracy slots) of the
overlapped path must equal the single all-reduce's up to the run-to-run noise of the fp32 atomics in the EdgeConv backward.
/root/reference/dgcnn/trainval.py:59-80."""
import os
import socket
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dynamic-gcnn_b200"))

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", rank=rank, world_size=world)
    import dgcnn as dg
    from dgcnn import model as M
    B, N = 2, 512
    dev = torch.device("cuda", torch.cuda.current_device())
    orig = M.build
    mask = torch.ones((B, N, 1, 256), device=dev)
    M.build = lambda pc, fl, dropout_mask=None: orig(pc, fl, dropout_mask=mask * M.DROPOUT_KEEP)   # no dropout noise
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand((B, N, 3), generator=g)
    y = torch.randint(0, 2, (B, N), generator=g)

    def run(mode, reps):
        os.environ["DGCNN_OVERLAP_AR"] = mode
        fl = SimpleNamespace(NUM_CLASS=2, MODEL_NAME="dgcnn", TRAIN=True, KVALUE=8, DEBUG=False, EDGE_CONV_LAYERS=2,
                             EDGE_CONV_FILTERS=64, FC_LAYERS=2, FC_FILTERS=[256, 256], LEARNING_RATE=0.0,
                             GPUS=list(range(world)), MINIBATCH_SIZE=B, NUM_CHANNEL=3, WEIGHT_KEY="", SEED=0,
                             BATCH_SIZE=B * world, NUM_POINT=N)
        tr = dg.trainval(fl)
        tr.initialize()
        data = [x] * world           # entry i is tower i's slice; this rank only looks at its own
        label = [y] * world
        out = []
        for _ in range(reps):
            tr.zero_gradients(None)
            tr.accum_gradient(None, data, label, sync=False, last=True)
            tr.apply_gradient(None)                       # learning rate 0: the buffer still holds the reduced sums
            out.append(tr.variables.flat_grad.clone())
        torch.cuda.synchronize()
        return out, tr

    ref, tr0 = run("0", 1)
    assert tr0._buckets is None
    ovl, tr1 = run("force", 4)                            # 2 eager + capture / replay + replay
    assert tr1._buckets is not None and len(tr1._head) == 8 and len(tr1._graphs) == 1
    scale = ref[0].abs().max().item()
    for i, t in enumerate(ovl):
        err = (t - ref[0]).abs().max().item()
        assert err <= 2e-4 * scale, (i, err, scale)
        assert abs(t[-2].item() - ref[0][-2].item()) <= 1e-6 and abs(t[-1].item() - ref[0][-1].item()) <= 1e-6
    # every rank holds the same sums
    if world > 1:
        mine = ovl[-1].clone()
        dist.broadcast(mine, src=0)
        assert torch.equal(mine, ovl[-1])
    # the promise of last=True is enforced
    tr1.zero_gradients(None)
    tr1.accum_gradient(None, [x] * world, [y] * world, sync=False, last=True)
    try:
        tr1.accum_gradient(None, [x] * world, [y] * world, sync=False)
        raise AssertionError("second accum_gradient after last=True must raise")
    except RuntimeError:
        pass
    tr1.apply_gradient(None)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("OVERLAP_OK world=%d max|grad| %.3g" % (world, scale), flush=True)
    tr1.release_graphs()          # graphs with NCCL nodes keep the communicator referenced: they go before the group
    tr0.release_graphs()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
