"""Host-side logic that needs no GPU: flags, IO handlers, variable scopes, argument validation of the ops
mirror, tower assignment, and the one-all-reduce data-parallel rule on a 2-rank gloo group."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_flags_defaults_and_parsing(dg):
    f = dg.DGCNN_FLAGS()
    assert (f.KVALUE, f.EDGE_CONV_LAYERS, f.NUM_CLASS, f.FC_FILTERS, f.MINIBATCH_SIZE) == (20, 3, 2, "512,256", 1)
    a = f.parser.parse_args("train -io synthetic -bs 4 -mbs 2 -ecl 1 -kv 16 -np 512 -it 2 --gpus 0,1 -sd 3".split())
    f.update({k: v for k, v in vars(a).items()})
    assert f.GPUS == [0, 1] and f.FC_FILTERS == [512, 256] and f.EDGE_CONV_FILTERS == 64
    assert (f.BATCH_SIZE, f.MINIBATCH_SIZE, f.KVALUE, f.NUM_POINT, f.SEED, f.IO_TYPE) == (4, 2, 16, 512, 3, "synthetic")
    a = f.parser.parse_args("inference -ecf 32,64 -db 0".split())
    f.update(vars(a))
    assert f.EDGE_CONV_FILTERS == [32, 64] and f.DEBUG == 0


def test_io_synthetic_and_array(dg, tmp_path):
    from types import SimpleNamespace
    fl = SimpleNamespace(BATCH_SIZE=3, NUM_POINT=128, NUM_CHANNEL=-1, NUM_CLASS=2, LABEL_KEY="label", WEIGHT_KEY="",
                         OUTPUT_FILE="", SHUFFLE=0, IO_TYPE="synthetic", INPUT_FILE=[""], DATA_KEY="data")
    io = dg.io_factory(fl)
    io.initialize()
    idx, data, label, weight = io.next()
    assert data.shape == (3, 128, 3) and label.shape == (3, 128) and weight is None and idx.tolist() == [0, 1, 2]
    assert io.num_channels() == 3 and io.next()[0].tolist() == [3, 4, 5]
    # dense file source (the io_h5 contract) with output store
    path = str(tmp_path / "d.npz")
    np.savez(path, data=np.random.rand(5, 16, 4).astype(np.float32), label=np.zeros((5, 16), np.int32))
    fl.IO_TYPE, fl.INPUT_FILE, fl.OUTPUT_FILE, fl.BATCH_SIZE = "h5", [path], str(tmp_path / "o.npz"), 4
    io = dg.io_factory(fl)
    io.initialize()
    assert io.num_entries() == 5 and io.num_channels() == 4
    assert io.next()[0].tolist() == [0, 1, 2, 3] and io.next()[0].tolist() == [4, 0, 1, 2]  # sequential wraparound
    io.store(1, np.ones((16, 2), np.float32))
    with pytest.raises(ValueError):
        io.store(9, None)
    io.finalize()
    assert np.load(fl.OUTPUT_FILE)["softmax"].shape == (1, 16, 2)
    fl.IO_TYPE = "nope"
    with pytest.raises(NotImplementedError):
        dg.io_factory(fl)


def test_variable_store_scopes_and_reuse(dg):
    from dgcnn.variables import VariableStore
    st = VariableStore(device="cpu", seed=0)
    with st.variable_scope("dgcnn"), st.variable_scope("EdgeConv0"), st.variable_scope("conv0"):
        w = st.get_variable("weights", (6, 64), "xavier")
        assert st.get_variable("weights", (6, 64), "xavier") is w  # AUTO_REUSE
        with pytest.raises(ValueError):
            st.get_variable("weights", (6, 32), "xavier")
    assert list(st.vars) == ["dgcnn/EdgeConv0/conv0/weights"]
    assert w.abs().max() <= np.sqrt(6.0 / 70) and w.requires_grad
    st.flatten(extra=2)
    assert st.flat_grad.numel() == 6 * 64 + 2 and w.grad.data_ptr() == st.flat_grad.data_ptr()
    (w * 2).sum().backward()
    assert float(st.flat_grad[:384].sum()) == 768.0  # autograd accumulates straight into the flat bucket


def test_declared_variables_match_survey_inventory(dg, oracle):
    from dgcnn import model
    from dgcnn.variables import VariableStore, set_default_store
    fl = oracle.make_flags(EDGE_CONV_LAYERS=4)
    st = VariableStore(device="cpu")
    old = set_default_store(st)
    try:
        with st.variable_scope("dgcnn"):
            model.declare_variables(fl, 3, "cpu")
    finally:
        set_default_store(old)
    assert st.num_params() == 1895554
    assert {n[len("dgcnn/"):]: tuple(v.shape) for n, v in st.vars.items()} == oracle.param_shapes(fl, 3)


def test_ops_list_validation_raises_valueerror(dg):
    x = torch.zeros(1, 8, 3)
    with pytest.raises(ValueError):
        dg.ops.repeat_edge_conv(x, 2, [4], 64, True)
    with pytest.raises(ValueError):
        dg.ops.repeat_residual_edge_conv(x, 2, 4, [64], True)
    with pytest.raises(ValueError):
        dg.ops.fc(x, 2, [64], True)


def test_tower_assignment(dg):
    from dgcnn.parallel import tower_assignment as ta
    assert [ta(8, 8, r) for r in range(8)] == [[r] for r in range(8)]
    assert ta(4, 1, 0) == [0, 1, 2, 3]
    assert [ta(4, 2, r) for r in range(2)] == [[0, 1], [2, 3]]
    assert [ta(2, 4, r) for r in range(4)] == [[0], [], [1], []]
    with pytest.raises(ValueError):
        ta(3, 2, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dynamic-gcnn_b200"))
    from dgcnn.parallel import allreduce_flat_, tower_assignment
    from dgcnn.variables import VariableStore
    # a toy "model": loss_tower = mean((x_tower @ w)^2); 4 towers over 2 ranks, 2 micro-steps accumulated
    st = VariableStore(device="cpu", seed=0)
    w = st.get_variable("w", (3, 2), "xavier")
    st.flatten(extra=2)
    g = torch.Generator().manual_seed(5)
    xs = torch.randn(2, 4, 5, 3, generator=g)  # [micro, tower, rows, 3] identical on every rank
    towers = tower_assignment(4, world, rank)
    for micro in range(2):
        for t in towers:
            loss = ((xs[micro, t] @ w) ** 2).mean()
            (loss / 4.0).backward()
            st.flat_grad[-2] += loss.detach() / 4.0
    allreduce_flat_(st.flat_grad)
    q.put((rank, st.flat_grad.clone()))
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_tower_mean_summed_over_microsteps():
    """trainval.py:64-69 (mean over towers) + :79 (sum over micro-steps) == one all-reduce of the flat bucket."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    from dgcnn.variables import VariableStore
    st = VariableStore(device="cpu", seed=0)
    w = st.get_variable("w", (3, 2), "xavier")
    g = torch.Generator().manual_seed(5)
    xs = torch.randn(2, 4, 5, 3, generator=g)
    total = sum(torch.stack([((xs[m, t] @ w) ** 2).mean() for t in range(4)]).mean() for m in range(2))
    total.backward()
    for r in (0, 1):
        assert torch.allclose(got[r][:6], w.grad.reshape(-1), atol=1e-6)
        assert torch.allclose(got[r][-2], total.detach(), atol=1e-6)
    assert torch.equal(got[0], got[1])


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dynamic-gcnn_b200"))
    from dgcnn.parallel import GradBuckets, allreduce_flat_, head_split_offset
    from dgcnn.variables import VariableStore
    st = VariableStore(device="cpu", seed=0)
    with st.variable_scope("dgcnn"):
        with st.variable_scope("EdgeConv0"):
            with st.variable_scope("conv0"):
                w0 = st.get_variable("weights", (3, 4), "xavier")
        with st.variable_scope("FC0"):
            w1 = st.get_variable("weights", (4, 2), "xavier")
    st.flatten(extra=2)
    names = st.trainable_names()
    split = head_split_offset(names, [st.vars[n].numel() for n in names])
    g = torch.Generator().manual_seed(11 + rank)            # every rank its own micro-batch
    x = torch.randn(6, 3, generator=g)
    loss = (torch.relu(x @ w0) @ w1).pow(2).mean()
    # the trainer's sequence: loss slot, head gradient folded -> head bucket starts, tail folded -> tail bucket, wait
    views = [w0.grad, w1.grad]
    w0.grad = w1.grad = None
    loss.backward()
    st.flat_grad[-2:].add_(torch.stack([loss.detach(), torch.tensor(float(rank))]), alpha=0.5)
    bk = GradBuckets(st.flat_grad, split)
    views[1].add_(w1.grad, alpha=0.5)
    bk.reduce_head_async()
    views[0].add_(w0.grad, alpha=0.5)
    bk.reduce_tail()
    bk.wait()
    two = st.flat_grad.clone()
    # the single all-reduce of the same local sums
    st.flat_grad.zero_()
    views[0].add_(w0.grad, alpha=0.5)
    views[1].add_(w1.grad, alpha=0.5)
    st.flat_grad[-2:].add_(torch.stack([loss.detach(), torch.tensor(float(rank))]), alpha=0.5)
    allreduce_flat_(st.flat_grad)
    q.put((rank, split, two, st.flat_grad.clone()))
    dist.destroy_process_group()


def test_two_bucket_overlapped_allreduce_equals_single_allreduce():
    """parallel.GradBuckets (head bucket reduced while the rest of backward runs, tail bucket after) gives exactly the
    flat buffer a single all-reduce gives; head_split_offset finds the head as the tail of the declaration order."""
    from dgcnn.parallel import GradBuckets, head_split_offset
    names = ["dgcnn/EdgeConv0/conv0/weights", "dgcnn/EdgeConv0/conv0/BatchNorm/beta", "dgcnn/MergedEdgeConv/weights",
             "dgcnn/FC0/weights", "dgcnn/FC1/BatchNorm/beta", "dgcnn/Final/weights"]
    assert head_split_offset(names, [384, 64, 1000, 10, 5, 2]) == 448
    assert head_split_offset(names[:2], [384, 64]) is None                       # no head at all
    assert head_split_offset(names[2:], [1000, 10, 5, 2]) is None                # nothing below the head
    assert head_split_offset([names[2], names[0]], [1000, 384]) is None           # head is not a suffix
    assert head_split_offset(["dgcnn/EdgeConv1/conv1/weights", "dgcnn/Final/weights"], [8192, 128]) == 8192   # -nofc
    with pytest.raises(ValueError):
        GradBuckets(torch.zeros(8), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, split, two, one in got:
        assert split == 12
        assert torch.equal(two, one)
    assert torch.equal(got[0][2], got[1][2])
    assert got[0][2][-1].item() == 0.5                                            # (0 + 1) / 2: the accuracy slot


def test_plane_sink_layout_matches_the_concat_order():
    """model._make_sinks: the column each producer is told to fill equals the position of its tensor in the
    reference's concats (model.py:60-63 MergedEdgeConv input; model.py:83-85 FC0 input = tensors ++ merged)."""
    import torch
    from types import SimpleNamespace
    from dgcnn import model as M, ops
    fl = SimpleNamespace(MODEL_NAME="dgcnn", EDGE_CONV_LAYERS=3, EDGE_CONV_FILTERS=[64, 32, 64], FC_LAYERS=2,
                         FC_FILTERS=[512, 256])
    P = 2048
    sk = M._make_sinks(fl, P, torch.device("meta"))
    assert sk is not None
    widths = [64, 64, 64, 32, 32, 64, 64, 64, 64]                 # (max, mean, net) per layer
    assert tuple(sk.planes["FC0"].shape) == (2, P, sum(widths) + 1024)
    assert tuple(sk.planes["MergedEdgeConv"].shape) == (2, P, 3 * 64)
    assert tuple(sk.planes["FC1"].shape) == (2, P, 512)
    col = 0
    for i, f in enumerate([64, 32, 64]):
        (pl, c), = sk.targets[("ec", i, "both")]
        assert pl is sk.planes["FC0"] and c == col                # max | mean adjacent
        (p0, c0), (p1, c1) = sk.targets[("ec", i, "net")]
        assert p0 is sk.planes["FC0"] and c0 == col + 2 * f
        assert p1 is sk.planes["MergedEdgeConv"] and c1 == 64 * i
        col += 2 * f + 64
    assert sk.targets[("layer", "MergedEdgeConv")] == [(sk.planes["FC0"], col)]
    assert sk.targets[("layer", "FC0")] == [(sk.planes["FC1"], 0)]
    # no FC head / tiny clouds: producers keep their separate split passes
    fl.MODEL_NAME = "residual-dgcnn-nofc"
    assert M._make_sinks(fl, P, torch.device("meta")) is None
    fl.MODEL_NAME = "dgcnn"
    assert M._make_sinks(fl, 512, torch.device("meta")) is None
    assert ops._sinks is None


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU oracle timed on the host cores) runs without a GPU and prints ONE JSON line
    carrying the keys the driver reads."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("points/sec")


def test_checkpoint_pruning_ignores_foreign_files(dg, tmp_path):
    """_prune_checkpoints keeps the newest CHECKPOINT_NUM files named PREFIX-<iteration> and never touches (or trips over)
    neighbours such as PREFIX-100.bak or PREFIX-final (main_funcs.py:82-83: Saver(max_to_keep))."""
    from dgcnn.main_funcs import _prune_checkpoints
    prefix = str(tmp_path / "snapshot")
    for name in ("-3", "-10", "-200", "-100.bak", "-final"):
        open(prefix + name, "w").close()
    _prune_checkpoints(prefix, 2)
    left = sorted(p.name for p in tmp_path.iterdir())
    assert left == ["snapshot-10", "snapshot-100.bak", "snapshot-200", "snapshot-final"]


def test_oracle_bf16_matmul_mode_is_scoped_and_close(oracle):
    """oracle.bf16_matmul(): the yardstick of the reduced-precision variant rounds matmul operands to bf16 inside the
    context only; outside it the fp32 path is untouched."""
    import torch
    fl = oracle.make_flags(EDGE_CONV_LAYERS=1, KVALUE=4, FC_FILTERS=[16, 8], TRAIN=False)
    P = oracle.init_params(fl, 3, seed=1)
    x = torch.rand((1, 32, 3), generator=torch.Generator().manual_seed(0))
    ref = oracle.build(x, fl, P)
    with oracle.bf16_matmul():
        low = oracle.build(x, fl, P)
    again = oracle.build(x, fl, P)
    assert torch.equal(ref, again)
    d = (low - ref).abs().max().item()
    assert 0.0 < d < 0.2
