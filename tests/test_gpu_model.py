"""GPU parity of the whole model / training step through the public API (dgcnn.model.build, dgcnn.trainval)
against the oracle and the committed golden vectors.  north_star tolerance: logits within 1e-3 (fp32).

kNN in feature space (layers >= 1) is a discontinuous function of activations that differ from the oracle's by
fp32 reassociation (~1e-6), so a near-tie can legitimately pick another neighbour.  Strict comparisons therefore
teacher-force the oracle's indices; free-running runs are checked for index agreement and logits statistics."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _flags_for(oracle, name, z):
    P = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    L = len([k for k in z.files if k.startswith("idx")])
    filt = [P["EdgeConv%d/conv0/weights" % i].shape[1] for i in range(L)]
    nfc = len([k for k in P if k.startswith("FC") and k.endswith("weights")])
    fcf = [P["FC%d/weights" % j].shape[1] for j in range(nfc)]
    fl = oracle.make_flags(EDGE_CONV_LAYERS=L, EDGE_CONV_FILTERS=filt, KVALUE=z["idx0"].shape[-1], FC_LAYERS=nfc,
                           FC_FILTERS=fcf, NUM_CLASS=P["Final/weights"].shape[1],
                           MODEL_NAME="residual-dgcnn" if "residual" in name else "dgcnn",
                           TRAIN="dropout_mask" in z.files, NUM_CHANNEL=z["x"].shape[-1])
    return fl, P, L


def _trainer(dg, fl, P):
    tr = dg.trainval(fl)
    tr.initialize()
    tr.variables.load_state_dict({"dgcnn/" + n: t for n, t in P.items()})
    return tr


# ref_*: vectors produced by the reference's own ops.py / model.py (tests/golden/make_reference_golden.py)
@pytest.mark.parametrize("name", ["cfg1_dgcnn", "residual", "lattice", "ref_dgcnn", "ref_residual"])
def test_model_forward_backward_vs_golden(dg, oracle, cuda, name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    fl, P, L = _flags_for(oracle, name, z)
    tr = _trainer(dg, fl, P)
    x = torch.from_numpy(z["x"]).cuda()
    y = torch.from_numpy(z["labels"]).cuda().long()
    mask = torch.from_numpy(z["dropout_mask"]).cuda() if fl.TRAIN else None
    dg.ops._knn_forced = iter([torch.from_numpy(z["idx%d" % i]) for i in range(L)])
    try:
        if fl.TRAIN:
            tr.zero_gradients(None)
        ctx = torch.enable_grad() if fl.TRAIN else torch.no_grad()
        with ctx:
            from dgcnn.variables import set_default_store
            old = set_default_store(tr.variables)
            with tr.variables.variable_scope("dgcnn"):
                logits = dg.build(x, fl, dropout_mask=mask)
            set_default_store(old)
            loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), y.reshape(-1))
            if fl.TRAIN:
                loss.backward()
    finally:
        dg.ops._knn_forced = None
    ref = torch.from_numpy(z["logits"])
    err = (logits.detach().cpu() - ref).abs().max().item()
    assert err <= 1e-3, "logits differ by %g" % err                       # north_star bound
    assert abs(loss.item() - float(z["loss"])) <= 1e-4
    if fl.TRAIN:
        for n in P:
            a = tr.variables.vars["dgcnn/" + n].grad.cpu()
            b = torch.from_numpy(z["grad:" + n])
            scale = max(float(b.abs().max()), 1e-3)
            # Head GEMMs run as a bf16 hi/lo split on the tensor cores (~2^-16 relative per operand, 64x finer than
            # TF32).  Two things amplify that on gradients: BN-backward cancellation on the earliest layers, and the
            # global max-pool (model.py:77) whose argmax can move between near-tied points -- a discontinuity exactly
            # like a kNN flip.  So: (almost) every element within 1e-2 of max|grad|, none off by more than 10 %.
            err = (a - b).abs()
            assert (err > 1e-2 * scale).float().mean().item() <= 2e-3, (n, err.max().item(), scale)
            assert err.max().item() <= 0.1 * scale, (n, err.max().item(), scale)


@pytest.mark.parametrize("name", ["cfg1_dgcnn", "lattice"])
def test_model_free_running_vs_golden(dg, oracle, cuda, name):
    """No teacher forcing: every layer's kNN computed on the GPU's own activations."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    fl, P, L = _flags_for(oracle, name, z)
    fl.TRAIN = False
    tr = _trainer(dg, fl, P)
    dg.ops._knn_trace = []
    try:
        out = tr.inference(None, [z["x"]], [z["labels"]])
    finally:
        trace, dg.ops._knn_trace = dg.ops._knn_trace, None
    assert torch.equal(trace[0].cpu(), torch.from_numpy(z["idx0"]))       # layer 0: bit-exact
    for i in range(1, L):
        agree = (trace[i].cpu() == torch.from_numpy(z["idx%d" % i])).float().mean().item()
        assert agree >= 0.995, (i, agree)
    softmax = out[0]
    assert softmax.shape == z["logits"].shape and np.allclose(softmax.sum(-1), 1.0, atol=1e-5)
    if not ("dropout_mask" in z.files):   # golden was produced in inference mode: logits comparable
        ref = torch.softmax(torch.from_numpy(z["logits"]), -1).numpy()
        frac = (np.abs(softmax - ref) <= 1e-3).mean()
        assert frac >= 0.99, frac


def test_trainer_api_step_matches_oracle_adam(dg, oracle, cuda):
    """zero_gradients -> accum_gradient x2 (2 towers x 2 micro-steps) -> apply_gradient == oracle: tower mean,
    micro-step SUM (trainval.py:64-69,79) and the TF-form Adam update."""
    torch.manual_seed(0)
    fl = oracle.make_flags(EDGE_CONV_LAYERS=1, KVALUE=8, FC_FILTERS=[32, 16], GPUS=[0, 1], MINIBATCH_SIZE=2,
                           NUM_CHANNEL=3, TRAIN=True, MODEL_NAME="residual-dgcnn-nofc", LEARNING_RATE=0.01)
    P = oracle.init_params(fl, 3, seed=5)
    tr = _trainer(dg, fl, P)
    data = torch.rand(2, 2, 2, 96, 3)          # [micro, tower, MBS, N, C]
    label = torch.randint(0, 2, (2, 2, 2, 96))
    tr.zero_gradients(None)
    for m in range(2):
        res = tr.accum_gradient(None, [data[m, 0].numpy(), data[m, 1].numpy()], [label[m, 0].numpy(), label[m, 1].numpy()])
        assert len(res) == 3 and np.isfinite(res[2])
    grads_gpu = {n[len("dgcnn/"):]: v.grad.detach().cpu().clone() for n, v in tr.variables.vars.items()}
    tr.apply_gradient(None)
    # oracle
    for t in P.values():
        t.requires_grad_(True)
    total = 0
    for m in range(2):
        tower_losses = []
        for g in range(2):
            lg = oracle.build(data[m, g], fl, P)
            tower_losses.append(oracle.softmax_loss_accuracy(lg, label[m, g])[2])
        total = total + torch.stack(tower_losses).mean()
    total.backward()
    for n, t in P.items():
        scale = max(float(t.grad.abs().max()), 1e-3)
        assert (grads_gpu[n] - t.grad).abs().max().item() <= 1e-2 * scale, n
    for n, t in P.items():
        p = t.detach().clone()
        # the Adam kernel is checked on the gradient the GPU actually produced (near-zero gradient entries
        # would otherwise turn a 1e-9 gradient difference into a visible update difference)
        oracle.adam_tf_step(p, grads_gpu[n], torch.zeros_like(p), torch.zeros_like(p), 1, lr=0.01)
        got = tr.variables.vars["dgcnn/" + n].detach().cpu()
        assert torch.allclose(got, p, atol=1e-6), (n, (got - p).abs().max())
    with pytest.raises(NotImplementedError):
        fl2 = oracle.make_flags(TRAIN=False, NUM_CHANNEL=3, EDGE_CONV_LAYERS=1)
        t2 = dg.trainval(fl2)
        t2.initialize()
        t2.accum_gradient(None, [], [])


def test_checkpoint_roundtrip(dg, oracle, cuda, tmp_path):
    fl = oracle.make_flags(EDGE_CONV_LAYERS=1, KVALUE=8, FC_FILTERS=[32, 16], NUM_CHANNEL=3)
    tr = dg.trainval(fl)
    tr.initialize()
    path = tr.save(str(tmp_path / "snapshot"), 41)
    assert path.endswith("snapshot-41")
    from dgcnn.main_funcs import iteration_from_filename
    assert iteration_from_filename(path) == 41
    before = {n: v.detach().clone() for n, v in tr.variables.vars.items()}
    with torch.no_grad():
        tr.variables.flat_param.add_(1.0)
    tr.restore(path)
    for n, v in tr.variables.vars.items():
        assert torch.equal(v.detach(), before[n])
    assert "dgcnn/EdgeConv0/conv0/BatchNorm/beta" in before          # TF variable names


def test_cli_train_config1_plumbing(dg, cuda, tmp_path):
    """BASELINE configs[0]: bin/dgcnn.py train, 1 EdgeConv layer, N=512, k=20, C=3, bs=2, synthetic data."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "dynamic-gcnn_b200", "bin", "dgcnn.py"), "train", "-io", "synthetic",
           "-bs", "2", "-mbs", "2", "-ecl", "1", "-kv", "20", "-np", "512", "-it", "3", "-rs", "1", "-db", "0",
           "-ld", str(tmp_path / "log"), "-wp", str(tmp_path / "w" / "snapshot"), "-chks", "2", "-sd", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Iteration 2" in out.stdout and "saved @" in out.stdout
    csv = open(str(tmp_path / "log" / "train_log-0000000.csv")).read().splitlines()
    assert csv[0].startswith("iter,epoch,titer,ttrain,tio,tsave,tsummary") and len(csv) == 4
    assert os.path.exists(str(tmp_path / "w" / "snapshot-1"))


def test_cli_inference_from_checkpoint(dg, cuda, tmp_path):
    """bin/dgcnn.py train (2 iterations, checkpoint every step) then bin/dgcnn.py inference -mp <checkpoint> -of out.npz:
    the reference's second sub-command (main_funcs.py:47-51,212-305) through the same files: softmax rows are stored per
    entry in request order, sum to one, and the CSV carries the reference's inference columns."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = [sys.executable, os.path.join(root, "dynamic-gcnn_b200", "bin", "dgcnn.py")]
    common = ["-io", "synthetic", "-bs", "2", "-mbs", "2", "-ecl", "2", "-kv", "16", "-np", "384", "-db", "0", "-sd", "3",
              "-rs", "1"]
    tr = subprocess.run(exe + ["train"] + common + ["-it", "2", "-wp", str(tmp_path / "w" / "snap"), "-chks", "1"],
                        capture_output=True, text=True, timeout=600)
    assert tr.returncode == 0, tr.stderr[-2000:]
    ckpt = str(tmp_path / "w" / "snap-1")
    assert os.path.exists(ckpt)
    out = str(tmp_path / "pred.npz")
    inf = subprocess.run(exe + ["inference"] + common + ["-it", "3", "-sh", "0", "-mp", ckpt, "-of", out,
                                                         "-ld", str(tmp_path / "ilog")],
                         capture_output=True, text=True, timeout=600)
    assert inf.returncode == 0, inf.stderr[-2000:]
    z = np.load(out)
    # 3 iterations x 2 entries, in sequential order; the first batch (entries 0, 1) is consumed by prepare()'s throw-away
    # next() exactly like the reference (main_funcs.py:66)
    assert z["softmax"].shape == (6, 384, 2) and list(z["index"]) == [2, 3, 4, 5, 6, 7]
    assert np.allclose(z["softmax"].sum(-1), 1.0, atol=1e-5)
    csv = open(str(tmp_path / "ilog" / "inference_log-0000001.csv")).read().splitlines()
    assert csv[0] == "iter,epoch,titer,tinference,tio,tsumiter,tsuminference,tsumio,loss,accuracy" and len(csv) == 4


def test_cli_config0_hdf5_in_hdf5_out(dg, cuda, tmp_path):
    """BASELINE configs[0] through the reference's own file formats: bin/dgcnn.py train -io h5 on an HDF5 file (keys data /
    label, iotool.py:213-224), then inference writing the PyTables-style output file (iotool.py:226-250); both files go
    through dgcnn.h5lite (no h5py / PyTables in the image)."""
    import subprocess
    import sys
    from dgcnn import h5lite
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = [sys.executable, os.path.join(root, "dynamic-gcnn_b200", "bin", "dgcnn.py")]
    rng = np.random.RandomState(5)
    src = str(tmp_path / "clouds.h5")
    data = rng.rand(8, 512, 3).astype(np.float32)
    label = (data[..., 0] > 0.5).astype(np.int32)
    h5lite.write(src, {"data": data, "label": label}, compress=5)
    common = ["-io", "h5", "-if", src, "-bs", "2", "-mbs", "2", "-ecl", "1", "-kv", "20", "-np", "512", "-db", "0",
              "-sd", "3", "-rs", "1"]
    tr = subprocess.run(exe + ["train"] + common + ["-it", "2", "-wp", str(tmp_path / "w" / "snap"), "-chks", "1",
                                                    "-ld", str(tmp_path / "log")],
                        capture_output=True, text=True, timeout=600)
    assert tr.returncode == 0, tr.stderr[-2000:]
    assert "Iteration 1" in tr.stdout                                  # iterations are reported 0-based
    out = str(tmp_path / "pred.h5")
    inf = subprocess.run(exe + ["inference"] + common + ["-it", "2", "-sh", "0", "-mp", str(tmp_path / "w" / "snap-1"),
                                                         "-of", out, "-ld", str(tmp_path / "ilog")],
                         capture_output=True, text=True, timeout=600)
    assert inf.returncode == 0, inf.stderr[-2000:]
    with h5lite.File(out) as f:
        sm, idx = f["softmax"], f["index"]
        assert sm.shape == (4, 512, 2) and idx.tolist() == [2, 3, 4, 5]      # entries 0, 1 go to prepare()'s throw-away next()
        assert np.allclose(sm.sum(-1), 1.0, atol=1e-5)
        assert np.array_equal(f["data"], data[idx]) and np.array_equal(f["label"], label[idx].astype(np.float32))
        assert f.attrs("softmax")["CLASS"] == "EARRAY"
