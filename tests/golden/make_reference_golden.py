"""Golden vectors produced by the REFERENCE'S OWN SOURCE: /root/reference/dgcnn/ops.py, model.py and trainval.py,
unmodified, are loaded by path and executed on top of oracle/tf1_shim (a stand-in for the few TensorFlow 1.x entry points they
call; TF1 itself cannot be installed here).  The reference's index arithmetic (ops.py:21-40), scopes, concat orders,
residual wiring (ops.py:100-140) and head (model.py:60-104) therefore come from the reference, not from this repo's
restatement; only the TF primitives are restated (oracle/tf1_shim/tensorflow/__init__.py, contrib/slim.py).

    python tests/golden/make_reference_golden.py        # needs /root/reference (this container only)

writes tests/golden/ref_*.npz in the layout of make_golden.py (x, labels, logits, loss, acc, dropout_mask, idx<i>,
tensor<i>, param:<name>, grad:<name>) plus the k_nn / edges cases `ref_knn_edges.npz`.  tests/test_oracle_vs_reference.py
checks oracle/ against these files (CPU) and, where /root/reference exists, re-runs this script's cases live; the GPU
tests compare the CUDA path with them like with any other golden file.

Loss / accuracy follow /root/reference/dgcnn/trainval.py:39-52 (softmax, argmax == label, mean sparse cross-entropy).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("DGCNN_REFERENCE_DIR", "/root/reference")
SHIM = os.path.join(ROOT, "oracle", "tf1_shim")


def load_reference():
    """-> (tf shim module, reference ops module, reference model module).  The reference's package __init__ uses
    Python-2 implicit relative imports, so its two hot-path files are loaded by path into a synthetic `dgcnn` package
    (model.py:7 does `import dgcnn` and calls dgcnn.ops.*)."""
    if "dgcnn" in sys.modules and not getattr(sys.modules["dgcnn"], "_is_reference", False):
        raise RuntimeError("the product package `dgcnn` is already imported in this process; run the reference in its own")
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    import tensorflow as tf
    if getattr(tf, "_get_variable", None) is None:
        raise RuntimeError("a real tensorflow shadows oracle/tf1_shim")
    if "dgcnn" in sys.modules:
        pkg = sys.modules["dgcnn"]
        return tf, pkg.ops, pkg.model
    pkg = types.ModuleType("dgcnn")
    pkg._is_reference = True
    pkg.__path__ = []
    sys.modules["dgcnn"] = pkg
    mods = {}
    for name in ("ops", "model", "trainval"):
        spec = importlib.util.spec_from_file_location("dgcnn." + name, os.path.join(REFERENCE, "dgcnn", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["dgcnn." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        mods[name] = mod
    return tf, mods["ops"], mods["model"]


def _flags(**kw):
    d = dict(NUM_CLASS=2, MODEL_NAME="dgcnn", TRAIN=True, KVALUE=20, DEBUG=False, EDGE_CONV_LAYERS=3, EDGE_CONV_FILTERS=64,
             FC_LAYERS=2, FC_FILTERS=[512, 256])     # /root/reference/dgcnn/flags.py:9-45 defaults of what build() reads
    d.update(kw)
    return types.SimpleNamespace(**d)


def _params(flags, C0, seed):
    """Initial values only (xavier weights, small random betas), keyed by TF variable name; drawn here with numpy so that
    the fixture does not depend on any code of this repo."""
    rng = np.random.RandomState(seed)
    L = int(flags.EDGE_CONV_LAYERS)
    filt = flags.EDGE_CONV_FILTERS if isinstance(flags.EDGE_CONV_FILTERS, list) else [int(flags.EDGE_CONV_FILTERS)] * L
    shapes = []
    cin = C0
    for i in range(L):
        shapes.append(("EdgeConv%d/conv0" % i, 2 * cin, filt[i]))
        shapes.append(("EdgeConv%d/conv1" % i, 2 * filt[i], 64))
        if flags.MODEL_NAME != "dgcnn" and i > 0 and filt[i] != filt[i - 1]:
            shapes.append(("EdgeConv%d/shortcut" % i, 64, filt[i]))
        cin = 64
    if flags.MODEL_NAME == "residual-dgcnn-nofc":
        shapes.append(("Final", 64, int(flags.NUM_CLASS)))
    else:
        shapes.append(("MergedEdgeConv", 64 * L, 1024))
        width = 1024 + sum(2 * f + 64 for f in filt) + 1024
        for j, f in enumerate(flags.FC_FILTERS):
            shapes.append(("FC%d" % j, width, f))
            width = f
        shapes.append(("Final", width, int(flags.NUM_CLASS)))
    P = {}
    for scope, ci, co in shapes:
        lim = np.sqrt(6.0 / (ci + co))
        P[scope + "/weights"] = ((rng.random_sample((ci, co)) * 2 - 1) * lim).astype(np.float32)
        P[scope + "/BatchNorm/beta"] = (0.1 * rng.standard_normal(co)).astype(np.float32)
    return P


def run_model(flags, x, labels, seed, dtype=torch.float32):
    """One forward + backward of the reference's graph.  -> dict in make_golden.py's layout."""
    tf, ref_ops, ref_model = load_reference()
    tf.reset(seed)
    P = _params(flags, x.shape[-1], seed)
    for n, v in P.items():
        tf.PRESET["dgcnn/" + n] = v
    mask = None
    if flags.TRAIN and flags.MODEL_NAME != "residual-dgcnn-nofc":
        rng = np.random.RandomState(1000 + seed)
        mask = torch.from_numpy((rng.random_sample((x.shape[0], x.shape[1], 1, flags.FC_FILTERS[-1])) < 0.7).astype(np.float32))
        tf.DROPOUT_MASK = mask
    captured = {}
    wrapped = {}
    for fn_name in ("repeat_edge_conv", "repeat_residual_edge_conv"):      # instrumentation: keep the returned tensors
        orig = getattr(ref_ops, fn_name)
        wrapped[fn_name] = orig

        def keep(*a, _orig=orig, **k):
            res = _orig(*a, **k)
            captured["tensors"] = list(res)
            return res
        setattr(ref_ops, fn_name, keep)
    try:
        with tf.variable_scope("dgcnn", reuse=tf.AUTO_REUSE):              # trainval.py:29
            pred = ref_model.build(torch.from_numpy(x).to(dtype), flags)   # trainval.py:38
    finally:
        for fn_name, orig in wrapped.items():
            setattr(ref_ops, fn_name, orig)
    lab = torch.from_numpy(labels).long()
    correct = (pred.argmax(dim=2) == lab)                                  # trainval.py:41
    acc = correct.to(torch.float32).mean()                                 # trainval.py:42
    xent = torch.nn.functional.cross_entropy(pred.reshape(-1, pred.shape[-1]), lab.reshape(-1), reduction="none")
    loss = xent.mean()                                                     # trainval.py:46-52 (no WEIGHT_KEY)
    loss.backward()
    out = {"x": x, "labels": labels, "logits": pred.detach().numpy().astype(np.float32),
           "loss": np.float32(loss.item()), "acc": np.float32(acc.item())}
    if mask is not None:
        out["dropout_mask"] = mask.numpy()
    for i, ix in enumerate(tf.TRACE["top_k"]):
        out["idx%d" % i] = ix.numpy()
        out["knn_input%d" % i] = tf.TRACE["top_k_input"][i].numpy().astype(np.float32)   # -distance matrix row source
    for i, t in enumerate(captured["tensors"]):
        out["tensor%d" % i] = t.detach().numpy().astype(np.float32)
    for n in P:
        v = tf.VARIABLES["dgcnn/" + n]
        out["param:" + n] = P[n]
        out["grad:" + n] = v.grad.reshape(P[n].shape).numpy().astype(np.float32)
    assert set(tf.VARIABLES) == set("dgcnn/" + n for n in P), sorted(set(tf.VARIABLES) ^ set("dgcnn/" + n for n in P))
    for i in range(len(tf.TRACE["top_k"])):
        del out["knn_input%d" % i]                                         # recomputable from x / tensor<3i-1>
    return out


def run_model_both(flags, x, labels, seed):
    """fp32 (the reference's arithmetic: every key of make_golden.py's layout) + the same graph in fp64 (`logits64`,
    `loss64`, `knn64_<i>`, `grad64:<name>`, stored rounded to fp32): fp32 gradients of this network carry ~1e-2 relative
    rounding noise (train-mode BN on random labels), the fp64 run pins the FORMULAE to ~1e-7."""
    out = run_model(flags, x, labels, seed)
    hi = run_model(flags, x, labels, seed, dtype=torch.float64)
    out["logits64"] = hi["logits"]
    out["loss64"] = hi["loss"]
    for k, v in hi.items():
        if k.startswith("idx"):
            out["knn64_" + k[3:]] = v
        elif k.startswith("grad:"):
            out["grad64:" + k[5:]] = v
    return out


def case_ref_dgcnn():
    rng = np.random.RandomState(21)
    x = rng.random_sample((2, 160, 3)).astype(np.float32)
    y = rng.randint(0, 2, (2, 160)).astype(np.int64)
    return run_model_both(_flags(EDGE_CONV_LAYERS=2, KVALUE=10, FC_FILTERS=[32, 16]), x, y, seed=3)


def case_ref_residual():
    rng = np.random.RandomState(22)
    x = rng.random_sample((2, 128, 4)).astype(np.float32)
    y = rng.randint(0, 3, (2, 128)).astype(np.int64)
    return run_model_both(_flags(EDGE_CONV_LAYERS=3, KVALUE=8, MODEL_NAME="residual-dgcnn", NUM_CLASS=3, FC_FILTERS=[32, 16],
                                 EDGE_CONV_FILTERS=[32, 64, 64]), x, y, seed=4)


def case_ref_residual_nofc():
    rng = np.random.RandomState(23)
    x = rng.random_sample((2, 96, 3)).astype(np.float32)
    y = rng.randint(0, 2, (2, 96)).astype(np.int64)
    return run_model_both(_flags(EDGE_CONV_LAYERS=2, KVALUE=6, MODEL_NAME="residual-dgcnn-nofc", EDGE_CONV_FILTERS=[32, 64]),
                          x, y, seed=5)


def case_ref_knn_edges():
    """k_nn (ops.py:8-19) and edges (ops.py:21-40) alone, on inputs whose pairwise distances are EXACT in fp32 whatever the
    summation order (coordinates are small dyadic rationals), so the only freedom left is the tie rule."""
    tf, ref_ops, _ = load_reference()
    out = {}
    rng = np.random.RandomState(31)
    clouds = {
        "dyadic3": (rng.randint(0, 64, (2, 200, 3)) / 64.0).astype(np.float32),         # multiples of 1/64 in [0, 1)
        "lattice3": rng.randint(0, 6, (2, 150, 3)).astype(np.float32),                   # voxel grid: massive ties, duplicates
        "feat8": (rng.randint(0, 32, (1, 96, 8)) / 16.0).astype(np.float32),             # 8 "feature" channels
    }
    for name, x in clouds.items():
        for k in (1, 7, 20):
            tf.reset()
            idx = ref_ops.k_nn(torch.from_numpy(x), k)
            out["%s:k%d:idx" % (name, k)] = idx.numpy().astype(np.int32)
        tf.reset()
        out["%s:edges" % name] = ref_ops.edges(torch.from_numpy(x), k=5).numpy()
        out["%s:x" % name] = x
    return out


TRAINER_SEED = 6


def trainer_flags():
    return _flags(EDGE_CONV_LAYERS=1, KVALUE=6, FC_FILTERS=[32, 16], GPUS=[0, 1], MINIBATCH_SIZE=2, NUM_CHANNEL=3,
                  LEARNING_RATE=0.001, WEIGHT_KEY="", TRAIN=True)


def case_ref_trainer():
    """/root/reference/dgcnn/trainval.py executed unmodified (graph mode of the shim: placeholders, towers, compute_gradients,
    tower mean, accumulation variables, apply_gradients, Session.run): two optimizer steps, each of two micro-steps on two
    towers (flags.GPUS = [0, 1], MINIBATCH_SIZE = 2), i.e. trainval.py:59-80's mean over towers, sum over micro-steps and
    Adam update as the reference's code composes them.  Stored: what accum_gradient returns per micro-step, the accumulated
    gradients before each apply, the variables after each apply, and inference() on the final variables."""
    tf, ref_ops, ref_model = load_reference()
    ref_trainval = sys.modules["dgcnn.trainval"]
    seed = TRAINER_SEED
    flags = trainer_flags()
    tf.reset(seed)
    P = _params(flags, 3, seed)
    for n, v in P.items():
        tf.PRESET["dgcnn/" + n] = v
    rng = np.random.RandomState(41)
    STEPS, MICRO, T, B, N = 2, 2, 2, 2, 48
    x = rng.random_sample((STEPS, MICRO, T, B, N, 3)).astype(np.float32)
    y = rng.randint(0, 2, (STEPS, MICRO, T, B, N)).astype(np.int32)
    masks = (rng.random_sample((STEPS, MICRO, T, B, N, 1, 16)) < 0.7).astype(np.float32)
    trainer = ref_trainval.trainval(flags)
    trainer.initialize()                                                   # builds the two-tower graph (trainval.py:12-85)
    sess = tf.Session()
    out = {"x": x, "labels": y, "dropout_masks": masks, "lr": np.float32(flags.LEARNING_RATE)}
    names = [v.name[:-2][len("dgcnn/"):] for v in tf.trainable_variables()]
    assert sorted(names) == sorted(P)
    for s in range(STEPS):
        trainer.zero_gradients(sess)                                       # main_funcs.py:135
        for m in range(MICRO):
            tf.DROPOUT_MASK = [torch.from_numpy(masks[s, m, t]) for t in range(T)]
            res = trainer.accum_gradient(sess, [x[s, m, t] for t in range(T)], [y[s, m, t] for t in range(T)])   # :155
            out["acc:%d:%d" % (s, m)] = np.float32(res[1])
            out["loss:%d:%d" % (s, m)] = np.float32(res[2])
            out["knn:%d:%d" % (s, m)] = np.stack([t.numpy() for t in tf.TRACE["top_k"]])     # [towers * layers, B, N, k]
        accum = sess.run([v for v in trainer._apply_grad.args])              # the accumulation variables (reads only)
        for n, g in zip(names, accum):
            out["accum:%d:%s" % (s, n)] = g.reshape(P[n].shape).astype(np.float32)
        trainer.apply_gradient(sess)                                       # main_funcs.py:166
        for n, v in zip(names, tf.trainable_variables()):
            out["var:%d:%s" % (s, n)] = v.tensor.detach().numpy().reshape(P[n].shape).astype(np.float32)
    # (initial values: _params(trainer_flags(), 3, TRAINER_SEED), not stored)
    tf.DROPOUT_MASK = [torch.from_numpy(masks[0, 0, t]) for t in range(T)]
    res = trainer.inference(sess, [x[0, 0, t] for t in range(T)], [y[0, 0, t] for t in range(T)])   # trainval.py:103-108
    out["inference:softmax"] = np.stack(res[:T]).astype(np.float32)       # one per tower, then accuracy, loss
    out["inference:acc"], out["inference:loss"] = np.float32(res[T]), np.float32(res[T + 1])
    return out


def case_ref_small():
    """Two light cases (no big tensors stored).
    weighted: trainval.py:46-52 with WEIGHT_KEY set -- loss = mean(xent * weight) -- one tower, one micro-step, through the
              reference's accum_gradient; stored: loss, accuracy, accumulated gradients of the EdgeConv / Final variables.
    lattice:  model.build (inference graph, TRAIN=False) on a voxel lattice with duplicated points: distance ties inside a
              whole model, resolved by tf.nn.top_k's rule in layer 0 (exact arithmetic) -- stored: indices, logits."""
    tf, ref_ops, ref_model = load_reference()
    ref_trainval = sys.modules["dgcnn.trainval"]
    out = {}
    # ---- weighted loss through the trainer
    flags = _flags(EDGE_CONV_LAYERS=1, KVALUE=5, FC_FILTERS=[16, 8], GPUS=[0], MINIBATCH_SIZE=2, NUM_CHANNEL=3,
                   LEARNING_RATE=0.001, WEIGHT_KEY="weight", TRAIN=True, NUM_CLASS=3)
    tf.reset(7)
    P = _params(flags, 3, 7)
    for n, v in P.items():
        tf.PRESET["dgcnn/" + n] = v
    rng = np.random.RandomState(51)
    x = rng.random_sample((2, 40, 3)).astype(np.float32)
    y = rng.randint(0, 3, (2, 40)).astype(np.int32)
    w = (rng.random_sample((2, 40)) * 2.0).astype(np.float32)
    mask = (rng.random_sample((2, 40, 1, 8)) < 0.7).astype(np.float32)
    trainer = ref_trainval.trainval(flags)
    trainer.initialize()
    sess = tf.Session()
    trainer.zero_gradients(sess)
    tf.DROPOUT_MASK = [torch.from_numpy(mask)]
    res = trainer.accum_gradient(sess, [x], [y], [w])
    knn = np.stack([t.numpy() for t in tf.TRACE["top_k"]])                  # (the next run resets the trace)
    names = [v.name[:-2][len("dgcnn/"):] for v in tf.trainable_variables()]
    accum = sess.run([v for v in trainer._apply_grad.args])
    out.update({"weighted:x": x, "weighted:labels": y, "weighted:weight": w, "weighted:mask": mask,
                "weighted:acc": np.float32(res[1]), "weighted:loss": np.float32(res[2]), "weighted:knn": knn})
    for n, g_ in zip(names, accum):
        if n.startswith(("EdgeConv", "Final")) or n.endswith("beta"):
            out["weighted:accum:" + n] = g_.reshape(P[n].shape).astype(np.float32)
    # ---- lattice model, inference graph
    flags = _flags(EDGE_CONV_LAYERS=2, KVALUE=9, FC_FILTERS=[16, 8], TRAIN=False)
    tf.reset(8)
    P = _params(flags, 3, 8)
    for n, v in P.items():
        tf.PRESET["dgcnn/" + n] = v
    rng = np.random.RandomState(52)
    xl = rng.randint(0, 5, (2, 90, 3)).astype(np.float32)
    xl[:, 60:] = xl[:, :30]                                                  # 30 duplicated points per cloud
    with tf.variable_scope("dgcnn", reuse=tf.AUTO_REUSE):
        pred = ref_model.build(torch.from_numpy(xl), flags)
    out.update({"lattice:x": xl, "lattice:logits": pred.detach().numpy().astype(np.float32)})
    for i, ix in enumerate(tf.TRACE["top_k"]):
        out["lattice:idx%d" % i] = ix.numpy()
    return out


SMALL_SEEDS = {"weighted": 7, "lattice": 8}

CASES = {"ref_dgcnn": case_ref_dgcnn, "ref_trainer": case_ref_trainer, "ref_small": case_ref_small, "ref_residual": case_ref_residual, "ref_residual_nofc": case_ref_residual_nofc,
         "ref_knn_edges": case_ref_knn_edges}

if __name__ == "__main__":
    for name, fn in CASES.items():
        out = fn()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if not k.startswith(("param", "grad"))})
