"""Generates tests/golden/*.npz from the oracle (oracle/dgcnn_oracle.py), fixed seeds.

The reference itself cannot produce vectors here (Python 2 + TensorFlow 1.x), so these
fixtures pin (a) the oracle against regressions on CPU and (b) the CUDA path on the GPU box, where
/root/reference and large oracle runs are not needed.  Run:  python tests/golden/make_golden.py
Each case stores inputs, every layer's kNN indices (bit-exact target), the [max, mean, net] tensors of each
EdgeConv layer, logits, loss/accuracy and all parameter gradients (TRAIN=True cases use a fixed dropout mask).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dgcnn_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(flags, x, labels, seed=0, train=True):
    torch.manual_seed(seed)
    P = O.init_params(flags, x.shape[-1], seed=seed)
    g = torch.Generator().manual_seed(100 + seed)
    for n, t in P.items():  # non-zero betas so that beta gradients / offsets are exercised
        if n.endswith("beta"):
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
    for t in P.values():
        t.requires_grad_(True)
    idx_out, tensors = [], []
    mask = None
    if train and flags.MODEL_NAME != "residual-dgcnn-nofc":
        nfc = O._listify(flags.FC_FILTERS, int(flags.FC_LAYERS), "f")
        mask = (torch.rand((x.shape[0], x.shape[1], 1, nfc[-1]), generator=g) < O.DROPOUT_KEEP).float()
    flags.TRAIN = train
    logits = O.build(x, flags, P, idx_out=idx_out, dropout_mask=mask, tensors_out=tensors)
    _, acc, loss = O.softmax_loss_accuracy(logits, labels)
    loss.backward()
    out = {"x": x.numpy(), "labels": labels.numpy(), "logits": logits.detach().numpy(),
           "loss": np.float32(loss.item()), "acc": np.float32(acc.item())}
    if mask is not None:
        out["dropout_mask"] = mask.numpy()
    for i, ix in enumerate(idx_out):
        out["idx%d" % i] = ix.numpy()
    for i, t in enumerate(tensors):
        out["tensor%d" % i] = t.detach().numpy()
    for n, t in P.items():
        out["param:" + n] = t.detach().numpy()
        out["grad:" + n] = t.grad.numpy()
    return out


def case_cfg1():
    """BASELINE.json configs[0]: 1 EdgeConv layer, N=512, k=20, C=3, bs=2."""
    g = torch.Generator().manual_seed(1234)
    x = torch.rand((2, 512, 3), generator=g)
    y = torch.randint(0, 2, (2, 512), generator=g)
    return _run(O.make_flags(EDGE_CONV_LAYERS=1, KVALUE=20, FC_FILTERS=[64, 32]), x, y)


def case_residual():
    g = torch.Generator().manual_seed(7)
    x = torch.rand((2, 192, 4), generator=g)
    y = torch.randint(0, 3, (2, 192), generator=g)
    fl = O.make_flags(EDGE_CONV_LAYERS=3, KVALUE=12, MODEL_NAME="residual-dgcnn", NUM_CLASS=3, FC_FILTERS=[48, 24],
                      EDGE_CONV_FILTERS=[32, 64, 64])
    return _run(fl, x, y, seed=1)


def case_lattice():
    """Voxel-like integer coordinates (tie-heavy kNN), inference mode."""
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, 24, (2, 256, 3), generator=g).float()
    y = torch.randint(0, 2, (2, 256), generator=g)
    return _run(O.make_flags(EDGE_CONV_LAYERS=2, KVALUE=16, FC_FILTERS=[32, 16]), x, y, seed=2, train=False)


CASES = {"cfg1_dgcnn": case_cfg1, "residual": case_residual, "lattice": case_lattice}

if __name__ == "__main__":
    for name, fn in CASES.items():
        out = fn()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if not k.startswith(("param", "grad"))})
