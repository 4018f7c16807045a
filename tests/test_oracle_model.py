"""Pins the torch-CPU restatement of ops.py / model.py (oracle/dgcnn_oracle.py) with known answers and with the
committed golden vectors (tests/golden/, produced by tests/golden/make_golden.py from the oracle itself --
the reference cannot run here: Python 2 + TensorFlow 1.x)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_parameter_inventory(oracle):
    fl = oracle.make_flags(EDGE_CONV_LAYERS=4)
    shapes = oracle.param_shapes(fl, 3)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 1895554
    assert shapes["EdgeConv0/conv0/weights"] == (6, 64) and shapes["FC0/weights"] == (2816, 512)
    assert shapes["Final/BatchNorm/beta"] == (2,)
    assert "EdgeConv0/conv0/BatchNorm/gamma" not in shapes  # scale=False


def test_edges_known_answer(oracle):
    x = torch.tensor([[[0.0, 0.0], [1.0, 0.0], [3.0, 0.0]]])
    e = oracle.edges(x, k=2)
    # point 0: neighbours (0, 1); point 2: neighbours (2, 1)
    assert e[0, 0].tolist() == [[0, 0, 0, 0], [0, 0, 1, 0]]
    assert e[0, 2].tolist() == [[3, 0, 0, 0], [3, 0, -2, 0]]


def test_bn_train_known_answer(oracle):
    t = torch.tensor([[1.0, 10.0], [3.0, 10.0]])
    y = oracle.bn_train(t, torch.tensor([0.5, -1.0]))
    # channel 0: mean 2, biased var 1 -> +-1/sqrt(1.001) ; channel 1: var 0 -> 0 ; then + beta
    r = 1.0 / np.sqrt(1.001)
    assert np.allclose(y.numpy(), [[-r + 0.5, -1.0], [r + 0.5, -1.0]], atol=1e-6)


def test_adam_tf_form(oracle):
    p, g = torch.tensor([1.0]), torch.tensor([0.5])
    m, v = torch.zeros(1), torch.zeros(1)
    oracle.adam_tf_step(p, g, m, v, 1, lr=0.1)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.isclose(p.item(), 1.0 - lr_t * 0.05 / (np.sqrt(0.00025) + 1e-8), rtol=1e-6)


def test_logits_are_relu_and_shapes(oracle):
    torch.manual_seed(0)
    x = torch.rand(2, 64, 3)
    for name in ("dgcnn", "residual-dgcnn", "residual-dgcnn-nofc"):
        fl = oracle.make_flags(EDGE_CONV_LAYERS=2, KVALUE=8, TRAIN=False, MODEL_NAME=name)
        out = oracle.build(x, fl, oracle.init_params(fl, 3))
        assert out.shape == (2, 64, 2) and (out >= 0).all()
    with pytest.raises(NotImplementedError):
        oracle.build(x, oracle.make_flags(MODEL_NAME="nope"), {})
    with pytest.raises(ValueError):
        oracle.repeat_edge_conv(x, 2, [8], 64, {})


def test_max_commutes_with_bn_relu(oracle):
    """The identity the fused kernel relies on: max_k relu(bn(z)) == relu(bn(max_k z)) (rstd > 0, no gamma)."""
    torch.manual_seed(1)
    z = torch.randn(2, 16, 5, 8)
    beta = torch.randn(8)
    dims = (0, 1, 2)
    mean, var = z.mean(dims, keepdim=True), z.var(dims, unbiased=False, keepdim=True)
    f = lambda t: torch.relu((t - mean) / torch.sqrt(var + 1e-3) + beta)  # noqa: E731
    assert torch.equal(f(z).amax(dim=2, keepdim=True), f(z.amax(dim=2, keepdim=True)))


@pytest.mark.parametrize("name", ["cfg1_dgcnn", "residual", "lattice"])
def test_golden_vectors_reproduce(oracle, name):
    from tests.golden import make_golden as mg
    path = os.path.join(GOLD, name + ".npz")
    assert os.path.exists(path), "run python tests/golden/make_golden.py"
    z = np.load(path)
    got = mg.CASES[name]()
    for key in z.files:
        a, b = z[key], got[key]
        if a.dtype.kind in "iu":
            assert np.array_equal(a, b), key
        else:
            assert np.allclose(a, b, rtol=1e-4, atol=1e-5), key
