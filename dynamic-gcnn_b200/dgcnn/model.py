"""dgcnn.model -- mirror of /root/reference/dgcnn/model.py:9-106 (same `build(point_cloud, flags)` signature).

The EdgeConv stack (the hot path) runs on the hand-written kernels behind dgcnn.ops; the head follows the
reference op for op.  Variables are created-or-reused in the default VariableStore under the caller's scope
(trainval opens "dgcnn" like trainval.py:29), with the reference's TF variable names.
"""
from __future__ import annotations

import torch

from . import ops
from .head import conv_bn_act, conv_bn_relu_dense
from .variables import default_store

DROPOUT_KEEP = 0.7  # model.py:91: tf.nn.dropout(net, 0.7, None) -> keep_prob


def build(point_cloud, flags, dropout_mask=None):
    """-> logits [B,N,NUM_CLASS] (ReLU'd and batch-normalised like the reference's `Final`, model.py:94-101).

    `dropout_mask` (optional, {0,1} keep mask [B,N,1,width]) replaces the random draw -- used by parity tests.
    """
    num_edge_conv = int(flags.EDGE_CONV_LAYERS)
    num_edge_filters = flags.EDGE_CONV_FILTERS
    num_fc = int(flags.FC_LAYERS)
    num_fc_filters = flags.FC_FILTERS
    is_training = bool(flags.TRAIN)
    k = int(flags.KVALUE)
    debug = bool(flags.DEBUG)
    num_class = int(flags.NUM_CLASS)

    net = point_cloud
    batch_size, num_point = net.shape[0], net.shape[1]
    if debug:
        print("\n")
        print("Shape %s ... Name %s" % (tuple(net.shape), "points"))

    if flags.MODEL_NAME not in ("dgcnn", "residual-dgcnn", "residual-dgcnn-nofc"):
        print("Unsupported MODEL_NAME: %s" % flags.MODEL_NAME)
        raise NotImplementedError
    dtype = str(getattr(flags, "DTYPE", "f32") or "f32").lower()
    if dtype not in ("f32", "fp32", "bf16"):
        print("Unsupported DTYPE: %s (f32 | bf16)" % dtype)
        raise NotImplementedError
    old_sinks, old_prec = ops._sinks, ops._precision
    ops._precision = "bf16" if dtype == "bf16" else "f32"
    ops._sinks = _make_sinks(flags, batch_size * num_point, net.device)
    try:
        return _build_body(net, flags, dropout_mask, num_edge_conv, num_edge_filters, num_fc, num_fc_filters, is_training,
                           k, debug, num_class, batch_size, num_point)
    finally:
        ops._sinks, ops._precision = old_sinks, old_prec


def _make_sinks(flags, P, device):
    """Pre-allocate the bf16 operand planes of MergedEdgeConv / FC0 / FC1 (model.py:60-63,83-88: the channel concats)
    and tell every producer where its output goes in them, so that producers fill the operands directly."""
    import os
    if flags.MODEL_NAME == "residual-dgcnn-nofc" or os.environ.get("DGCNN_PLANE_SINKS", "1") == "0":
        return None
    L = int(flags.EDGE_CONV_LAYERS)
    filt = [int(f) for f in ops._listify(flags.EDGE_CONV_FILTERS, L, "num_filters")]
    nfc = int(flags.FC_LAYERS)
    fcf = [int(f) for f in ops._listify(flags.FC_FILTERS, nfc, "num_filters")] if nfc else []
    widths = []
    for f in filt:
        widths += [f, f, ops.CONV1_WIDTH]
    if nfc == 0 or not ops._tc_ok(P, 1024, fcf[0], *widths):
        return None
    sk = ops.PlaneSinks()
    k_merged = ops.CONV1_WIDTH * L
    k_fc0 = sum(widths) + 1024
    bf = dict(dtype=torch.bfloat16, device=device)
    npl = ops._npl()
    sk.planes["MergedEdgeConv"] = torch.empty((npl, P, k_merged), **bf)
    sk.planes["FC0"] = torch.empty((npl, P, k_fc0), **bf)
    col = 0
    for i, f in enumerate(filt):
        sk.targets[("ec", i, "both")] = [(sk.planes["FC0"], col)]
        sk.targets[("ec", i, "net")] = [(sk.planes["FC0"], col + 2 * f), (sk.planes["MergedEdgeConv"], ops.CONV1_WIDTH * i)]
        col += 2 * f + ops.CONV1_WIDTH
    sk.targets[("layer", "MergedEdgeConv")] = [(sk.planes["FC0"], col)]
    if nfc > 1 and ops._tc_ok(P, fcf[0], fcf[1]):
        sk.planes["FC1"] = torch.empty((npl, P, fcf[0]), **bf)
        sk.targets[("layer", "FC0")] = [(sk.planes["FC1"], 0)]
    return sk


def _build_body(net, flags, dropout_mask, num_edge_conv, num_edge_filters, num_fc, num_fc_filters, is_training, k, debug,
                num_class, batch_size, num_point):
    if flags.MODEL_NAME == "dgcnn":
        tensors = ops.repeat_edge_conv(net, repeat=num_edge_conv, k=k, num_filters=num_edge_filters,
                                       trainable=is_training, debug=debug)
    else:
        tensors = ops.repeat_residual_edge_conv(net, repeat=num_edge_conv, k=k, num_filters=num_edge_filters,
                                                trainable=is_training, debug=debug)

    if flags.MODEL_NAME == "residual-dgcnn-nofc":                      # model.py:45-58
        net = conv_bn_relu_dense(tensors[-1], "Final", num_class, True)
        if debug: print("Shape %s ... Name %s" % (tuple(net.shape), "Final"))
        return net.squeeze(-2)

    P = batch_size * num_point
    flat = [x.reshape(P, x.shape[-1]) for x in tensors]
    # model.py:60-72: concat of every layer's `net` -> MergedEdgeConv (1024) ; the concat is never built
    # ... and model.py:77's global max pool of it, produced by the same BN apply pass
    merged, g = conv_bn_act([flat[3 * i + 2] for i in range(num_edge_conv)], "MergedEdgeConv", 1024, True,
                            pool_rows=num_point)
    if debug: print("Shape %s ... Name %s" % ((batch_size, num_point, 1, 1024), "MergedEdgeConv"))
    if debug: print("Shape %s ... Name %s" % ((batch_size, 1, 1, 1024), "maxpool0"))
    # model.py:80-88: tile(g) ++ tensors ++ merged -> FC stack.  First FC layer: the tiled global feature is a
    # per-cloud term, the rest a multi-source GEMM; later FC layers are plain.
    num_fc_filters = ops._listify(num_fc_filters, num_fc, "num_filters")
    if debug: print("Shape %s ... Name %s" % ((batch_size, num_point, 1, 1024 + sum(x.shape[1] for x in flat) + 1024),
                                              "concat (never materialised)"))
    srcs = flat + [merged]
    if num_fc == 0:
        net = torch.cat([g.view(batch_size, 1, 1024).expand(batch_size, num_point, 1024).reshape(P, 1024)] + srcs, dim=1)
    for i in range(num_fc):
        if i == 0:
            net = conv_bn_act(srcs, "FC0", int(num_fc_filters[0]), is_training, cloud_feature=g,
                              points_per_cloud=num_point)
        else:
            net = conv_bn_act([net], "FC%d" % i, int(num_fc_filters[i]), is_training)
        if debug: print("Shape %s ... Name %s" % ((batch_size, num_point, 1, net.shape[1]), "FC%d" % i))
    net = net.view(batch_size, num_point, 1, net.shape[1])

    if is_training:                                                                  # model.py:90-91
        if dropout_mask is None:
            net = torch.nn.functional.dropout(net, p=1.0 - DROPOUT_KEEP, training=True)
        else:
            net = net * dropout_mask / DROPOUT_KEEP
        if debug: print("Shape %s ... Name %s" % (tuple(net.shape), "dropout"))

    net = conv_bn_relu_dense(net, "Final", num_class, True)                          # model.py:94-101
    if debug: print("Shape %s ... Name %s" % (tuple(net.shape), "Final"))
    net = net.squeeze(-2)                                                            # model.py:104
    return net


def declare_variables(flags, num_channel: int, device) -> None:
    """Create every variable build() would create, without running it (the TF graph is built -- and its
    variables exist -- before the first batch: trainval.py:26-55, main_funcs.py:70-93 restore)."""
    st = default_store()
    L = int(flags.EDGE_CONV_LAYERS)
    filt = ops._listify(flags.EDGE_CONV_FILTERS, L, "num_filters")
    train = bool(flags.TRAIN)
    cin = int(num_channel)
    for i in range(L):
        with st.variable_scope("EdgeConv%d" % i):
            ops._conv_bn_vars("conv0", 2 * cin, int(filt[i]), train, device)
            ops._conv_bn_vars("conv1", 2 * int(filt[i]), ops.CONV1_WIDTH, train, device)
            if flags.MODEL_NAME != "dgcnn" and i > 0 and filt[i] != filt[i - 1]:
                ops._conv_bn_vars("shortcut", ops.CONV1_WIDTH, int(filt[i]), train, device)
        cin = ops.CONV1_WIDTH
    if flags.MODEL_NAME == "residual-dgcnn-nofc":
        ops._conv_bn_vars("Final", ops.CONV1_WIDTH, int(flags.NUM_CLASS), True, device)
        return
    if flags.MODEL_NAME not in ("dgcnn", "residual-dgcnn"):
        print("Unsupported MODEL_NAME: %s" % flags.MODEL_NAME)
        raise NotImplementedError
    ops._conv_bn_vars("MergedEdgeConv", ops.CONV1_WIDTH * L, 1024, True, device)
    width = 1024 + sum(2 * int(f) + ops.CONV1_WIDTH for f in filt) + 1024
    nfc = int(flags.FC_LAYERS)
    fcf = ops._listify(flags.FC_FILTERS, nfc, "num_filters")
    for j in range(nfc):
        ops._conv_bn_vars("FC%d" % j, width, int(fcf[j]), train, device)
        width = int(fcf[j])
    ops._conv_bn_vars("Final", width, int(flags.NUM_CLASS), True, device)
