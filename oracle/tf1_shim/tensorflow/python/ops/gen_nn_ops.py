"""gen_nn_ops.max_pool_v2 (model.py:76-77): NHWC max pooling with a run-time ksize; 'VALID' padding, unit strides."""
import torch

import tensorflow as tf


@tf.dual
def max_pool_v2(x, ksize, strides, padding, name=None):
    assert padding == "VALID" and list(strides) == [1, 1, 1, 1] and int(ksize[0]) == 1 and int(ksize[3]) == 1
    kh, kw = int(ksize[1]), int(ksize[2])
    y = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), kernel_size=(kh, kw), stride=1)
    return y.permute(0, 2, 3, 1)
