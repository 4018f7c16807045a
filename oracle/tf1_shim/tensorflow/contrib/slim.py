"""tf.contrib.slim.conv2d / batch_norm with the defaults the reference relies on (ops.py:47-54 and every other conv).

slim.conv2d(inputs NHWC, num_outputs, kernel_size=1, stride=1, padding='VALID', normalizer_fn=slim.batch_norm,
            activation_fn=tf.nn.relu (default) | None, scope=..., trainable=...):
  * variable `<scope>/weights`, shape [1, 1, Cin, Cout], weights_initializer = xavier_initializer();
  * NO bias when a normalizer_fn is given ("biases ... ignored if normalizer_fn is not None");
  * output = activation_fn(normalizer_fn(conv(inputs))) -- the activation comes AFTER the normalisation.
slim.batch_norm defaults: decay 0.999, center=True (variable `<scope>/BatchNorm/beta`, zeros), scale=False (no gamma),
  epsilon=0.001, is_training=True -> normalises with the batch mean and the BIASED batch variance over every axis but the
  last (tf.nn.moments / fused batch norm); the moving averages it also maintains are never read by the reference.
A 1x1 / stride-1 / VALID convolution is a matmul over the channel axis.
"""
import torch

import tensorflow as tf


def batch_norm(x, scope_name):
    beta = tf._get_variable(scope_name + "/BatchNorm/beta", (x.shape[-1],), "zeros", x.dtype)
    dims = tuple(range(x.dim() - 1))
    mean = x.mean(dim=dims, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=dims, keepdim=True)
    inv = torch.rsqrt(var + 0.001)                    # tf.nn.batch_normalization: x * inv + (beta - mean * inv)
    return x * inv + (beta - mean * inv)


_DEFAULT = object()


def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=_DEFAULT, normalizer_fn=None,
           trainable=True, scope=None):
    assert int(kernel_size) == 1 and int(stride) == 1 and padding == "VALID" and normalizer_fn is batch_norm and scope
    if activation_fn is _DEFAULT:
        activation_fn = tf.nn.relu
    # the variable scope is a property of the CALL SITE (graph construction), not of the moment the op executes
    return _conv2d(inputs, int(num_outputs), activation_fn, tf._scope_name(scope))


@tf.dual
def _conv2d(inputs, num_outputs, activation_fn, name):
    cin = int(inputs.shape[-1])
    w = tf._get_variable(name + "/weights", (1, 1, cin, num_outputs), "xavier", inputs.dtype)
    out = torch.matmul(inputs, w[0, 0])
    out = batch_norm(out, name)
    if activation_fn is not None:
        out = activation_fn(out)
    tf.TRACE["conv2d"][name] = out
    return out
