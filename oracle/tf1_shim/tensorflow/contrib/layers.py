"""`import tensorflow.contrib.layers as L` (ops.py:6): imported by the reference, never used on the path."""
