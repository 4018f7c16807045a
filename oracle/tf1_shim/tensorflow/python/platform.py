"""`import tensorflow.python.platform` (model.py:4) has no effect in TF1 beyond importing the module."""
