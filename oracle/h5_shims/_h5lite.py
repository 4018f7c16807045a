"""Loads dynamic-gcnn_b200/dgcnn/h5lite.py by path (it depends on numpy only) without importing the product package."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dynamic-gcnn_b200", "dgcnn",
                     "h5lite.py")
_spec = importlib.util.spec_from_file_location("_dgcnn_h5lite", _path)
h5lite = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(h5lite)
