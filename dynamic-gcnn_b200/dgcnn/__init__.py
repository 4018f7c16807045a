"""dgcnn -- B200-native drop-in for the Python API of DeepLearnPhysics/dynamic-gcnn
(/root/reference/dgcnn/__init__.py:1-5 exports the same five names)."""
from .iotool import io_factory
from .model import build
from .trainval import trainval
from . import ops
from .flags import DGCNN_FLAGS

__all__ = ["io_factory", "build", "trainval", "ops", "DGCNN_FLAGS"]
