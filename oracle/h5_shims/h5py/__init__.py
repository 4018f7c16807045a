"""oracle/h5_shims/h5py -- TEST INFRASTRUCTURE ONLY.  The one h5py entry point the reference's io_h5 uses
(/root/reference/dgcnn/iotool.py:216-224: `h5.File(f, 'r')[key]` wrapped in np.array), served by dgcnn.h5lite, so that the
reference's own IO handler can be executed unmodified on HDF5 files in an image without h5py / libhdf5."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _h5lite import h5lite  # noqa: E402


def File(path, mode="r"):
    return h5lite.File(path, mode)
