"""oracle/h5_shims/tables -- TEST INFRASTRUCTURE ONLY.  The PyTables calls of the reference's io_h5 output path
(/root/reference/dgcnn/iotool.py:226-250: Filters, open_file, create_earray, Float32Atom, EArray.append, close), buffered in
memory and written by dgcnn.h5lite on close() as the extendable zlib EArrays PyTables would have written."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _h5lite import h5lite  # noqa: E402


class Filters(object):
    def __init__(self, complib="zlib", complevel=0):
        assert complib == "zlib"
        self.complevel = int(complevel)


class Float32Atom(object):
    dtype = np.float32


class _EArray(object):
    def __init__(self, atom, shape):
        assert shape[0] == 0, "the extendable axis is the first one"
        self.dtype, self.row_shape, self.rows = atom.dtype, tuple(int(s) for s in shape[1:]), []

    def append(self, arr):
        arr = np.asarray(arr)
        if tuple(arr.shape[1:]) != self.row_shape:        # PyTables refuses rows of another shape
            raise ValueError("the appended object has shape %s, the EArray rows have shape %s" % (arr.shape[1:], self.row_shape))
        self.rows.append(arr.astype(self.dtype))


class _File(object):
    def __init__(self, path, filters):
        self.path, self.filters, self.root, self.arrays = path, filters, "/", {}

    def create_earray(self, where, name, atom, shape):
        assert where == "/"
        ea = self.arrays[name] = _EArray(atom, list(shape))
        return ea

    def close(self):
        out = {n: (np.concatenate(a.rows, axis=0) if a.rows else np.zeros((0,) + a.row_shape, a.dtype))
               for n, a in self.arrays.items()}
        h5lite.write(self.path, out, compress=self.filters.complevel if self.filters else 0)


def open_file(path, mode="r", filters=None):
    assert mode == "w"
    return _File(path, filters)
