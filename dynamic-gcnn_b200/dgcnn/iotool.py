"""dgcnn.iotool -- IO handlers behind the reference's interface (/root/reference/dgcnn/iotool.py:6-33,282-287):
io_factory(flags) -> handler with initialize() / next() -> (idx, data, label, weight) / store(idx, softmax) /
finalize() / num_entries() / num_channels() / batch_size().

Out of scope for performance.  What exists here:
  io_array  : the dense fixed-N layout of io_h5 (data [n,N,C], label [n,N], optional weight [n,N], whole
              dataset in RAM, shuffle / sequential-wraparound batching: iotool.py:220-231,258-278) read from
              HDF5 files (.h5 / .hdf5: h5py when it is importable, else the built-in reader dgcnn.h5lite) or .npz
              files with the same keys; the inference output goes to an HDF5 file with the reference's layout
              (iotool.py:226-250: extendable zlib-5 EArrays DATA_KEY / softmax / LABEL_KEY, the label stored as
              float32 like the reference's Float32Atom) when OUTPUT_FILE ends in .h5 / .hdf5, else to .npz.  -io h5
  io_synth  : seeded in-memory clouds of the benchmark shape (uniform [0,1)^C points, labels in {0..NUM_CLASS-1})
              -io synthetic
  io_larcv  : needs larcv2 + ROOT (external HEP stack, not in this image) -> NotImplementedError on use.
The reference's truthiness bugs on numpy arrays (iotool.py:228-229,241,274-275) are fixed (`is not None`), and
the stored softmax has shape [N, NUM_CLASS] (not the data's shape, iotool.py:237-240).
"""
from __future__ import annotations

import numpy as np


class io_base(object):
    def __init__(self, flags):
        self._batch_size = flags.BATCH_SIZE
        self._num_entries = -1
        self._num_channels = -1

    def batch_size(self, size=None):
        if size is None:
            return self._batch_size
        self._batch_size = int(size)

    def num_entries(self):
        return self._num_entries

    def num_channels(self):
        return self._num_channels

    def initialize(self):
        raise NotImplementedError

    def store(self, idx, softmax):
        raise NotImplementedError

    def next(self):
        raise NotImplementedError

    def finalize(self):
        raise NotImplementedError


class io_array(io_base):
    """Dense dataset held in RAM (the io_h5 contract)."""

    def __init__(self, flags):
        super(io_array, self).__init__(flags=flags)
        self._flags = flags
        self._data = self._label = self._weight = None
        self._out = None

    def _read(self, path):
        f = self._flags
        if path.endswith((".h5", ".hdf5", ".hdf")):
            try:
                import h5py
                opener = lambda: h5py.File(path, "r")  # noqa: E731
            except ImportError:
                from . import h5lite
                opener = lambda: h5lite.File(path)  # noqa: E731
            with opener() as h:
                get = lambda key: np.array(h[key])  # noqa: E731
                return (get(f.DATA_KEY), get(f.LABEL_KEY) if f.LABEL_KEY else None,
                        get(f.WEIGHT_KEY) if f.WEIGHT_KEY else None)
        z = np.load(path, allow_pickle=False)
        return (np.asarray(z[f.DATA_KEY]), np.asarray(z[f.LABEL_KEY]) if f.LABEL_KEY and f.LABEL_KEY in z else None,
                np.asarray(z[f.WEIGHT_KEY]) if f.WEIGHT_KEY else None)

    def initialize(self):
        self._last_entry = -1
        parts = [self._read(p) for p in self._flags.INPUT_FILE]
        cat = lambda i: (None if parts[0][i] is None else np.concatenate([p[i] for p in parts], axis=0))  # noqa: E731
        self._data, self._label, self._weight = cat(0).astype(np.float32), cat(1), cat(2)
        self._num_channels = self._data.shape[-1]
        self._num_entries = len(self._data)
        if self._flags.OUTPUT_FILE:
            self._out = {"idx": [], "softmax": []}

    def store(self, idx, softmax):
        if self._out is None:
            raise NotImplementedError
        idx = int(idx)
        if idx >= self.num_entries():
            raise ValueError
        self._out["idx"].append(idx)
        self._out["softmax"].append(np.asarray(softmax, dtype=np.float32))

    def next(self):
        n, bs = self.num_entries(), self.batch_size()
        if self._flags.SHUFFLE:
            idx = np.arange(n)
            np.random.shuffle(idx)
            idx = idx[0:bs]
        else:
            idx = (np.arange(bs) + self._last_entry + 1) % n
        self._last_entry = int(idx[-1])
        data = self._data[idx, ...]
        label = self._label[idx, ...] if self._label is not None else None
        weight = self._weight[idx, ...] if self._weight is not None else None
        return idx, data, label, weight

    def finalize(self):
        if self._out is not None and self._out["idx"]:
            order = np.asarray(self._out["idx"])
            out = {self._flags.DATA_KEY: self._data[order], "softmax": np.stack(self._out["softmax"]), "index": order}
            if self._label is not None:
                out[self._flags.LABEL_KEY] = self._label[order]
            if str(self._flags.OUTPUT_FILE).endswith((".h5", ".hdf5", ".hdf")):
                from . import h5lite
                if self._label is not None:                       # iotool.py:235: create_earray(..., Float32Atom())
                    out[self._flags.LABEL_KEY] = out[self._flags.LABEL_KEY].astype(np.float32)
                h5lite.write(self._flags.OUTPUT_FILE, out, compress=5)
            else:
                np.savez_compressed(self._flags.OUTPUT_FILE, **out)


io_h5 = io_array  # the reference's name


class io_synth(io_array):
    """Synthetic clouds of the benchmark shape: INPUT_FILE is ignored."""

    N_ENTRIES = 64

    def initialize(self):
        f = self._flags
        self._last_entry = -1
        rng = np.random.RandomState(1234)
        n = max(self.N_ENTRIES, int(f.BATCH_SIZE))
        npts = int(f.NUM_POINT) if int(f.NUM_POINT) > 0 else 2048
        ch = int(f.NUM_CHANNEL) if int(getattr(f, "NUM_CHANNEL", -1)) > 0 else 3
        self._data = rng.random_sample((n, npts, ch)).astype(np.float32)
        self._label = rng.randint(0, int(f.NUM_CLASS), size=(n, npts)).astype(np.int32) if f.LABEL_KEY else None
        self._weight = np.ones((n, npts), np.float32) if getattr(f, "WEIGHT_KEY", "") else None
        self._num_channels, self._num_entries = ch, n
        if f.OUTPUT_FILE:
            self._out = {"idx": [], "softmax": []}


class io_larcv(io_base):
    """iotool.py:35-197 reads sparse3d ROOT trees through larcv2; neither library exists in this image."""

    def __init__(self, flags):
        super(io_larcv, self).__init__(flags=flags)

    def initialize(self):
        raise NotImplementedError("io_larcv needs larcv2 + ROOT, which are external to the reference and absent here")


def io_factory(flags):
    if flags.IO_TYPE == "h5":
        return io_h5(flags)
    if flags.IO_TYPE == "synthetic":
        return io_synth(flags)
    if flags.IO_TYPE == "larcv":
        return io_larcv(flags)
    raise NotImplementedError
