"""Native FASTA reader: ``read_fasta_file`` of karma.py:40-61 backed by
the C packer ``kb_fasta_*`` in libkarma_b200.so.

Returns the same ``OrderedDict{">name": sequence}`` karma.py builds, as a ``PackedSequences``
that also carries the packed buffers (bases, offsets, key_len), so that
``KmerClustering`` hands them to the GPU without re-joining a million Python strings.
"""
from collections import OrderedDict
from ctypes import byref, c_int64, c_void_p

import numpy as np

from . import _lib


class PackedSequences(OrderedDict):
    """OrderedDict of the records plus their packed form.  ``packed()`` returns
    (bases uint8, offsets int64, key_len int32) while the mapping is unmodified
    (same number of records), else None."""

    def __init__(self):
        super().__init__()
        self._packed = None
        self._n = -1

    def packed(self):
        if self._packed is not None and len(self) == self._n:
            return self._packed
        return None


def read_packed(path):
    """(bases, offsets, key_len, keys, key_offsets) straight from the C packer."""
    lib = _lib.load()
    h = c_void_p()
    n, nb, nk = c_int64(), c_int64(), c_int64()
    _lib.check(lib.kb_fasta_open(str(path).encode(), byref(h), byref(n), byref(nb), byref(nk)))
    try:
        bases = np.empty(nb.value, dtype=np.uint8)
        offsets = np.empty(n.value + 1, dtype=np.int64)
        key_len = np.empty(n.value, dtype=np.int32)
        keys = np.empty(nk.value, dtype=np.uint8)
        key_offsets = np.empty(n.value + 1, dtype=np.int64)
        _lib.check(lib.kb_fasta_fill(h, bases.ctypes.data_as(c_void_p), offsets.ctypes.data_as(c_void_p),
                                     key_len.ctypes.data_as(c_void_p), keys.ctypes.data_as(c_void_p),
                                     key_offsets.ctypes.data_as(c_void_p)))
    finally:
        lib.kb_fasta_close(h)
    return bases, offsets, key_len, keys, key_offsets


def read_fasta_file(fasta_file):
    """Drop-in for karma.py:40-61."""
    bases, offsets, key_len, keys, key_offsets = read_packed(fasta_file)
    raw = bases.tobytes().decode("ascii")
    rawk = keys.tobytes().decode("ascii")
    off = offsets.tolist()
    ko = key_offsets.tolist()
    d = PackedSequences()
    n = len(off) - 1
    for i in range(n):
        d[rawk[ko[i]:ko[i + 1]]] = raw[off[i]:off[i + 1]]     # duplicate keys: first position, last value
    if len(d) == n:                                           # no duplicates: the packed form is the mapping
        d._packed = (bases, offsets, key_len)
        d._n = n
    return d
