"""MatrixMarket I/O (reference apps/mf/io.h:38-352: coordinate files for the data matrix, array files
for the factors)."""
from __future__ import annotations

import numpy as np


def read_matrix_market_coo(path: str):
    """(i, j, x, rows, cols) with 0-based int64 indices and float32 values; parsed by the native reader
    (csrc/adapm/io.cc), ``read_matrix_market_coo_py`` is the numpy reference."""
    from .. import _C

    return tuple(_C.read_matrix_market_coo(path))


def read_matrix_market_coo_py(path: str):
    with open(path) as f:
        header = f.readline()
        assert header.startswith("%%MatrixMarket matrix coordinate"), f"{path}: not a MatrixMarket coordinate file"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, nnz = (int(t) for t in line.split())
        d = np.loadtxt(f, dtype=np.float64).reshape(-1, 3)
    assert d.shape[0] == nnz
    return d[:, 0].astype(np.int64) - 1, d[:, 1].astype(np.int64) - 1, d[:, 2].astype(np.float32), m, n


def read_matrix_market_array(path: str) -> np.ndarray:
    with open(path) as f:
        header = f.readline()
        assert header.startswith("%%MatrixMarket matrix array"), f"{path}: not a MatrixMarket array file"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n = (int(t) for t in line.split())
        v = np.loadtxt(f, dtype=np.float64)
    return v.reshape(n, m).T.copy()  # column-major on disk
