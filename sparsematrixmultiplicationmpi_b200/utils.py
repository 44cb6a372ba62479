"""Host utilities of the path, same names and behaviour as the reference's utils.cpp.

  readMatrixMarketFile    utils.cpp:70-185   (text -> COO on the host, CSR assembly on the device)
  generateLargeFatVector  utils.cpp:193-209
  serialize / deserialize utils.cpp:216-253
  areMatricesEqual        utils.cpp:38-63
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .matrix import DeviceCSR, SparseMatrix, as_fat_vector


def parse_matrix_market(filename: str):
    """Text front end of readMatrixMarketFile (utils.cpp:70-153) -> (n_rows, n_cols, rows, cols, vals, symmetric).

    Header lines start with '%'; any of them containing "symmetric" marks the matrix symmetric (so
    "skew-symmetric" counts, with no sign flip), "pattern" gives every record the value 1.0 (:84-105).
    The first other line is "rows cols nnz" (:108-109); then nnz whitespace-separated records
    "r c [v]", 1-based (:124-153). Errors are the reference's messages (:77, :114, :140).
    """
    try:
        f = open(filename, "rb")
    except OSError:
        raise RuntimeError("Unable to open file: " + filename)
    with f:
        symmetric = pattern = False
        size_line = None
        for raw in f:
            if raw.startswith(b"%"):
                symmetric |= b"symmetric" in raw
                pattern |= b"pattern" in raw
            else:
                size_line = raw
                break
        try:
            n_rows, n_cols, nnz = (int(t) for t in size_line.split()[:3])
        except Exception:
            raise RuntimeError("Failed to read matrix dimensions from file: " + filename)
        tokens = f.read().split()
    per = 2 if pattern else 3
    if len(tokens) < per * nnz:
        raise RuntimeError("Failed to read data from file: " + filename)
    tokens = tokens[:per * nnz]
    try:
        rows = np.array(tokens[0::per], dtype=np.int64) - 1
        cols = np.array(tokens[1::per], dtype=np.int64) - 1
        vals = np.ones(nnz, dtype=np.float64) if pattern else np.array(tokens[2::per], dtype=np.float64)
    except ValueError:
        raise RuntimeError("Failed to read data from file: " + filename)
    return n_rows, n_cols, rows.astype(np.int32), cols.astype(np.int32), vals, symmetric


def readMatrixMarketFile(filename: str, device: int = 0, return_device: bool = False):
    """MatrixMarket coordinate file -> SparseMatrix, bit-identical to the reference loader's CSR.

    The records are parsed on the host; mirroring, the per-row (column, value) order and the row
    pointer are built in HBM by spmm_csr_from_coo_host. With return_device=True the resident
    DeviceCSR is returned alongside, so the multiply can reuse it without another upload.
    """
    n_rows, n_cols, rows, cols, vals, symmetric = parse_matrix_market(filename)
    dev = DeviceCSR.from_coo_host(n_rows, n_cols, rows, cols, vals, symmetric, device)
    host = dev.download()
    if return_device:
        return host, dev
    dev.close()
    return host


def generateLargeFatVector(n: int, m: int) -> np.ndarray:
    """n x m doubles, each rand() % 100 + 1 from libc's never-seeded generator (utils.cpp:193-209)."""
    out = np.empty((n, m), dtype=np.float64)
    _cabi.lib().spmm_generate_fat_vector(n, m, out.ctypes.data)
    return out


def serialize(fatVec) -> np.ndarray:
    """Row-major flatten (utils.cpp:216-228)."""
    return as_fat_vector(fatVec).reshape(-1).copy()


def deserialize(flat, rows: int, cols: int) -> np.ndarray:
    """Inverse of serialize (utils.cpp:237-253)."""
    return np.ascontiguousarray(flat, dtype=np.float64).reshape(rows, cols).copy()


def areMatricesEqual(mat1, mat2, tolerance: float) -> bool:
    """Same shape and every |a-b| <= tolerance, absolute (utils.cpp:38-63)."""
    a, b = as_fat_vector(mat1), as_fat_vector(mat2)
    if a.shape != b.shape:
        return False
    return bool(_cabi.lib().spmm_are_equal(a.ctypes.data, b.ctypes.data, a.size, C.c_double(tolerance)))
