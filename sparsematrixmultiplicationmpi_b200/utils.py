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

    Native tokenizer (csrc/mm_read.cu, spmm_mm_read): header lines start with '%'; any of them containing
    "symmetric" marks the matrix symmetric (so "skew-symmetric" counts, with no sign flip), "pattern" gives every
    record the value 1.0 (:84-105). The first other line is "rows cols nnz" (:108-109); then nnz records
    "r c [v]", 1-based, as a whitespace-separated token stream (:124-153). Errors are the reference's messages
    (:77, :114, :140), raised as RuntimeError.
    """
    L = _cabi.lib()
    nr, nc, sym, ne = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    pr, pc, pv = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _cabi.check(L.spmm_mm_read(filename.encode(), C.byref(nr), C.byref(nc), C.byref(ne), C.byref(sym), C.byref(pr),
                               C.byref(pc), C.byref(pv)))
    n = ne.value
    try:
        if n and pr.value:
            rows = np.ctypeslib.as_array(C.cast(pr, C.POINTER(C.c_int)), shape=(n,)).copy()
            cols = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_int)), shape=(n,)).copy()
            vals = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(n,)).copy()
        else:
            rows, cols, vals = np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.float64)
    finally:
        L.spmm_mm_free(pr, pc, pv)
    return nr.value, nc.value, rows, cols, vals, bool(sym.value)


def readMatrixMarketFile(filename: str, device: int = 0, return_device: bool = False):
    """MatrixMarket coordinate file -> SparseMatrix, bit-identical to the reference loader's CSR.

    The records are tokenized on the host by all its threads (spmm_csr_from_matrix_market); mirroring, the per-row (column, value) order and the row
    pointer are built in HBM by spmm_csr_from_coo_host. With return_device=True the resident
    DeviceCSR is returned alongside, so the multiply can reuse it without another upload.
    """
    h = C.c_void_p()
    _cabi.check(_cabi.lib().spmm_csr_from_matrix_market(device, filename.encode(), C.byref(h)))
    dev = DeviceCSR(h.value)
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
