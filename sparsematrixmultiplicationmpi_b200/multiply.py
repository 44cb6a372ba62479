"""The four entry points of the reference, same names, argument meaning and result.

  sparseMatrixFatVectorMultiply                 SparseMatrixFatVectorMultiply.h:14-15, .cpp:11-31
  sparseMatrixFatVectorMultiplyRowWise          SparseMatrixFatVectorMultiplyRowWise.h:15-17, .cpp:12-126
  sparseMatrixFatVectorMultiplyColumnWise       SparseMatrixFatVectorMultiplyColumnWise.h:15, .cpp:13-131
  sparseMatrixFatVectorMultiplyNonZeroElement   SparseMatrixFatVectorMultiplyNonZeroElement.h:15, .cpp:12-120

Inputs are host objects (SparseMatrix, FatVector = (N,k) float64 array), replicated on every
rank as main.cpp:106-146 leaves them; the result is a host FatVector on rank 0 and an empty one
elsewhere (RowWise.cpp:125). A "rank" is one process driving one B200 (torchrun); without an
initialised process group there is one rank. All arithmetic runs in the sm_100a kernels; a
missing extension or GPU raises (no CPU fallback).

The device copy of a matrix shard is cached per (matrix buffers, strategy, rank layout), so
main()'s four back-to-back calls on the same M upload A once per strategy shard. The reference
checks nothing about its arguments; here sizes are validated and errors surface as RuntimeError
(the reference's only convention: std::runtime_error).
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from .matrix import DeviceCSR, SparseMatrix, as_fat_vector
from .strategies import ColumnBlocks, ColumnSlabs, CudaCompute, NonZeroRanges, RowWise, world

_compute: CudaCompute | None = None
_cache: dict = {}
_CACHE_MAX = 8


def _engine() -> CudaCompute:
    global _compute
    if _compute is None:
        _compute = CudaCompute()
    return _compute


def _close(plan) -> None:
    A = getattr(plan, "A", plan)
    if isinstance(A, DeviceCSR):
        A.close()


def clear_cache() -> None:
    """Drop every cached device shard."""
    for _, _, plan in _cache.values():
        _close(plan)
    _cache.clear()


def _fingerprint(m: SparseMatrix) -> int:
    """Cheap content check of a cache hit (the reference functions are pure: a matrix edited in place, or a new one
    at a recycled address, must not meet the old shard): every array whole up to 4 Ki elements, else 4 Ki evenly
    spaced probes."""
    h = 0
    for a in (m.rowPtr, m.colIndices, m.values):
        step = 1 if a.size <= (1 << 12) else a.size // (1 << 12)
        h = hash((h, a[::step].tobytes(), a[-1:].tobytes()))
    return h


def _key(m: SparseMatrix, tag: str, k: int | None):
    rank, P = world()
    return (m.values.ctypes.data, m.colIndices.ctypes.data, m.rowPtr.ctypes.data, m.nnz, m.numRows, m.numCols,
            tag, k, rank, P)


def _evict(key) -> None:
    entry = _cache.pop(key, None)
    if entry is not None:
        _close(entry[2])


def _cached(m: SparseMatrix, tag: str, k: int | None, make):
    """Device shard of (matrix, strategy, rank layout). An entry lives as long as its SparseMatrix (weakref.finalize
    evicts it) and is rebuilt when the matrix's contents no longer match the fingerprint taken at upload."""
    key = _key(m, tag, k)
    print_now = _fingerprint(m)
    hit = _cache.get(key)
    if hit is not None and hit[0]() is m and hit[1] == print_now:
        return hit[2]
    _evict(key)
    if len(_cache) >= _CACHE_MAX:
        clear_cache()
    plan = make()
    _cache[key] = (weakref.ref(m), print_now, plan)
    weakref.finalize(m, _evict, key)
    return plan


def _check(m: SparseMatrix, v, k: int) -> np.ndarray:
    if k < 0:
        raise RuntimeError("vecCols must be non-negative")
    if m.rowPtr.size != m.numRows + 1 or m.values.size != m.colIndices.size:
        raise RuntimeError("SparseMatrix arrays are inconsistent (rowPtr needs numRows+1 entries)")
    B = as_fat_vector(v)
    if B.shape[0] < m.numCols or B.shape[1] < k:
        raise RuntimeError(f"fatVector must hold at least {m.numCols} rows of {k} values")
    if B.shape[1] != k or B.shape[0] != m.numCols:
        B = np.ascontiguousarray(B[:m.numCols, :k])
    return B


def _to_device(B: np.ndarray, eng: CudaCompute, plan) -> torch.Tensor:
    """Host fat vector -> device tensor through the shard's pinned staging and the library's host threads."""
    out = torch.empty(B.shape, dtype=torch.float64, device=eng.device)
    if B.size:
        plan.A.upload_dense(B, out.data_ptr(), torch.cuda.current_stream(eng.device).cuda_stream)
    return out


def _to_host(Cd: torch.Tensor | None, k: int, eng: CudaCompute, plan) -> np.ndarray:
    if Cd is None:
        return np.empty((0, k), dtype=np.float64)  # FatVector{} on non-root ranks
    if Cd.numel() == 0:
        return np.empty(tuple(Cd.shape), dtype=np.float64)
    Cd = Cd.contiguous()
    return plan.A.download_dense(Cd.data_ptr(), Cd.shape[0], Cd.shape[1], torch.cuda.current_stream(eng.device).cuda_stream)


def sparseMatrixFatVectorMultiply(sparseMatrix: SparseMatrix, fatVector, vecCols: int, out=None) -> np.ndarray:
    """C = A * B on one B200 (the sequential reference function; no communication).
    `out` (optional, not in the reference signature) reuses a caller-owned result buffer."""
    B = _check(sparseMatrix, fatVector, vecCols)
    eng = _engine()
    A = _cached(sparseMatrix, "seq", None, lambda: eng.upload(sparseMatrix))
    return A.multiply_host(B, vecCols, eng.kernel, out)


def sparseMatrixFatVectorMultiplyRowWise(sparseMatrix: SparseMatrix, fatVector, vecCols: int) -> np.ndarray:
    B = _check(sparseMatrix, fatVector, vecCols)
    eng = _engine()
    plan = _cached(sparseMatrix, "row", vecCols, lambda: RowWise.from_host(eng, sparseMatrix, vecCols))
    return _to_host(plan.run(_to_device(B, eng, plan)), vecCols, eng, plan)


def sparseMatrixFatVectorMultiplyColumnWise(sparseMatrix: SparseMatrix, fatVector, vecCols: int,
                                            mode: str = "blocks") -> np.ndarray:
    """mode="blocks": column blocks of A + reduce-scatter (BASELINE.json); mode="slabs": the reference's
    split of B's k columns. Same C either way (up to FP64 summation order for "blocks")."""
    B = _check(sparseMatrix, fatVector, vecCols)
    eng = _engine()
    if mode == "slabs":
        plan = _cached(sparseMatrix, "colslab", vecCols, lambda: ColumnSlabs.from_host(eng, sparseMatrix, vecCols))
        return _to_host(plan.run(_to_device(B, eng, plan)), vecCols, eng, plan)
    if mode != "blocks":
        raise RuntimeError("mode must be 'blocks' or 'slabs'")
    plan = _cached(sparseMatrix, "colblk", vecCols, lambda: ColumnBlocks.from_host(eng, sparseMatrix, vecCols))
    B_local = _to_device(np.ascontiguousarray(B[plan.col_start:plan.col_end]), eng, plan)
    return _to_host(plan.run(B_local), vecCols, eng, plan)


def sparseMatrixFatVectorMultiplyNonZeroElement(sparseMatrix: SparseMatrix, fatVector, vecCols: int) -> np.ndarray:
    B = _check(sparseMatrix, fatVector, vecCols)
    eng = _engine()
    plan = _cached(sparseMatrix, "nnz", vecCols, lambda: NonZeroRanges.from_host(eng, sparseMatrix, vecCols))
    return _to_host(plan.run(_to_device(B, eng, plan)), vecCols, eng, plan)
