"""B200-native sparse matrix x fat vector multiply (C = A*B, A CSR FP64/int32, B dense N x k FP64).

Drop-in for the one hot path of AlexisBalayre/SparseMatrixMultiplicationMPI: the four multiply
entry points and the utils.cpp helpers keep their names. Everything computes in hand-written
sm_100a CUDA kernels behind the C-ABI of include/spmm_b200.h (libspmm_b200.so, built in-tree by
`make -C sparsematrixmultiplicationmpi_b200/csrc`). There is no CPU fallback.
"""
from . import _cabi
from .matrix import DeviceCSR, SparseMatrix, as_fat_vector
from .multiply import (clear_cache, sparseMatrixFatVectorMultiply, sparseMatrixFatVectorMultiplyColumnWise,
                       sparseMatrixFatVectorMultiplyNonZeroElement, sparseMatrixFatVectorMultiplyRowWise)
from .strategies import (ColumnBlocks, ColumnSlabs, CudaCompute, NonZeroRanges, RowWise, partition_cols,
                         partition_nnz, partition_rows)
from .utils import (areMatricesEqual, deserialize, generateLargeFatVector, parse_matrix_market,
                    readMatrixMarketFile, serialize)

__all__ = [
    "SparseMatrix", "DeviceCSR", "as_fat_vector", "CudaCompute",
    "sparseMatrixFatVectorMultiply", "sparseMatrixFatVectorMultiplyRowWise",
    "sparseMatrixFatVectorMultiplyColumnWise", "sparseMatrixFatVectorMultiplyNonZeroElement",
    "RowWise", "ColumnBlocks", "ColumnSlabs", "NonZeroRanges",
    "partition_rows", "partition_cols", "partition_nnz",
    "readMatrixMarketFile", "parse_matrix_market", "generateLargeFatVector", "serialize", "deserialize",
    "areMatricesEqual", "clear_cache",
]
