// Entry point kept from the reference ("Source Code/SparseMatrixFatVectorMultiply.h":14-15).
// C = A * B with A in CSR (SparseMatrix) and B a FatVector of vecCols columns.
// Reference body: sequential triple loop on the host. Here: one B200, the
// row-length-binned sm_100a kernels behind spmm_b200.h (no CPU fallback).
#ifndef SPARSEMATRIXFATVECTORMULTIPLY_H
#define SPARSEMATRIXFATVECTORMULTIPLY_H

#include "MatrixDefinitions.h"

FatVector sparseMatrixFatVectorMultiply(const SparseMatrix &sparseMatrix,
                                        const FatVector &fatVector, int vecCols);

#endif
