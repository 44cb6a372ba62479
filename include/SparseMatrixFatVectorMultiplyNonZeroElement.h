// Entry point kept from the reference ("Source Code/SparseMatrixFatVectorMultiplyNonZeroElement.h":15).
// Equal non-zero ranges per rank (NonZeroElement.cpp:24-39). Here each rank runs
// the nnz-balanced merge-path kernel on its range; rows cut by a range boundary
// are summed in rank order. Rank 0 returns the full C, others return {}.
#ifndef SPARSEMATRIXFATVECTORMULTIPLYNONZEROELEMENT_H
#define SPARSEMATRIXFATVECTORMULTIPLYNONZEROELEMENT_H

#include "MatrixDefinitions.h"
#include <iostream>

FatVector sparseMatrixFatVectorMultiplyNonZeroElement(const SparseMatrix &sparseMatrix, const FatVector &fatVector, int vecCols);

#endif
