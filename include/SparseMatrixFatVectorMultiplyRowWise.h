// Entry point kept from the reference ("Source Code/SparseMatrixFatVectorMultiplyRowWise.h":15-17).
// Row-block decomposition: rank r owns the contiguous rows given by the
// reference formula (RowWise.cpp:26-29). Here each rank drives one B200 and
// multiplies only its row block; rank 0 returns the full C, others return {}.
#ifndef SPARSEMATRIXFATVECTORMULTIPLYROWWIZE_H
#define SPARSEMATRIXFATVECTORMULTIPLYROWWIZE_H

#include "MatrixDefinitions.h"
#include <iostream>

FatVector sparseMatrixFatVectorMultiplyRowWise(const SparseMatrix &sparseMatrix,
                                               const FatVector &fatVector,
                                               int vecCols);

#endif
