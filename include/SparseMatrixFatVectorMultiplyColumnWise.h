// Entry point kept from the reference ("Source Code/SparseMatrixFatVectorMultiplyColumnWise.h":15).
// Column decomposition. The reference splits the k columns of B between ranks
// (ColumnWise.cpp:25-28); BASELINE.json's north_star asks for column blocks of A
// with the partial C summed. Both give the same C; see DESIGN.md for which one
// each layer implements. Rank 0 returns the full C, others return {}.
#ifndef SPARSEMATRIXFATVECTORMULTIPLYCOLUMNWIZE_H
#define SPARSEMATRIXFATVECTORMULTIPLYCOLUMNWIZE_H

#include "MatrixDefinitions.h"
#include <iostream>

FatVector sparseMatrixFatVectorMultiplyColumnWise(const SparseMatrix &sparseMatrix, const FatVector &fatVector, int vecCols);

#endif
