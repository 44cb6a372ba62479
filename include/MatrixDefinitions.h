// Data model of the drop-in boundary: CSR matrix + "fat vector" (dense N x k).
//
// Mirrors the reference's types so its main.cpp / utils.cpp compile against
// this header unchanged:
//   struct SparseMatrix  <- /root/reference "Source Code/MatrixDefinitions.h":14-19
//   typedef FatVector    <- /root/reference "Source Code/MatrixDefinitions.h":22
// The shipped reference header is stale: every reference .cpp reads
// sparseMatrix.numRows / numCols (utils.cpp:180-181, main.cpp:60,111-112,
// SparseMatrixFatVectorMultiply.cpp:15) but the struct lacks them. The two
// ints below are that fix; they sit after the three vectors so aggregate
// initialisation of {values, colIndices, rowPtr} keeps working.
//
// The include guard is deliberately the reference's own so that force-including
// this file (-include) pre-empts the stale header when reference sources are
// compiled where they lie (oracle/Makefile).
#ifndef MATRIXDEFINITIONS_H
#define MATRIXDEFINITIONS_H

#include <vector>

struct SparseMatrix
{
    std::vector<double> values;  // nnz FP64 values, row by row
    std::vector<int> colIndices; // nnz 0-based column ids, ascending inside a row, duplicates allowed
    std::vector<int> rowPtr;     // numRows+1 offsets into the two arrays above
    int numRows = 0;
    int numCols = 0;
};

// N rows, each an independently allocated vector of k doubles (not contiguous).
typedef std::vector<std::vector<double>> FatVector;

#endif
