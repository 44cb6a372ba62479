// ref_wrap.cpp — extern "C" doors into the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/spmm_oracle.c header). This file is
// linked with the reference's own .cpp files compiled from where they lie under
// /root/reference (oracle/Makefile); it only converts flat arrays to the
// reference's SparseMatrix / FatVector and back, and starts the rank-threads of
// compat/mpi.h for the three MPI strategies. Outputs land in oracle/_ref/.
#include <cstring>
#include <exception>
#include <string>
#include <thread>

#include "utils.h" // the reference's own header (pulls compat mpi.h / petsc.h)
#include "SparseMatrixFatVectorMultiply.h"
#include "SparseMatrixFatVectorMultiplyRowWise.h"
#include "SparseMatrixFatVectorMultiplyColumnWise.h"
#include "SparseMatrixFatVectorMultiplyNonZeroElement.h"

namespace
{
thread_local std::string g_err;

SparseMatrix make_matrix(int n_rows, int n_cols, const int *rowptr, const int *colidx, const double *vals)
{
    SparseMatrix m;
    const int nnz = rowptr[n_rows];
    m.values.assign(vals, vals + nnz);
    m.colIndices.assign(colidx, colidx + nnz);
    m.rowPtr.assign(rowptr, rowptr + n_rows + 1);
    m.numRows = n_rows;
    m.numCols = n_cols;
    return m;
}

FatVector make_fat(const double *flat, int n, int k)
{
    FatVector v(n, std::vector<double>(k));
    for (int i = 0; i < n; ++i)
        std::memcpy(v[i].data(), flat + (size_t)i * k, sizeof(double) * k);
    return v;
}

void flatten(const FatVector &v, int k, double *out)
{
    for (size_t i = 0; i < v.size(); ++i)
        std::memcpy(out + i * (size_t)k, v[i].data(), sizeof(double) * k);
}
} // namespace

extern "C"
{

const char *ref_last_error() { return g_err.c_str(); }

// strategy: 0 sequential, 1 row-wise, 2 column-wise, 3 non-zero-element.
// P rank-threads for 1..3. seconds (optional) = wall time of the multiply call
// alone on rank 0, measured like main.cpp:77-79,161-163 (conversion excluded).
int ref_spmm(int strategy, int P, int n_rows, int n_cols, const int *rowptr, const int *colidx,
             const double *vals, const double *B, int k, double *C, double *seconds)
{
    try
    {
        const SparseMatrix M = make_matrix(n_rows, n_cols, rowptr, colidx, vals);
        const FatVector v = make_fat(B, n_cols, k);
        FatVector out;
        double dt = 0.0;
        if (strategy == 0)
        {
            double t0 = MPI_Wtime();
            out = sparseMatrixFatVectorMultiply(M, v, k);
            dt = MPI_Wtime() - t0;
        }
        else
        {
            compat_mpi::run(P, [&](int rank) {
                MPI_Barrier(MPI_COMM_WORLD);
                double t0 = MPI_Wtime();
                FatVector r = strategy == 1   ? sparseMatrixFatVectorMultiplyRowWise(M, v, k)
                              : strategy == 2 ? sparseMatrixFatVectorMultiplyColumnWise(M, v, k)
                                              : sparseMatrixFatVectorMultiplyNonZeroElement(M, v, k);
                double t1 = MPI_Wtime();
                if (rank == 0)
                {
                    out = std::move(r);
                    dt = t1 - t0;
                }
            });
        }
        if (C)
            flatten(out, k, C);
        if (seconds)
            *seconds = dt;
        return 0;
    }
    catch (const std::exception &e)
    {
        g_err = e.what();
        return 1;
    }
}

// readMatrixMarketFile (utils.cpp:70-185). Two-call protocol: arrays NULL -> sizes only.
int ref_read_mtx(const char *path, int *n_rows, int *n_cols, int *nnz,
                 int *rowptr, int *colidx, double *vals)
{
    try
    {
        SparseMatrix m = readMatrixMarketFile(path);
        *n_rows = m.numRows;
        *n_cols = m.numCols;
        *nnz = (int)m.values.size();
        if (rowptr && colidx && vals)
        {
            std::memcpy(rowptr, m.rowPtr.data(), sizeof(int) * m.rowPtr.size());
            std::memcpy(colidx, m.colIndices.data(), sizeof(int) * m.colIndices.size());
            std::memcpy(vals, m.values.data(), sizeof(double) * m.values.size());
        }
        return 0;
    }
    catch (const std::exception &e)
    {
        g_err = e.what();
        return 1;
    }
}

// generateLargeFatVector (utils.cpp:193-209); srand(1) = the never-seeded state.
void ref_generate_fatvector(int n, int k, double *out)
{
    srand(1);
    FatVector v = generateLargeFatVector(n, k);
    flatten(v, k, out);
}

// serialize / deserialize round trip (utils.cpp:216-253): returns 1 when the flat
// image equals the row-major layout this repo uses at the C-ABI.
int ref_serialize_is_rowmajor(const double *flat, int n, int k)
{
    FatVector v = deserialize(std::vector<double>(flat, flat + (size_t)n * k), n, k);
    std::vector<double> s = serialize(v);
    return s.size() == (size_t)n * k && std::memcmp(s.data(), flat, sizeof(double) * s.size()) == 0;
}

int ref_are_equal(const double *a, const double *b, int n, int k, double tol)
{
    return areMatricesEqual(make_fat(a, n, k), make_fat(b, n, k), tol) ? 1 : 0;
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
}
