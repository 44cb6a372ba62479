// entry_points.cpp — the reference's four C++ entry points, bodies replaced by calls into the
// extern "C" layer (include/spmm_b200.h). Signatures, headers and the SparseMatrix / FatVector
// types are the reference's, so its main.cpp and utils.cpp drive this file unchanged:
//   sparseMatrixFatVectorMultiply                "Source Code/SparseMatrixFatVectorMultiply.h":14-15
//   sparseMatrixFatVectorMultiplyRowWise         "Source Code/SparseMatrixFatVectorMultiplyRowWise.h":15-17
//   sparseMatrixFatVectorMultiplyColumnWise      "Source Code/SparseMatrixFatVectorMultiplyColumnWise.h":15
//   sparseMatrixFatVectorMultiplyNonZeroElement  "Source Code/SparseMatrixFatVectorMultiplyNonZeroElement.h":15
//
// An MPI rank drives GPU (rank mod device count). Rank and size come from the caller's MPI
// (<mpi.h>: a real one, or include/compat/mpi.h in this image), and — because the result has to
// land in a host FatVector on rank 0 anyway — the result collective stays the reference's own
// MPI call on host buffers (Gatherv / Reduce). The NCCL collectives of the one-process-per-GPU
// layout live in the torch.distributed host layer (strategies.py); see INTEGRATION.md.
//
// Nothing here computes: a non-zero status from the C-ABI becomes std::runtime_error (the
// reference's only error convention, utils.cpp:77,114,140). No CPU fallback.
#include <mpi.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "MatrixDefinitions.h"
#include "SparseMatrixFatVectorMultiply.h"
#include "SparseMatrixFatVectorMultiplyColumnWise.h"
#include "SparseMatrixFatVectorMultiplyNonZeroElement.h"
#include "SparseMatrixFatVectorMultiplyRowWise.h"
#include "spmm_b200.h"

namespace
{

void ok(int status)
{
    if (status != SPMM_OK)
        throw std::runtime_error(std::string("spmm_b200: ") + spmm_last_error());
}

int device_for_rank(int rank)
{
    int count = 0;
    ok(spmm_device_count(&count));
    if (count < 1)
        throw std::runtime_error("spmm_b200: no CUDA device (no CPU fallback)");
    return rank % count;
}

// serialize() layout (utils.cpp:216-228): row-major flatten of the first n rows x k columns
std::vector<double> pack(const FatVector &v, size_t n, int k)
{
    if (v.size() < n)
        throw std::runtime_error("spmm_b200: fatVector has fewer rows than the matrix has columns");
    std::vector<double> flat(n * (size_t)k);
    for (size_t i = 0; i < n; ++i)
    {
        if (v[i].size() < (size_t)k)
            throw std::runtime_error("spmm_b200: fatVector row shorter than vecCols");
        std::memcpy(flat.data() + i * (size_t)k, v[i].data(), sizeof(double) * (size_t)k);
    }
    return flat;
}

FatVector unpack(const double *flat, size_t n, int k)
{
    FatVector out(n, std::vector<double>((size_t)k));
    for (size_t i = 0; i < n; ++i)
        std::memcpy(out[i].data(), flat + i * (size_t)k, sizeof(double) * (size_t)k);
    return out;
}

void validate(const SparseMatrix &m, int k)
{
    if (k < 0)
        throw std::runtime_error("spmm_b200: vecCols is negative");
    if (m.numRows < 0 || m.rowPtr.size() != (size_t)m.numRows + 1 || m.values.size() != m.colIndices.size() ||
        (size_t)m.rowPtr[m.numRows] != m.values.size())
        throw std::runtime_error("spmm_b200: SparseMatrix arrays are inconsistent");
}

// Device copies of matrix shards, keyed on the host buffers' identity, so main()'s four
// back-to-back calls on one M upload each shard once. Inputs are borrowed for the call only:
// the key also carries the sizes, and spmm_entry_clear_cache() drops everything.
struct Key
{
    const void *vals, *cols, *rowptr;
    size_t nnz;
    int n_rows, n_cols, tag, a, b;
    bool operator<(const Key &o) const
    {
        return std::tie(vals, cols, rowptr, nnz, n_rows, n_cols, tag, a, b) <
               std::tie(o.vals, o.cols, o.rowptr, o.nnz, o.n_rows, o.n_cols, o.tag, o.a, o.b);
    }
};
std::mutex g_mu;
std::map<Key, spmm_csr_t> g_cache;

template <typename Make>
spmm_csr_t cached(const SparseMatrix &m, int tag, int a, int b, Make make)
{
    Key key{m.values.data(), m.colIndices.data(), m.rowPtr.data(), m.values.size(), m.numRows, m.numCols, tag, a, b};
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end())
        return it->second;
    if (g_cache.size() >= 32)
    {
        for (auto &kv : g_cache)
            spmm_csr_destroy(kv.second);
        g_cache.clear();
    }
    spmm_csr_t h = make();
    g_cache[key] = h;
    return h;
}

spmm_csr_t whole_matrix(const SparseMatrix &m, int device)
{
    return cached(m, 0, device, 0, [&] {
        spmm_csr_t h = nullptr;
        ok(spmm_csr_create_host(device, m.numRows, m.numCols, (long long)m.values.size(), m.rowPtr.data(),
                                m.colIndices.data(), m.values.data(), &h));
        return h;
    });
}

} // namespace

extern "C" void spmm_entry_clear_cache()
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &kv : g_cache)
        spmm_csr_destroy(kv.second);
    g_cache.clear();
}

FatVector sparseMatrixFatVectorMultiply(const SparseMatrix &sparseMatrix, const FatVector &fatVector, int vecCols)
{
    validate(sparseMatrix, vecCols);
    const size_t n = (size_t)sparseMatrix.numRows;
    if (n == 0 || vecCols == 0)
        return FatVector(n, std::vector<double>((size_t)vecCols, 0.0));
    spmm_csr_t A = whole_matrix(sparseMatrix, device_for_rank(0));
    const std::vector<double> B = pack(fatVector, (size_t)sparseMatrix.numCols, vecCols);
    std::vector<double> C(n * (size_t)vecCols);
    ok(spmm_multiply_host(A, B.data(), vecCols, C.data(), SPMM_KERNEL_AUTO));
    return unpack(C.data(), n, vecCols);
}

FatVector sparseMatrixFatVectorMultiplyRowWise(const SparseMatrix &sparseMatrix, const FatVector &fatVector,
                                               int vecCols)
{
    validate(sparseMatrix, vecCols);
    int worldSize = 1, worldRank = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &worldSize);
    MPI_Comm_rank(MPI_COMM_WORLD, &worldRank);
    int start = 0, end = 0;
    spmm_partition_rows(sparseMatrix.numRows, worldSize, worldRank, &start, &end); // RowWise.cpp:26-29
    const int device = device_for_rank(worldRank);

    std::vector<double> local((size_t)(end - start) * (size_t)vecCols);
    if (end > start && vecCols > 0)
    {
        // the rank's shard: rows [start,end) with the row pointer rebased
        spmm_csr_t A = cached(sparseMatrix, 1, start, end, [&] {
            const int lo = sparseMatrix.rowPtr[start], hi = sparseMatrix.rowPtr[end];
            std::vector<int> rp(sparseMatrix.rowPtr.begin() + start, sparseMatrix.rowPtr.begin() + end + 1);
            for (int &x : rp)
                x -= lo;
            spmm_csr_t h = nullptr;
            ok(spmm_csr_create_host(device, end - start, sparseMatrix.numCols, (long long)hi - lo, rp.data(),
                                    sparseMatrix.colIndices.data() + lo, sparseMatrix.values.data() + lo, &h));
                return h;
        });
        const std::vector<double> B = pack(fatVector, (size_t)sparseMatrix.numCols, vecCols);
        ok(spmm_multiply_host(A, B.data(), vecCols, local.data(), SPMM_KERNEL_AUTO));
    }

    // Gatherv of the row blocks to rank 0 (RowWise.cpp:63-87)
    std::vector<int> counts(worldSize), displs(worldSize);
    for (int r = 0, off = 0; r < worldSize; ++r)
    {
        int s, e;
        spmm_partition_rows(sparseMatrix.numRows, worldSize, r, &s, &e);
        counts[r] = (e - s) * vecCols;
        displs[r] = off;
        off += counts[r];
    }
    std::vector<double> gathered;
    if (worldRank == 0)
        gathered.resize((size_t)sparseMatrix.numRows * (size_t)vecCols + 1);
    MPI_Gatherv(local.data(), (int)local.size(), MPI_DOUBLE, gathered.data(), counts.data(), displs.data(),
                MPI_DOUBLE, 0, MPI_COMM_WORLD);
    if (worldRank != 0)
        return FatVector{}; // RowWise.cpp:125
    return unpack(gathered.data(), (size_t)sparseMatrix.numRows, vecCols);
}

FatVector sparseMatrixFatVectorMultiplyColumnWise(const SparseMatrix &sparseMatrix, const FatVector &fatVector,
                                                  int vecCols)
{
    // Column blocks of A (BASELINE.json's reading): rank r owns columns J_r of A and
    // rows J_r of B, produces a full-size partial C, and the partials are summed to rank 0.
    validate(sparseMatrix, vecCols);
    int worldSize = 1, worldRank = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &worldSize);
    MPI_Comm_rank(MPI_COMM_WORLD, &worldRank);
    int c0 = 0, c1 = 0;
    spmm_partition_rows(sparseMatrix.numCols, worldSize, worldRank, &c0, &c1);
    const int device = device_for_rank(worldRank);
    const size_t n = (size_t)sparseMatrix.numRows;

    std::vector<double> partial(n * (size_t)vecCols + 1, 0.0);
    if (n && vecCols > 0)
    {
        spmm_csr_t A = cached(sparseMatrix, 2, c0, c1, [&] {
            spmm_csr_t whole = whole_matrix(sparseMatrix, device), h = nullptr;
            ok(spmm_csr_column_block(whole, c0, c1, &h));
            return h;
        });
        if (fatVector.size() < (size_t)sparseMatrix.numCols)
            throw std::runtime_error("spmm_b200: fatVector has fewer rows than the matrix has columns");
        const FatVector slab(fatVector.begin() + c0, fatVector.begin() + c1);
        const std::vector<double> B = pack(slab, (size_t)(c1 - c0), vecCols);
        ok(spmm_multiply_host(A, B.data(), vecCols, partial.data(), SPMM_KERNEL_AUTO));
    }
    std::vector<double> total;
    if (worldRank == 0)
        total.resize(n * (size_t)vecCols + 1);
    MPI_Reduce(partial.data(), total.data(), (int)(n * (size_t)vecCols), MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
    if (worldRank != 0)
        return FatVector{};
    return unpack(total.data(), n, vecCols);
}

FatVector sparseMatrixFatVectorMultiplyNonZeroElement(const SparseMatrix &sparseMatrix, const FatVector &fatVector,
                                                      int vecCols)
{
    validate(sparseMatrix, vecCols);
    int worldSize = 1, worldRank = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &worldSize);
    MPI_Comm_rank(MPI_COMM_WORLD, &worldRank);
    long long b = 0, e = 0;
    spmm_partition_nnz((long long)sparseMatrix.values.size(), worldSize, worldRank, &b, &e); // NonZeroElement.cpp:24-39
    const int device = device_for_rank(worldRank);
    const size_t n = (size_t)sparseMatrix.numRows;

    // full-size zeroed accumulator as in the reference (:54); only the rank's rows are written
    std::vector<double> local(n * (size_t)vecCols + 1, 0.0);
    if (e > b && vecCols > 0)
    {
        spmm_csr_t A = whole_matrix(sparseMatrix, device);
        int first = 0, last = -1;
        ok(spmm_nnz_range_rows(A, b, e, &first, &last));
        const std::vector<double> B = pack(fatVector, (size_t)sparseMatrix.numCols, vecCols);
        ok(spmm_multiply_nnz_range_host(A, b, e, first, last, B.data(), vecCols,
                                        local.data() + (size_t)first * (size_t)vecCols, SPMM_KERNEL_AUTO));
    }
    std::vector<double> total;
    if (worldRank == 0)
        total.resize(n * (size_t)vecCols + 1);
    MPI_Reduce(local.data(), total.data(), (int)(n * (size_t)vecCols), MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD); // :88
    if (worldRank != 0)
        return FatVector{};
    return unpack(total.data(), n, vecCols);
}
