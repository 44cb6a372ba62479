// spmm_stream.cu — narrow fat vectors (k = 1, 2, 4, 8): the CSR arrays are streamed in nnz order, independent of the
// row structure, and the rows are reduced from shared memory.
//
// Replaces, for k <= 8, the loop nest of
//   /root/reference "Source Code/SparseMatrixFatVectorMultiply.cpp":17-28
// Why: with few columns the row kernels are latency bound — a row of the cop20k_A shape is 22 non-zeros, one step of a
// warp, behind a chain of dependent loads (row pointer -> column/value -> B row) — and reach 20-30 % of the HBM
// bandwidth although the multiply is a pure stream of A (12 bytes per non-zero; B and C are small).
// Here a CTA takes a tile of consecutive rows holding at most TNZ non-zeros (tile cuts found once per matrix by binary
// search on rowptr and cached on the handle). Phase 1: every thread loads its non-zeros fully coalesced, gathers the B
// piece and leaves the products in shared memory — all loads of the tile are independent. Phase 2: one thread per
// (row, column) adds the row's products in ascending non-zero order. That is the reference's own arithmetic, product
// rounded, then added in column order (no FMA), so the result is bit-identical to the oracle's.
#include <algorithm>
#include <mutex>

#include "spmm_launch.cuh"

namespace spmm
{

namespace
{
constexpr int ST_TPB = 256;

__global__ void stream_cuts_kernel(const int *__restrict__ rowptr, int n_rows, int n_tiles, int tile_nnz, int *__restrict__ cuts)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles)
        return;
    if (t == n_tiles)
    {
        cuts[t] = n_rows;
        return;
    }
    // first row r with rowptr[r] >= t * tile_nnz
    const long long want = (long long)t * tile_nnz;
    int lo = 0, hi = n_rows;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (rowptr[mid] < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    cuts[t] = lo;
}

// K columns; LPE lanes share a non-zero, each holding W = K / LPE columns of its product
template <int K>
__global__ void __launch_bounds__(ST_TPB)
    spmm_stream_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                       const double *__restrict__ B, long long ldb, double *__restrict__ C, long long ldc,
                       const int *__restrict__ cuts, int cap)
{
    constexpr int LPE = K >= 4 ? K / 2 : 1;
    constexpr int W = K / LPE;
    extern __shared__ __align__(16) double prod[]; // cap x K
    const int r0 = cuts[blockIdx.x], r1 = cuts[blockIdx.x + 1];
    if (r0 >= r1)
        return;
    const int e0 = rowptr[r0], e1 = rowptr[r1];
    const int n = min(e1 - e0, cap); // the host guarantees e1 - e0 <= cap
    const int part = threadIdx.x % LPE;
    for (int i = threadIdx.x / LPE; i < n; i += ST_TPB / LPE)
    {
        const int c = __ldg(colidx + e0 + i);
        const double v = __ldg(vals + e0 + i);
        const double *b = B + (long long)c * ldb + part * W;
        if constexpr (W == 2)
        {
            const double2 bb = __ldg(reinterpret_cast<const double2 *>(b));
            *reinterpret_cast<double2 *>(prod + (size_t)i * K + part * 2) = make_double2(__dmul_rn(v, bb.x), __dmul_rn(v, bb.y));
        }
        else
            prod[(size_t)i * K + part] = __dmul_rn(v, __ldg(b));
    }
    __syncthreads();
    const int items = (r1 - r0) * K;
    for (int idx = threadIdx.x; idx < items; idx += ST_TPB)
    {
        const int r = r0 + idx / K, j = idx % K;
        const int a = rowptr[r] - e0, z = rowptr[r + 1] - e0;
        double sum = 0.0;
        for (int q = a; q < z; ++q)
            sum = __dadd_rn(sum, prod[(size_t)q * K + j]);
        C[(long long)r * ldc + j] = sum;
    }
}


template <int K>
int launch_stream_t(const spmm_csr_s *A, const double *d_B, long long ldb, double *d_C, long long ldc, int cap,
                    const int *cuts, int n_tiles, cudaStream_t stream)
{
    auto kern = spmm_stream_kernel<K>;
    const size_t smem = (size_t)cap * K * sizeof(double);
    static std::mutex mu;
    static std::map<int, size_t> configured; // per device
    {
        std::lock_guard<std::mutex> lk(mu);
        size_t &have = configured[A->device];
        if (smem > have && smem > 48 * 1024)
        {
            SPMM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            have = smem;
        }
    }
    kern<<<n_tiles, ST_TPB, smem, stream>>>(A->d_rowptr, A->d_colidx, A->d_vals, d_B, ldb, d_C, ldc, cuts, cap);
    SPMM_CUDA(cudaGetLastError());
    return SPMM_OK;
}

// non-zeros a tile may hold: about 32 KB of products per CTA (7 CTAs per SM), more for long rows
int stream_cap(const spmm_csr_s *A, int kc)
{
    const int t = tuning().stream_tile;
    int cap = t > 0 ? t : std::min(2048, 4096 / kc);
    cap = std::max(cap, 4 * A->sched.max_len);
    return cap;
}
} // namespace

bool stream_shape_ok(const spmm_csr_s *A, const double *d_B, long long ldb, const double *d_C, long long ldc, int kc)
{
    if (!(kc == 1 || kc == 2 || kc == 4 || kc == 8) || A->nnz == 0 || A->nnz >= (1ll << 31) - (1 << 20))
        return false;
    if (kc >= 2 && (ldb % 2 != 0 || (uintptr_t)d_B % 16 != 0))
        return false;
    (void)d_C;
    (void)ldc;
    return (size_t)stream_cap(A, kc) * kc * sizeof(double) <= 96 * 1024; // rows of up to 3072 / k non-zeros
}

int launch_stream(const spmm_csr_s *A, const double *d_B, long long ldb, double *d_C, long long ldc, int kc,
                  cudaStream_t stream)
{
    note_kernel("spmm_stream_kernel");
    const int cap = stream_cap(A, kc);
    const int tile_nnz = cap - A->sched.max_len; // a tile starts at the first row at or after a multiple of tile_nnz
    const int n_tiles = (int)((A->nnz + tile_nnz - 1) / tile_nnz);
    const int *cuts = nullptr;
    const int rc = cached_bounds(A, 3, n_tiles, stream, &cuts, [&](int *d) {
        stream_cuts_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, stream>>>(A->d_rowptr, A->n_rows, n_tiles, tile_nnz, d);
    });
    if (rc)
        return rc;
    switch (kc)
    {
    case 1: return launch_stream_t<1>(A, d_B, ldb, d_C, ldc, cap, cuts, n_tiles, stream);
    case 2: return launch_stream_t<2>(A, d_B, ldb, d_C, ldc, cap, cuts, n_tiles, stream);
    case 4: return launch_stream_t<4>(A, d_B, ldb, d_C, ldc, cap, cuts, n_tiles, stream);
    case 8: return launch_stream_t<8>(A, d_B, ldb, d_C, ldc, cap, cuts, n_tiles, stream);
    }
    set_error("stream kernel: k must be 1, 2, 4 or 8");
    return SPMM_ERR_INVALID;
}

} // namespace spmm
