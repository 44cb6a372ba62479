// spmm_rowblock.cu — row-block union format and its kernel (large k).
//
// Why: at k >= 16 the row kernel is bound by the L1 -> register gather of B rows
// (nnz * k * 8 bytes through a 128 B/clk/SM port), not by HBM. Neighbouring rows of
// FEM-like matrices share most of their columns, so a team that owns R consecutive
// rows and walks the UNION of their column lists loads every shared B row once and
// feeds R accumulators from registers. The union lists are built once per handle,
// on the device, next to the CSR arrays (which stay untouched and bit-exact):
//   blkptr[nb+1]           union entries of row block b (rows b*R .. b*R+R-1)
//   ucol[e]                column id of union entry e (ascending; a column that is
//                          duplicated inside one row gets one entry per duplicate)
//   uval[e*R + r]          value of row b*R+r at that column, 0.0 when absent
// A zero entry contributes fma(0, b, acc) = acc exactly when b is finite; the
// reference path never sees that product, so the kernel is only selected by AUTO
// when the fill ratio is modest, and B is assumed finite (DESIGN.md states it).
// The per-(row, column) accumulation order is still ascending column, as in
// SparseMatrixFatVectorMultiply.cpp:17-28.
#include <cub/device/device_scan.cuh>

#include "spmm_launch.cuh"

namespace spmm
{

namespace
{

// One thread per row block: R-way merge of the rows' ascending column lists.
// FILL=false counts entries (and flags unsorted rows); FILL=true writes them.
template <int R, bool FILL>
__global__ void rowblock_build_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                      const double *__restrict__ vals, int n_rows, int n_blocks, int *count,
                                      const int *__restrict__ blkptr, int *ucol, double *uval, int *unsorted)
{
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= n_blocks)
        return;
    int j[R], je[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
        const long long row = b * R + r;
        j[r] = row < n_rows ? rowptr[row] : 0;
        je[r] = row < n_rows ? rowptr[row + 1] : 0;
    }
    if constexpr (!FILL)
    {
        bool bad = false;
#pragma unroll
        for (int r = 0; r < R; ++r)
            for (int t = j[r] + 1; t < je[r]; ++t)
                bad |= colidx[t] < colidx[t - 1];
        if (bad)
            atomicExch(unsorted, 1);
    }
    int n = 0;
    long long o = FILL ? blkptr[b] : 0;
    while (true)
    {
        int c = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (j[r] < je[r])
                c = min(c, colidx[j[r]]);
        if (c == 0x7fffffff)
            break;
        if constexpr (FILL)
            ucol[o] = c;
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
            double v = 0.0;
            if (j[r] < je[r] && colidx[j[r]] == c)
            {
                if constexpr (FILL)
                    v = vals[j[r]];
                ++j[r];
            }
            if constexpr (FILL)
                uval[o * R + r] = v;
        }
        ++o;
        ++n;
    }
    if constexpr (!FILL)
        count[b] = n;
}

struct RbArgs
{
    const int *blkptr;
    const int *ucol;
    const double *uval;
    const double *B;
    double *C;
    long long ldb, ldc;
    int n_rows, n_blocks;
    const int *bounds; // CTA cuts over row blocks (gridDim.x+1 entries)
};

__device__ __forceinline__ void ld_vals(const double *p, double (&x)[2])
{
    double2 t;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "l"(p));
    x[0] = t.x;
    x[1] = t.y;
}
__device__ __forceinline__ void ld_vals(const double *p, double (&x)[4])
{
    double2 t, s;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "l"(p));
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(s.x), "=d"(s.y) : "l"(p + 2));
    x[0] = t.x;
    x[1] = t.y;
    x[2] = s.x;
    x[3] = s.y;
}

constexpr int RB_BLOCK_COST = 8;

__device__ __forceinline__ int rb_lower_bound(const int *__restrict__ blkptr, int n_blocks, long long target)
{
    int lo = 0, hi = n_blocks;
    while (lo < hi)
    {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)blkptr[mid] + (long long)RB_BLOCK_COST * mid < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__global__ void rb_bounds_kernel(const int *blkptr, int n_blocks, int grid, int *bounds)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > grid)
        return;
    const long long total = (long long)blkptr[n_blocks] + (long long)RB_BLOCK_COST * n_blocks;
    bounds[b] = b == 0 ? 0 : (b == grid ? n_blocks : rb_lower_bound(blkptr, n_blocks, (total * b + grid - 1) / grid));
}

// A team of KL lanes owns one row block (R rows) at a time; 32/KL teams per warp; every
// CTA sweeps one contiguous, equal-cost run of row blocks (same reasoning as the row kernel).
template <int R, int KL, int NV, int W, int U, int THREADS_>
__global__ void __launch_bounds__(THREADS_, min_blocks(NV, W, U, R, THREADS_)) spmm_rowblock_kernel(const RbArgs a)
{
    constexpr int RW = 32 / KL;
    constexpr int SLOTS = (THREADS_ / 32) * RW;
    using S = Slice<KL, NV, W>;

    __shared__ int s_chunk[2];
    if (threadIdx.x < 2)
        s_chunk[threadIdx.x] = a.bounds[blockIdx.x + threadIdx.x];
    __syncthreads();
    const int lo = s_chunk[0], hi = s_chunk[1];

    const int lane = threadIdx.x & 31;
    const int kl = lane % KL;
    const int slot = (threadIdx.x >> 5) * RW + lane / KL;
    const int tile0 = blockIdx.y * S::TILE;
    const double *__restrict__ Bk = a.B + tile0 + kl * W;

    int blk = lo + slot;
    int es = 0, ee = 0;
    if (blk < hi)
    {
        es = a.blkptr[blk];
        ee = a.blkptr[blk + 1];
    }
    for (int base = lo; base < hi; base += SLOTS)
    {
        const int nblk = blk + SLOTS; // the team's next row block: fetch its extent now
        int nes = 0, nee = 0;
        if (nblk < hi)
        {
            nes = a.blkptr[nblk];
            nee = a.blkptr[nblk + 1];
        }
        S acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
            acc[r].zero();
        if (es < ee)
        {
            int e = es;
            int c[U];
            double x[U][R];
#pragma unroll
            for (int u = 0; u < U; ++u)
            {
                const int ej = min(e + u, ee - 1);
                c[u] = ld_stream_i32(a.ucol + ej);
                ld_vals(a.uval + (long long)ej * R, x[u]);
            }
            while (true)
            {
                S b[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    b[u].template load<true>(Bk + (long long)c[u] * a.ldb, 0xffffffffu);
                const int en = e + U;
                const bool more = en < ee;
                int cn[U];
                double xn[U][R];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const int ej = min(en + u, ee - 1); // clamped, unconditional: overlaps the B loads
                    cn[u] = ld_stream_i32(a.ucol + ej);
                    ld_vals(a.uval + (long long)ej * R, xn[u]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (e + u < ee)
                    {
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            acc[r].fma(x[u][r], b[u]);
                    }
                if (!more)
                    break;
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    c[u] = cn[u];
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        x[u][r] = xn[u][r];
                }
                e = en;
            }
        }
        if (blk < hi)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
            {
                const long long row = (long long)blk * R + r;
                if (row < a.n_rows)
                    acc[r].store(a.C + row * a.ldc + tile0 + kl * W, 0xffffffffu);
            }
        }
        blk = nblk;
        es = nes;
        ee = nee;
    }
}

template <int R, int KL, int NV, int W, int U>
int launch_rb_one(const spmm_csr_s *A, const RbArgs &args, int tiles, int device, cudaStream_t stream)
{
    auto kern = spmm_rowblock_kernel<R, KL, NV, W, U, THREADS>;
    int per_sm = 1;
    int rc = kernel_info(kern, &per_sm);
    if (rc)
        return rc;
    const Tuning &t = tuning();
    if (t.rows_ctas_per_sm > 0)
        per_sm = std::min(per_sm, t.rows_ctas_per_sm);
    constexpr int SLOTS = (THREADS / 32) * (32 / KL);
    long long grid = (long long)device_props(device).sm_count * per_sm;
    grid = std::max(1LL, std::min(grid, ((long long)args.n_blocks + SLOTS - 1) / SLOTS));
    RbArgs a2 = args;
    rc = cached_bounds(A, 1, (int)grid, stream, &a2.bounds, [&](int *out) {
        rb_bounds_kernel<<<((unsigned)grid + 256) / 256, 256, 0, stream>>>(args.blkptr, args.n_blocks, (int)grid, out);
    });
    if (rc)
        return rc;
    kern<<<dim3((unsigned)grid, (unsigned)tiles), THREADS, 0, stream>>>(a2);
    SPMM_CUDA(cudaGetLastError());
    return SPMM_OK;
}

template <int R, int W>
int launch_rb_shape(const spmm_csr_s *A, int kl, int nv, int u, const RbArgs &a, int tiles, int dev, cudaStream_t s)
{
#define SPMM_RB_CASE(K, N)                                    \
    if (kl == K && nv == N)                                   \
    {                                                         \
        if (u >= 2)                                           \
            return launch_rb_one<R, K, N, W, 2>(A, a, tiles, dev, s); \
        return launch_rb_one<R, K, N, W, 1>(A, a, tiles, dev, s);    \
    }
    SPMM_RB_CASE(4, 1)
    SPMM_RB_CASE(8, 1)
    SPMM_RB_CASE(8, 2)
    SPMM_RB_CASE(8, 4)
    SPMM_RB_CASE(16, 4)
    SPMM_RB_CASE(32, 4)
#undef SPMM_RB_CASE
    return -1; // shape not covered: caller falls back to the row kernel
}

} // namespace

bool rowblock_shape_ok(int w, int kl, int nv, int tiles, int kc)
{
    if (kc != tiles * kl * nv * w)
        return false;
    return (kl == 4 && nv == 1) || (kl == 8 && (nv == 1 || nv == 2 || nv == 4)) || (kl == 16 && nv == 4) ||
           (kl == 32 && nv == 4);
}

int launch_rowblock(const spmm_csr_s *A, int w, int kl, int nv, int tiles, const double *d_B, long long ldb,
                    double *d_C, long long ldc, cudaStream_t stream)
{
    const Tuning &t = tuning();
    const int u = t.rows_unroll > 0 ? t.rows_unroll : (nv * A->rb_R >= 16 ? 1 : 2);
    RbArgs args;
    args.blkptr = A->d_blkptr;
    args.ucol = A->d_ucol;
    args.uval = A->d_uval;
    args.B = d_B;
    args.C = d_C;
    args.ldb = ldb;
    args.ldc = ldc;
    args.n_rows = A->n_rows;
    args.n_blocks = A->rb_blocks;
    args.bounds = nullptr;
    int rc = -1;
    if (A->rb_R == 2)
        rc = w == 2 ? launch_rb_shape<2, 2>(A, kl, nv, u, args, tiles, A->device, stream)
                    : launch_rb_shape<2, 1>(A, kl, nv, u, args, tiles, A->device, stream);
    else if (A->rb_R == 4)
        rc = w == 2 ? launch_rb_shape<4, 2>(A, kl, nv, u, args, tiles, A->device, stream)
                    : launch_rb_shape<4, 1>(A, kl, nv, u, args, tiles, A->device, stream);
    if (rc == -1)
    {
        set_error("row-block kernel: unsupported shape");
        return SPMM_ERR_UNSUPPORTED;
    }
    return rc;
}

void free_rowblocks(spmm_csr_s *A)
{
    drop_bounds(A, 1);
    cudaFree(A->d_blkptr);
    cudaFree(A->d_ucol);
    cudaFree(A->d_uval);
    A->d_blkptr = nullptr;
    A->d_ucol = nullptr;
    A->d_uval = nullptr;
    A->rb_R = 0;
    A->rb_blocks = 0;
    A->rb_entries = 0;
}

template <int R>
static int build_rowblocks_r(spmm_csr_s *A)
{
    const int nb = (A->n_rows + R - 1) / R;
    int *d_count = nullptr, *d_flag = nullptr;
    SPMM_CUDA(cudaMalloc(&d_count, sizeof(int) * ((size_t)nb + 1)));
    cudaError_t e = cudaMalloc(&d_flag, sizeof(int));
    if (e == cudaSuccess)
        e = cudaMemset(d_count, 0, sizeof(int) * ((size_t)nb + 1));
    if (e == cudaSuccess)
        e = cudaMemset(d_flag, 0, sizeof(int));
    if (e == cudaSuccess)
        e = cudaMalloc(&A->d_blkptr, sizeof(int) * ((size_t)nb + 1));
    const unsigned grid = (unsigned)std::max(1, (nb + 127) / 128);
    if (e == cudaSuccess && nb)
    {
        rowblock_build_kernel<R, false><<<grid, 128>>>(A->d_rowptr, A->d_colidx, A->d_vals, A->n_rows, nb, d_count,
                                                     nullptr, nullptr, nullptr, d_flag);
        e = cudaGetLastError();
    }
    size_t tmp_bytes = 0;
    void *d_tmp = nullptr;
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_count, A->d_blkptr, nb + 1);
    if (e == cudaSuccess)
        e = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_count, A->d_blkptr, nb + 1);
    int total = 0, unsorted = 0;
    if (e == cudaSuccess)
        e = cudaMemcpy(&total, A->d_blkptr + nb, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        e = cudaMemcpy(&unsorted, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_tmp);
    cudaFree(d_count);
    cudaFree(d_flag);
    if (e != cudaSuccess)
    {
        free_rowblocks(A);
        return cuda_fail(e, "row-block count", __FILE__, __LINE__);
    }
    if (unsorted)
    {
        free_rowblocks(A);
        set_error("row-block format needs ascending column ids inside every row");
        return SPMM_ERR_UNSUPPORTED;
    }
    e = cudaMalloc(&A->d_ucol, sizeof(int) * (size_t)std::max(total, 1));
    if (e == cudaSuccess)
        e = cudaMalloc(&A->d_uval, sizeof(double) * (size_t)std::max(total, 1) * R);
    if (e == cudaSuccess && nb)
    {
        rowblock_build_kernel<R, true><<<grid, 128>>>(A->d_rowptr, A->d_colidx, A->d_vals, A->n_rows, nb, nullptr,
                                                    A->d_blkptr, A->d_ucol, A->d_uval, nullptr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
        free_rowblocks(A);
        return cuda_fail(e, "row-block fill", __FILE__, __LINE__);
    }
    A->rb_R = R;
    A->rb_blocks = nb;
    A->rb_entries = total;
    return SPMM_OK;
}

} // namespace spmm

using namespace spmm;

extern "C"
{

int spmm_csr_build_rowblocks(spmm_csr_t A, int rows_per_block)
{
    SPMM_REQUIRE(A != nullptr, "handle is NULL");
    SPMM_REQUIRE(rows_per_block == -1 || rows_per_block == 0 || rows_per_block == 2 || rows_per_block == 4,
                 "rows_per_block must be -1 (auto), 0 (drop), 2 or 4");
    SPMM_CUDA(cudaSetDevice(A->device));
    free_rowblocks(A);
    if (rows_per_block == 0 || A->n_rows == 0 || A->nnz == 0)
        return SPMM_OK;
    if (rows_per_block == 2)
        return build_rowblocks_r<2>(A);
    if (rows_per_block == 4)
        return build_rowblocks_r<4>(A);
    // auto: only for short, regular rows (the row-length schedule decides); keep the widest
    // block whose zero fill stays modest, else none. Unsorted rows simply keep the CSR kernels.
    if (A->sched.auto_kernel != SPMM_KERNEL_ROWS || A->sched.mean_len < 4.0 || A->sched.max_len > 1024)
        return SPMM_OK;
    int rc = build_rowblocks_r<4>(A);
    if (rc == SPMM_OK && (double)A->rb_entries * 4 <= 1.75 * (double)A->nnz)
        return SPMM_OK;
    if (rc != SPMM_OK && rc != SPMM_ERR_UNSUPPORTED)
        return rc;
    free_rowblocks(A);
    if (rc == SPMM_ERR_UNSUPPORTED)
        return SPMM_OK;
    rc = build_rowblocks_r<2>(A);
    if (rc == SPMM_OK && (double)A->rb_entries * 2 <= 1.4 * (double)A->nnz)
        return SPMM_OK;
    free_rowblocks(A);
    return (rc == SPMM_ERR_UNSUPPORTED) ? SPMM_OK : rc;
}

int spmm_csr_rowblock_info(spmm_csr_t A, int *rows_per_block, long long *union_entries, double *fill_ratio)
{
    SPMM_REQUIRE(A != nullptr, "handle is NULL");
    if (rows_per_block)
        *rows_per_block = A->rb_R;
    if (union_entries)
        *union_entries = A->rb_entries;
    if (fill_ratio)
        *fill_ratio = (A->rb_R && A->nnz) ? (double)A->rb_entries * A->rb_R / (double)A->nnz : 0.0;
    return SPMM_OK;
}

} // extern "C"
