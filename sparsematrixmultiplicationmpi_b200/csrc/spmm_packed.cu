// spmm_packed.cu — warp-packed stream layout ("SELL-4-4" of row blocks) and its kernel.
//
// Why (ncu, profiles/): in the CSR row kernel every team fetches the column ids / values of its
// own row with 4- and 8-byte loads from four different places per warp instruction (2 of the ~6
// L1 wavefronts per non-zero at k=64), and the first group of every row waits for them. Here
// the A stream is re-laid once per handle so that a warp reads it as ONE contiguous,
// 16-byte-vectorised stream whose addresses do not depend on any loaded value:
//
//   slice  = TPW consecutive row blocks (TPW = 32/KL teams of a warp; a row block = R rows that
//            share the union of their column lists, R = 1 is plain rows),
//   group  = G = 4 consecutive steps of the slice; per group and team: 4 column ids (one int4)
//            and 4*R values (2*R double2), teams side by side -> a warp reads 64 B of ids and
//            128*R B of values per group, fully coalesced,
//   sptr[] = first group of every slice; rows shorter than the slice are padded with id -1.
//
// The kernel keeps the ids/values of the NEXT group in flight while the B-row slices of the
// current group are loaded and accumulated; nothing on its critical path waits for the A stream.
// The CSR arrays stay untouched (bit-exact CSR is the contract); this is a second, derived layout
// like the row-block one. Per-(row, column) accumulation order is still ascending column.
#include <cub/device/device_scan.cuh>

#include "spmm_launch.cuh"

namespace spmm
{

namespace
{
constexpr int PG = 4; // steps per group

// One thread per slice: groups needed = ceil(max entries over the slice's TPW units / PG).
// R == 1: a unit is a CSR row; R > 1: a unit is a row block of the handle's union layout.
template <int R>
__global__ void packed_count_kernel(const int *__restrict__ extent, int n_units, int tpw, int n_slices, int *groups)
{
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n_slices)
        return;
    int mx = 0;
    for (int t = 0; t < tpw; ++t)
    {
        const long long u = s * tpw + t;
        if (u < n_units)
            mx = max(mx, extent[u + 1] - extent[u]);
    }
    groups[s] = (mx + PG - 1) / PG;
}

// One thread per unit: scatter its entries into the slice layout, pad with id -1 / value 0.
template <int R>
__global__ void packed_fill_kernel(const int *__restrict__ extent, const int *__restrict__ cols,
                                   const double *__restrict__ vals, int n_units, int tpw, int n_slices,
                                   const int *__restrict__ sptr, int *pcol, double *pval)
{
    const long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= (long long)n_slices * tpw)
        return;
    const long long s = u / tpw;
    const int t = (int)(u % tpw);
    const int g0 = sptr[s], ng = sptr[s + 1] - g0;
    int e0 = 0, len = 0;
    if (u < n_units)
    {
        e0 = extent[u];
        len = extent[u + 1] - e0;
    }
    for (int j = 0; j < ng * PG; ++j)
    {
        const long long slot = ((long long)(g0 + j / PG) * tpw + t) * PG + (j % PG);
        const bool live = j < len;
        pcol[slot] = live ? cols[e0 + j] : -1;
#pragma unroll
        for (int r = 0; r < R; ++r)
            pval[slot * R + r] = live ? vals[(long long)(e0 + j) * R + r] : 0.0;
    }
}

struct PackedArgs
{
    const int *sptr;
    const int *pcol;
    const double *pval;
    const double *B;
    double *C;
    long long ldb, ldc;
    int n_rows, n_slices, tiles;
    const int *bounds; // CTA cuts over slices
};

__device__ __forceinline__ int4 ld_stream_i32x4(const int *p)
{
    int4 v;
    asm("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p)
{
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

constexpr int PK_SLICE_COST = 2;

__global__ void packed_bounds_kernel(const int *sptr, int n_slices, int grid, int *bounds)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > grid)
        return;
    const long long total = (long long)sptr[n_slices] + (long long)PK_SLICE_COST * n_slices;
    int r;
    if (b == 0)
        r = 0;
    else if (b == grid)
        r = n_slices;
    else
    {
        const long long target = (total * b + grid - 1) / grid;
        int lo = 0, hi = n_slices;
        while (lo < hi)
        {
            const int mid = lo + ((hi - lo) >> 1);
            if ((long long)sptr[mid] + (long long)PK_SLICE_COST * mid < target)
                lo = mid + 1;
            else
                hi = mid;
        }
        r = lo;
    }
    bounds[b] = r;
}

// A warp owns one slice at a time (TPW = 32/KL teams, each on its own unit of R rows).
template <int R, int KL, int NV, int THREADS_, bool SWEEP>
__global__ void __launch_bounds__(THREADS_, SWEEP ? 1 : min_blocks(NV, 2, PG, R, THREADS_))
    spmm_packed_kernel(const PackedArgs a)
{
    constexpr int TPW = 32 / KL;
    constexpr int WARPS = THREADS_ / 32;
    constexpr int W = 2;
    using S = Slice<KL, NV, W>;

    __shared__ int s_chunk[2];
    if (threadIdx.x < 2)
        s_chunk[threadIdx.x] = a.bounds[blockIdx.x + threadIdx.x];
    __syncthreads();
    const int lo = s_chunk[0], hi = s_chunk[1];

    const int lane = threadIdx.x & 31;
    const int team = lane / KL, kl = lane % KL;
    const int warp = threadIdx.x >> 5;

    for (int tile = SWEEP ? 0 : (int)blockIdx.y; tile < (SWEEP ? a.tiles : (int)blockIdx.y + 1); ++tile)
    {
        const int tile0 = tile * S::TILE;
        const double *__restrict__ Bk = a.B + tile0 + kl * W;
        int s = lo + warp;
        int g0 = 0, g1 = 0;
        if (s < hi)
        {
            g0 = a.sptr[s];
            g1 = a.sptr[s + 1];
        }
        for (; s < hi; s += WARPS)
        {
            const int ns = s + WARPS; // the warp's next slice: fetch its extent now
            int ng0 = 0, ng1 = 0;
            if (ns < hi)
            {
                ng0 = a.sptr[ns];
                ng1 = a.sptr[ns + 1];
            }
            S acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                acc[r].zero();
            if (g0 < g1)
            {
                const int *cp = a.pcol + ((long long)g0 * TPW + team) * PG;
                const double *vp = a.pval + ((long long)g0 * TPW + team) * PG * R;
                int4 c = ld_stream_i32x4(cp);
                double x[PG * R];
#pragma unroll
                for (int i = 0; i < PG * R; i += 2)
                {
                    const double2 t = ld_stream_f64x2(vp + i);
                    x[i] = t.x;
                    x[i + 1] = t.y;
                }
                for (int g = g0; g < g1; ++g)
                {
                    const int cc[PG] = {c.x, c.y, c.z, c.w};
                    S b[PG];
#pragma unroll
                    for (int u = 0; u < PG; ++u)
                        if (cc[u] >= 0)
                            b[u].template load<true>(Bk + (long long)cc[u] * a.ldb, 0xffffffffu);
                    // next group's ids / values: sequential addresses, independent of everything above
                    const int gn = min(g + 1, g1 - 1);
                    const int4 cn = ld_stream_i32x4(a.pcol + ((long long)gn * TPW + team) * PG);
                    double xn[PG * R];
                    const double *vn = a.pval + ((long long)gn * TPW + team) * PG * R;
#pragma unroll
                    for (int i = 0; i < PG * R; i += 2)
                    {
                        const double2 t = ld_stream_f64x2(vn + i);
                        xn[i] = t.x;
                        xn[i + 1] = t.y;
                    }
#pragma unroll
                    for (int u = 0; u < PG; ++u)
                        if (cc[u] >= 0)
                        {
#pragma unroll
                            for (int r = 0; r < R; ++r)
                                acc[r].fma(x[u * R + r], b[u]);
                        }
                    c = cn;
#pragma unroll
                    for (int i = 0; i < PG * R; ++i)
                        x[i] = xn[i];
                }
            }
            const long long unit = (long long)s * TPW + team;
#pragma unroll
            for (int r = 0; r < R; ++r)
            {
                const long long row = unit * R + r;
                if (row < a.n_rows)
                    acc[r].store(a.C + row * a.ldc + tile0 + kl * W, 0xffffffffu);
            }
            g0 = ng0;
            g1 = ng1;
        }
    }
}

template <int R, int KL, int NV, int TH, bool SWEEP>
int launch_packed_one(const spmm_csr_s *A, const PackedArgs &args, int tiles, cudaStream_t stream)
{
    auto kern = spmm_packed_kernel<R, KL, NV, TH, SWEEP>;
    int per_sm = 1;
    int rc = kernel_info(kern, &per_sm, TH);
    if (rc)
        return rc;
    const Tuning &t = tuning();
    if (SWEEP)
        per_sm = 1;
    if (t.rows_ctas_per_sm > 0)
        per_sm = SWEEP ? t.rows_ctas_per_sm : std::min(per_sm, t.rows_ctas_per_sm);
    long long grid = (long long)device_props(A->device).sm_count * per_sm;
    grid = std::max(1LL, std::min(grid, ((long long)args.n_slices + TH / 32 - 1) / (TH / 32)));
    PackedArgs a2 = args;
    a2.tiles = tiles;
    rc = cached_bounds(A, 2, (int)grid, stream, &a2.bounds, [&](int *out) {
        packed_bounds_kernel<<<((unsigned)grid + 256) / 256, 256, 0, stream>>>(args.sptr, args.n_slices, (int)grid, out);
    });
    if (rc)
        return rc;
    kern<<<dim3((unsigned)grid, SWEEP ? 1u : (unsigned)tiles), TH, 0, stream>>>(a2);
    SPMM_CUDA(cudaGetLastError());
    return SPMM_OK;
}

template <int R, int KL>
int launch_packed_nv(const spmm_csr_s *A, int nv, int threads, bool sweep, const PackedArgs &a, int tiles, cudaStream_t s)
{
#define SPMM_PK_CASE(N)                                                            \
    if (nv == N)                                                                   \
    {                                                                              \
        if (sweep)                                                                 \
            return threads >= 512 ? launch_packed_one<R, KL, N, 512, true>(A, a, tiles, s) \
                                  : launch_packed_one<R, KL, N, 256, true>(A, a, tiles, s); \
        return launch_packed_one<R, KL, N, 256, false>(A, a, tiles, s);           \
    }
    SPMM_PK_CASE(1)
    SPMM_PK_CASE(2)
    SPMM_PK_CASE(4)
#undef SPMM_PK_CASE
    return -1;
}

template <int R>
int build_packed_r(spmm_csr_s *A, int kl)
{
    const int tpw = 32 / kl;
    const int *extent = R == 1 ? A->d_rowptr : A->d_blkptr;
    const int *cols = R == 1 ? A->d_colidx : A->d_ucol;
    const double *vals = R == 1 ? A->d_vals : A->d_uval;
    const int n_units = R == 1 ? A->n_rows : A->rb_blocks;
    const int n_slices = (n_units + tpw - 1) / tpw;
    int *d_groups = nullptr;
    SPMM_CUDA(cudaMalloc(&d_groups, sizeof(int) * ((size_t)n_slices + 1)));
    cudaError_t e = cudaMemset(d_groups, 0, sizeof(int) * ((size_t)n_slices + 1));
    if (e == cudaSuccess)
        e = cudaMalloc(&A->d_sptr, sizeof(int) * ((size_t)n_slices + 1));
    if (e == cudaSuccess && n_slices)
    {
        packed_count_kernel<R><<<(n_slices + 127) / 128, 128>>>(extent, n_units, tpw, n_slices, d_groups);
        e = cudaGetLastError();
    }
    size_t tmp_bytes = 0;
    void *d_tmp = nullptr;
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_groups, A->d_sptr, n_slices + 1);
    if (e == cudaSuccess)
        e = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_groups, A->d_sptr, n_slices + 1);
    int total = 0;
    if (e == cudaSuccess)
        e = cudaMemcpy(&total, A->d_sptr + n_slices, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_tmp);
    cudaFree(d_groups);
    const size_t slots = (size_t)std::max(total, 1) * tpw * PG;
    if (e == cudaSuccess)
        e = cudaMalloc(&A->d_pcol, sizeof(int) * slots);
    if (e == cudaSuccess)
        e = cudaMalloc(&A->d_pval, sizeof(double) * slots * R);
    if (e == cudaSuccess && n_slices)
    {
        const long long threads = (long long)n_slices * tpw;
        packed_fill_kernel<R><<<(unsigned)((threads + 127) / 128), 128>>>(extent, cols, vals, n_units, tpw, n_slices,
                                                                          A->d_sptr, A->d_pcol, A->d_pval);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
        free_packed(A);
        return cuda_fail(e, "packed stream build", __FILE__, __LINE__);
    }
    A->pk_R = R;
    A->pk_kl = kl;
    A->pk_slices = n_slices;
    A->pk_groups = total;
    return SPMM_OK;
}

} // namespace

void free_packed(spmm_csr_s *A)
{
    drop_bounds(A, 2);
    cudaFree(A->d_sptr);
    cudaFree(A->d_pcol);
    cudaFree(A->d_pval);
    A->d_sptr = nullptr;
    A->d_pcol = nullptr;
    A->d_pval = nullptr;
    A->pk_R = 0;
    A->pk_kl = 0;
    A->pk_slices = 0;
    A->pk_groups = 0;
}

bool packed_shape_ok(const spmm_csr_s *A, int w, int kl, int nv, int tiles, int kc)
{
    return A->pk_R != 0 && w == 2 && kl == A->pk_kl && (nv == 1 || nv == 2 || nv == 4) && kc == tiles * kl * nv * w;
}

int launch_packed(const spmm_csr_s *A, int nv, int tiles, const double *d_B, long long ldb, double *d_C, long long ldc,
                  cudaStream_t stream)
{
    const Tuning &t = tuning();
    PackedArgs args;
    args.sptr = A->d_sptr;
    args.pcol = A->d_pcol;
    args.pval = A->d_pval;
    args.B = d_B;
    args.C = d_C;
    args.ldb = ldb;
    args.ldc = ldc;
    args.n_rows = A->n_rows;
    args.n_slices = A->pk_slices;
    args.tiles = tiles;
    args.bounds = nullptr;
    const bool sweep = t.rows_sweep > 0;
    const int threads = t.rows_threads > 0 ? t.rows_threads : 512;
    int rc = -1;
    if (A->pk_R == 1 && A->pk_kl == 8)
        rc = launch_packed_nv<1, 8>(A, nv, threads, sweep, args, tiles, stream);
    else if (A->pk_R == 2 && A->pk_kl == 8)
        rc = launch_packed_nv<2, 8>(A, nv, threads, sweep, args, tiles, stream);
    else if (A->pk_R == 1 && A->pk_kl == 16)
        rc = launch_packed_nv<1, 16>(A, nv, threads, sweep, args, tiles, stream);
    else if (A->pk_R == 2 && A->pk_kl == 16)
        rc = launch_packed_nv<2, 16>(A, nv, threads, sweep, args, tiles, stream);
    if (rc == -1)
    {
        set_error("packed kernel: unsupported shape");
        return SPMM_ERR_UNSUPPORTED;
    }
    return rc;
}

} // namespace spmm

using namespace spmm;

extern "C"
{

int spmm_csr_build_packed(spmm_csr_t A, int rows_per_unit, int lanes_per_row)
{
    SPMM_REQUIRE(A != nullptr, "handle is NULL");
    SPMM_REQUIRE(rows_per_unit == 0 || rows_per_unit == 1 || rows_per_unit == 2, "rows_per_unit must be 0 (drop), 1 or 2");
    SPMM_REQUIRE(rows_per_unit == 0 || lanes_per_row == 8 || lanes_per_row == 16, "lanes_per_row must be 8 or 16");
    SPMM_CUDA(cudaSetDevice(A->device));
    free_packed(A);
    if (rows_per_unit == 0 || A->n_rows == 0 || A->nnz == 0)
        return SPMM_OK;
    if (rows_per_unit == 1)
        return build_packed_r<1>(A, lanes_per_row);
    SPMM_REQUIRE(A->rb_R == 2, "rows_per_unit = 2 needs spmm_csr_build_rowblocks(A, 2) first");
    return build_packed_r<2>(A, lanes_per_row);
}

int spmm_csr_packed_info(spmm_csr_t A, int *rows_per_unit, int *lanes_per_row, long long *slots, double *fill_ratio)
{
    SPMM_REQUIRE(A != nullptr, "handle is NULL");
    const long long n = A->pk_R ? (long long)A->pk_groups * (32 / A->pk_kl) * PG : 0;
    if (rows_per_unit)
        *rows_per_unit = A->pk_R;
    if (lanes_per_row)
        *lanes_per_row = A->pk_kl;
    if (slots)
        *slots = n;
    if (fill_ratio)
        *fill_ratio = (A->pk_R && A->nnz) ? (double)n * A->pk_R / (double)A->nnz : 0.0;
    return SPMM_OK;
}

} // extern "C"
