// spmm_staged.cu — row kernel with the A stream staged through shared memory.
//
// Why (profiles/r1_ncu_rows_k64.md + the diagnostic builds): in the CSR row kernel the chain
// "column id load (L2/DRAM latency) -> dependent B-row load -> FMA" is paid by every team for
// every group of a ~20-element row, and no amount of unrolling, occupancy or cache residency
// moved the 105-118 us plateau; with the B rows forced into L1 the kernel still took 85-100 us,
// with the ids taken out of the dependency 56 us. So the ids/values must not be fetched by the
// consumer through a dependent long-latency load at all.
//
// Here a CTA (one per SM) streams the contiguous slice of colidx/vals that belongs to its rows
// through a 4-stage ring in shared memory with asynchronous 16-byte copies (cp.async.cg, L2 ->
// smem, no L1 allocation, no registers held), two tiles ahead of the consumers. Teams read ids
// and values with LDS (~30 cycles), so the only long-latency access left on the critical path is
// the B-row gather itself. The row pointer of the CTA's rows is staged once as well.
//
//   tile t      = CAP consecutive non-zeros starting at base + t*CAP (base = first nnz rounded
//                 down to 4, so every copy is 16-byte aligned)
//   phase t     = the rows whose first non-zero lies in tile t; they may run into tile t+1, so
//                 tiles t and t+1 are resident while tiles t+2, t+3 are in flight
//   ring index  = (j - base) mod 4*CAP
// A row longer than CAP would not fit in two tiles: the launcher only selects this kernel when
// the handle's schedule says max_row_len <= CAP.
#include "spmm_launch.cuh"

namespace spmm
{

namespace
{
constexpr int ST_CAP = 2048;   // non-zeros per tile
constexpr int ST_STAGES = 4;   // ring = 4 tiles (power of two -> mask)
constexpr int ST_RING = ST_CAP * ST_STAGES;
constexpr int ST_RT = 1024;    // rows per super-tile (row pointer staged per super-tile)

struct StagedArgs
{
    const int *rowptr;
    const int *colidx;
    const double *vals;
    const double *B;
    double *C;
    long long ldb, ldc;
    int n_rows, rows_per_cta, tiles;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int KL, int NV, int U, int THREADS_>
__global__ void __launch_bounds__(THREADS_, 1) spmm_staged_kernel(const StagedArgs a)
{
    constexpr int W = 2;
    constexpr int RW = 32 / KL;
    constexpr int SLOTS = (THREADS_ / 32) * RW;
    using S = Slice<KL, NV, W>;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_val = reinterpret_cast<double *>(smem_raw);                       // ST_RING doubles
    int *s_col = reinterpret_cast<int *>(smem_raw + sizeof(double) * ST_RING);  // ST_RING ints
    int *s_rp = s_col + ST_RING;                                                // ST_RT + 1 ints

    const int lane = threadIdx.x & 31;
    const int kl = lane % KL;
    const int slot = (threadIdx.x >> 5) * RW + lane / KL;
    const int cta_lo = min(a.n_rows, (int)blockIdx.x * a.rows_per_cta);
    const int cta_hi = min(a.n_rows, cta_lo + a.rows_per_cta);

    for (int tile = 0; tile < a.tiles; ++tile)
    {
        const int tile0 = tile * S::TILE;
        const double *__restrict__ Bk = a.B + tile0 + kl * W;
        for (int R0 = cta_lo; R0 < cta_hi; R0 += ST_RT)
        {
            const int R1 = min(cta_hi, R0 + ST_RT);
            const int nr = R1 - R0;
            __syncthreads(); // previous super-tile / column tile completely consumed
            for (int i = threadIdx.x; i <= nr; i += THREADS_)
                s_rp[i] = a.rowptr[R0 + i];
            __syncthreads();
            const int n0 = s_rp[0], n1 = s_rp[nr];
            const int base = n0 & ~3;
            const int n_t = (n1 - base + ST_CAP - 1) / ST_CAP;

            // copies of tile t: CAP/4 16-byte pieces of ids, CAP/2 of values; pieces past n1 are skipped
            auto issue = [&](int t) {
                if (t < n_t)
                {
                    const int g0 = base + t * ST_CAP;               // first global element of the tile
                    const int r0 = (t & (ST_STAGES - 1)) * ST_CAP;  // its ring offset
                    const int live = min(ST_CAP, n1 - g0);          // elements needed
                    for (int p = threadIdx.x; p * 4 < live; p += THREADS_)
                        cp_async16(s_col + r0 + p * 4, a.colidx + g0 + p * 4);
                    for (int p = threadIdx.x; p * 2 < live; p += THREADS_)
                        cp_async16(s_val + r0 + p * 2, a.vals + g0 + p * 2);
                }
                cp_async_commit(); // always: keeps the group count uniform across threads
            };
            issue(0);
            issue(1);
            issue(2);
            for (int t = 0; t < n_t; ++t)
            {
                __syncthreads(); // everyone finished phase t-1 -> the stage of tile t-1 may be overwritten
                issue(t + 3);
                cp_async_wait<2>(); // this thread's copies of tiles <= t+1 have landed
                __syncthreads();    // ... and everybody else's
                // rows of phase t: first non-zero in [base + t*CAP, base + (t+1)*CAP)
                auto first_row_at = [&](int nz) { // first local row r with s_rp[r] >= nz
                    int l = 0, h = nr;
                    while (l < h)
                    {
                        const int m = (l + h) >> 1;
                        if (s_rp[m] < nz)
                            l = m + 1;
                        else
                            h = m;
                    }
                    return l;
                };
                const int r_lo = t == 0 ? 0 : first_row_at(base + t * ST_CAP);
                const int r_hi = t == n_t - 1 ? nr : first_row_at(base + (t + 1) * ST_CAP);
                for (int lr = r_lo + slot; lr < r_hi; lr += SLOTS)
                {
                    const int js = s_rp[lr], je = s_rp[lr + 1];
                    S acc;
                    acc.zero();
                    if (js < je)
                    {
                        int c[U];
                        double x[U];
#pragma unroll
                        for (int u = 0; u < U; ++u)
                        {
                            const int r = (min(js + u, je - 1) - base) & (ST_RING - 1);
                            c[u] = s_col[r];
                            x[u] = s_val[r];
                        }
                        for (int j = js; j < je; j += U)
                        {
                            S b[U];
#if defined(SPMM_ABLATE) && (SPMM_ABLATE & 2) // diagnostic: no B gather
#pragma unroll
                            for (int u = 0; u < U; ++u)
#pragma unroll
                                for (int i = 0; i < NV * W; ++i)
                                    b[u].v[i] = (double)c[u];
#elif defined(SPMM_ABLATE) && (SPMM_ABLATE & 8) // diagnostic: every B row folded into 64 rows (all L1 hits)
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                b[u].template load<true>(Bk + (long long)(c[u] & 63) * a.ldb, 0xffffffffu);
#else
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                b[u].template load<true>(Bk + (long long)c[u] * a.ldb, 0xffffffffu);
#endif
                            int cn[U];
                            double xn[U];
#pragma unroll
                            for (int u = 0; u < U; ++u)
                            {
                                const int r = (min(j + U + u, je - 1) - base) & (ST_RING - 1);
                                cn[u] = s_col[r];
                                xn[u] = s_val[r];
                            }
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                if (j + u < je)
                                    acc.fma(x[u], b[u]);
#pragma unroll
                            for (int u = 0; u < U; ++u)
                            {
                                c[u] = cn[u];
                                x[u] = xn[u];
                            }
                        }
                    }
                    acc.store(a.C + (long long)(R0 + lr) * a.ldc + tile0 + kl * W, 0xffffffffu);
                }
            }
            if (n_t == 0) // only empty rows in this super-tile: C rows are still zeros
            {
                S z;
                z.zero();
                for (int lr = slot; lr < nr; lr += SLOTS)
                    z.store(a.C + (long long)(R0 + lr) * a.ldc + tile0 + kl * W, 0xffffffffu);
            }
            cp_async_wait<0>();
        }
    }
}

template <int KL, int NV, int U, int TH>
int launch_staged_one(const spmm_csr_s *A, StagedArgs args, int tiles, cudaStream_t stream)
{
    auto kern = spmm_staged_kernel<KL, NV, U, TH>;
    constexpr size_t smem = sizeof(double) * ST_RING + sizeof(int) * ST_RING + sizeof(int) * (ST_RT + 4);
    static bool configured = false;
    if (!configured)
    {
        SPMM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const Tuning &t = tuning();
    const int per_sm = t.rows_ctas_per_sm > 0 ? t.rows_ctas_per_sm : 1;
    long long grid = (long long)device_props(A->device).sm_count * per_sm;
    // contiguous, equal-row chunks, multiples of 32 rows
    long long rows_per = ((long long)A->n_rows + grid - 1) / grid;
    rows_per = std::max(32LL, (rows_per + 31) / 32 * 32);
    grid = ((long long)A->n_rows + rows_per - 1) / rows_per;
    args.rows_per_cta = (int)rows_per;
    args.tiles = tiles;
    kern<<<(unsigned)grid, TH, smem, stream>>>(args);
    SPMM_CUDA(cudaGetLastError());
    return SPMM_OK;
}

} // namespace

bool staged_shape_ok(const spmm_csr_s *A, int w, int kl, int nv, int tiles, int kc)
{
    return A->owns && A->sched.max_len <= ST_CAP && w == 2 && kl == 8 && (nv == 1 || nv == 2 || nv == 4) &&
           kc == tiles * kl * nv * w && ((uintptr_t)A->d_colidx % 16 == 0) && ((uintptr_t)A->d_vals % 16 == 0);
}

int launch_staged(const spmm_csr_s *A, int nv, int tiles, const double *d_B, long long ldb, double *d_C, long long ldc,
                  cudaStream_t stream)
{
    const Tuning &t = tuning();
    StagedArgs args;
    args.rowptr = A->d_rowptr;
    args.colidx = A->d_colidx;
    args.vals = A->d_vals;
    args.B = d_B;
    args.C = d_C;
    args.ldb = ldb;
    args.ldc = ldc;
    args.n_rows = A->n_rows;
    args.rows_per_cta = 0;
    args.tiles = tiles;
    const int u = t.rows_unroll > 0 ? t.rows_unroll : (nv >= 4 ? 2 : 4);
    const int th = t.rows_threads > 0 ? t.rows_threads : 512;
#define SPMM_ST_CASE(N, UU)                                                      \
    if (nv == N && u == UU)                                                      \
        return th >= 1024  ? launch_staged_one<8, N, UU, 1024>(A, args, tiles, stream) \
               : th >= 512 ? launch_staged_one<8, N, UU, 512>(A, args, tiles, stream)  \
                           : launch_staged_one<8, N, UU, 256>(A, args, tiles, stream);
    SPMM_ST_CASE(1, 2) SPMM_ST_CASE(1, 4) SPMM_ST_CASE(2, 2) SPMM_ST_CASE(2, 4) SPMM_ST_CASE(4, 2) SPMM_ST_CASE(4, 4)
    SPMM_ST_CASE(1, 1) SPMM_ST_CASE(2, 1) SPMM_ST_CASE(4, 1)
#undef SPMM_ST_CASE
    set_error("staged kernel: unsupported shape");
    return SPMM_ERR_UNSUPPORTED;
}

} // namespace spmm
