// spmm_rows.cu — launcher + instantiations of the team-per-row kernel (spmm_kernels.cuh).
//
// Team shape is picked from k and from the handle's row-length schedule:
//   KL  lanes across columns   = next pow2 >= ceil(k_tile / VEC), k_tile <= 32*VEC per blockIdx.y
//   VEC doubles per lane       = 2 (128-bit loads) when k, ldb, ldc are even and pointers 16 B aligned,
//                                4 for k >= 128 multiples of 4, else 1
//   NP  non-zeros side by side = pow2 near mean_row_len/2, capped at 32/KL
//   U   steps in flight        = 4/NP (at least 1)
// All of these can be overridden through spmm_tune_set() for measurement.
#include <algorithm>
#include <mutex>
#include <unordered_set>

#include "spmm_internal.h"
#include "spmm_kernels.cuh"

namespace spmm
{

namespace
{
constexpr int THREADS = 256;

std::mutex g_attr_mu;
std::unordered_set<const void *> g_attr_done;

template <int KL, int VEC, int NP, int U>
int launch_one(const RowsArgs &args, int tiles, int device, cudaStream_t stream)
{
    auto kern = spmm_rows_kernel<KL, VEC, NP, U, THREADS>;
    const void *fn = (const void *)kern;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lk(g_attr_mu);
        if (!g_attr_done.count(fn))
        {
            // no dynamic smem: give the whole unified array to L1D (B-row reuse lives there)
            SPMM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
            g_attr_done.insert(fn);
        }
    }
    SPMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0));
    if (per_sm < 1)
        per_sm = 1;
    const Tuning &t = tuning();
    if (t.rows_ctas_per_sm > 0)
        per_sm = std::min(per_sm, t.rows_ctas_per_sm);
    constexpr int SLOTS = (THREADS / 32) * (32 / (KL * NP));
    const long long rows = (long long)args.row_end - args.row_begin;
    long long grid = (long long)device_props(device).sm_count * per_sm;
    grid = std::max(1LL, std::min(grid, (rows + SLOTS - 1) / SLOTS));
    kern<<<dim3((unsigned)grid, (unsigned)tiles), THREADS, 0, stream>>>(args);
    SPMM_CUDA(cudaGetLastError());
    return SPMM_OK;
}

template <int KL, int VEC, int NP>
int pick_u(int u, const RowsArgs &a, int tiles, int dev, cudaStream_t s)
{
    if constexpr (NP == 1)
    {
        if (u >= 8)
            return launch_one<KL, VEC, NP, 8>(a, tiles, dev, s);
        if (u >= 4)
            return launch_one<KL, VEC, NP, 4>(a, tiles, dev, s);
        return launch_one<KL, VEC, NP, 2>(a, tiles, dev, s);
    }
    else if constexpr (NP <= 4)
    {
        if (u >= 4)
            return launch_one<KL, VEC, NP, 4>(a, tiles, dev, s);
        if (u >= 2)
            return launch_one<KL, VEC, NP, 2>(a, tiles, dev, s);
        return launch_one<KL, VEC, NP, 1>(a, tiles, dev, s);
    }
    else
    {
        if (u >= 2)
            return launch_one<KL, VEC, NP, 2>(a, tiles, dev, s);
        return launch_one<KL, VEC, NP, 1>(a, tiles, dev, s);
    }
}

template <int KL, int VEC>
int pick_np(int np, int u, const RowsArgs &a, int tiles, int dev, cudaStream_t s)
{
    constexpr int MAXNP = 32 / KL;
    np = std::max(1, std::min(np, MAXNP));
#define SPMM_NP_CASE(N)                               \
    if constexpr (N <= MAXNP)                         \
        if (np >= N)                                  \
            return pick_u<KL, VEC, N>(u, a, tiles, dev, s);
    SPMM_NP_CASE(32)
    SPMM_NP_CASE(16)
    SPMM_NP_CASE(8)
    SPMM_NP_CASE(4)
    SPMM_NP_CASE(2)
#undef SPMM_NP_CASE
    return pick_u<KL, VEC, 1>(u, a, tiles, dev, s);
}

template <int VEC>
int pick_kl(int kl, int np, int u, const RowsArgs &a, int tiles, int dev, cudaStream_t s)
{
    switch (kl)
    {
    case 32:
        return pick_np<32, VEC>(np, u, a, tiles, dev, s);
    case 16:
        return pick_np<16, VEC>(np, u, a, tiles, dev, s);
    case 8:
        if constexpr (VEC <= 2)
            return pick_np<8, VEC>(np, u, a, tiles, dev, s);
        break;
    case 4:
        if constexpr (VEC <= 2)
            return pick_np<4, VEC>(np, u, a, tiles, dev, s);
        break;
    case 2:
        if constexpr (VEC <= 2)
            return pick_np<2, VEC>(np, u, a, tiles, dev, s);
        break;
    case 1:
        if constexpr (VEC <= 2)
            return pick_np<1, VEC>(np, u, a, tiles, dev, s);
        break;
    }
    set_error("rows kernel: unsupported team shape");
    return SPMM_ERR_UNSUPPORTED;
}

int next_pow2(int x)
{
    int p = 1;
    while (p < x)
        p <<= 1;
    return p;
}
} // namespace

int launch_rows(const spmm_csr_s *A, int row_begin, int row_end, int c_row0, const double *d_B, long long ldb,
                double *d_C, long long ldc, int kc, cudaStream_t stream)
{
    if (row_end <= row_begin || kc <= 0)
        return SPMM_OK;
    const Tuning &t = tuning();
    const bool even = (kc % 2 == 0) && (ldb % 2 == 0) && (ldc % 2 == 0) &&
                      ((uintptr_t)d_B % 16 == 0) && ((uintptr_t)d_C % 16 == 0);
    int vec = even ? ((kc % 4 == 0 && kc >= 128) ? 4 : 2) : 1;
    if (t.rows_vec == 1 || (t.rows_vec == 2 && even) || (t.rows_vec == 4 && even && kc % 4 == 0 && kc >= 64))
        vec = t.rows_vec;
    const int kq = (kc + vec - 1) / vec; // lanes needed across the columns
    const int kl = std::min(32, next_pow2(kq));
    const int tiles = (kq + kl - 1) / kl;
    int np = t.rows_np > 0 ? t.rows_np : next_pow2(std::max(1, (int)(A->sched.mean_len * 0.5 + 0.5)));
    np = std::max(1, std::min(np, 32 / kl));
    const int u = t.rows_unroll > 0 ? t.rows_unroll : std::max(1, 4 / np);

    RowsArgs args;
    args.rowptr = A->d_rowptr;
    args.colidx = A->d_colidx;
    args.vals = A->d_vals;
    args.B = d_B;
    args.C = d_C;
    args.ldb = ldb;
    args.ldc = ldc;
    args.row_begin = row_begin;
    args.row_end = row_end;
    args.c_row0 = c_row0;
    args.kc = kc;
    switch (vec)
    {
    case 4:
        return pick_kl<4>(kl, np, u, args, tiles, A->device, stream);
    case 2:
        return pick_kl<2>(kl, np, u, args, tiles, A->device, stream);
    default:
        return pick_kl<1>(kl, np, u, args, tiles, A->device, stream);
    }
}

} // namespace spmm
