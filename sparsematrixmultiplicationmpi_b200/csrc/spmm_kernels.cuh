// spmm_kernels.cuh — sm_100a SpMM kernels: C = A * B, A CSR (FP64 values, int32 ids),
// B/C dense row-major FP64 ("fat vectors").
//
// Replaces the reference's host triple loop
//   /root/reference "Source Code/SparseMatrixFatVectorMultiply.cpp":17-28
// and the per-rank loops of the three strategies (RowWise.cpp:36-50,
// ColumnWise.cpp:34-48, NonZeroElement.cpp:56-67).
//
// The path is not a dense contraction: no tensor cores. It is bound by HBM
// (A streamed once, B and C touched once) and, at large k, by the L1/L2 -> SM
// gather of B rows. Design rules used below (DESIGN.md has the arithmetic):
//   * lanes map across the k columns of a B row, 16 B (double2) per lane, so a B
//     row is one fully coalesced request; several non-zeros / rows share a warp
//     when k is small ("team" = KL lanes x NP concurrent non-zeros);
//   * every CTA owns ONE CONTIGUOUS chunk of rows, cut so that chunks cost the
//     same (nnz + row overhead), and the grid is one resident wave
//     (SMs x CTAs/SM): neighbouring rows of FEM-like matrices share B rows, and
//     a contiguous sweep turns that into L1 hits instead of L2 traffic;
//   * A (col ids, values) is streamed with L1::no_allocate so it does not evict
//     B rows from L1; C is written with streaming stores;
//   * UNROLL independent B-row loads are in flight per team before the FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spmm
{

// ---- cache-hinted accessors ------------------------------------------------------
__device__ __forceinline__ int ld_stream_i32(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_b1(const double *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_b2(const double *p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_c1(double *p, double v)
{
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_c2(double *p, double x, double y)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}

// VEC contiguous doubles of one B row / C row per lane.
template <int VEC>
struct Vec
{
    double v[VEC];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            v[i] = 0.0;
    }
    __device__ __forceinline__ void load(const double *p)
    {
        if constexpr (VEC == 1)
            v[0] = ld_b1(p);
        else
        {
#pragma unroll
            for (int i = 0; i < VEC; i += 2)
            {
                double2 t = ld_b2(p + i);
                v[i] = t.x;
                v[i + 1] = t.y;
            }
        }
    }
    __device__ __forceinline__ void store(double *p) const
    {
        if constexpr (VEC == 1)
            st_c1(p, v[0]);
        else
        {
#pragma unroll
            for (int i = 0; i < VEC; i += 2)
                st_c2(p + i, v[i], v[i + 1]);
        }
    }
    __device__ __forceinline__ void fma(double a, const Vec &b)
    {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            v[i] = ::fma(a, b.v[i], v[i]);
    }
};

// Cost model used to cut rows into equal-cost contiguous chunks: one unit per
// non-zero plus ROW_COST units per row (row-pointer loads, C store, loop set-up).
constexpr int ROW_COST = 4;

__device__ __forceinline__ long long chunk_cost(const int *__restrict__ rowptr, int row_begin, int r)
{
    return (long long)(rowptr[r] - rowptr[row_begin]) + (long long)ROW_COST * (r - row_begin);
}

// First row r in [row_begin,row_end] whose prefix cost reaches `target`.
__device__ __forceinline__ int chunk_lower_bound(const int *__restrict__ rowptr, int row_begin, int row_end,
                                                 long long target)
{
    int lo = row_begin, hi = row_end;
    while (lo < hi)
    {
        int mid = lo + ((hi - lo) >> 1);
        if (chunk_cost(rowptr, row_begin, mid) < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

struct RowsArgs
{
    const int *rowptr;
    const int *colidx;
    const double *vals;
    const double *B; // already offset to the first computed column
    double *C;       // already offset to the first computed column; row `c_row0` is at C[0]
    long long ldb, ldc;
    int row_begin, row_end; // rows computed by this launch
    int c_row0;             // row id stored at C[0]
    int kc;                 // columns computed (per blockIdx.y tile: KL*VEC of them)
};

// Team kernel. A team of KL*NP lanes owns one row at a time: KL lanes across the
// columns (VEC doubles each), NP non-zeros of the row in flight side by side,
// UNROLL steps issued back to back. 32/(KL*NP) teams share a warp.
template <int KL, int VEC, int NP, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) spmm_rows_kernel(const RowsArgs a)
{
    constexpr int T = KL * NP;
    constexpr int RW = 32 / T;
    constexpr int SLOTS = (THREADS / 32) * RW;
    static_assert(T <= 32 && 32 % T == 0, "team must divide a warp");

    __shared__ int s_chunk[2];
    if (threadIdx.x == 0)
    {
        const long long total = chunk_cost(a.rowptr, a.row_begin, a.row_end);
        const long long g = gridDim.x, b = blockIdx.x;
        s_chunk[0] = b == 0 ? a.row_begin : chunk_lower_bound(a.rowptr, a.row_begin, a.row_end, (total * b + g - 1) / g);
        s_chunk[1] = b == g - 1 ? a.row_end : chunk_lower_bound(a.rowptr, a.row_begin, a.row_end, (total * (b + 1) + g - 1) / g);
    }
    __syncthreads();
    const int lo = s_chunk[0], hi = s_chunk[1];

    const int lane = threadIdx.x & 31;
    const int lt = lane % T;   // lane inside the team
    const int g = lt / KL;     // which of the NP concurrent non-zeros
    const int kl = lt % KL;    // which column group
    const int slot = (threadIdx.x >> 5) * RW + lane / T;
    const int kcol = blockIdx.y * (KL * VEC) + kl * VEC;
    const bool kact = kcol < a.kc;
    const double *__restrict__ Bk = a.B + kcol;

    for (int base = lo; base < hi; base += SLOTS)
    {
        const int row = base + slot;
        const bool valid = row < hi;
        int js = 0, je = 0;
        if (valid)
        {
            js = a.rowptr[row];
            je = a.rowptr[row + 1];
        }
        Vec<VEC> acc;
        acc.zero();
        for (int j = js + g; j < je; j += NP * UNROLL)
        {
            int c[UNROLL];
            double x[UNROLL];
            Vec<VEC> b[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
            {
                const int jj = j + u * NP;
                c[u] = 0;
                x[u] = 0.0;
                if (jj < je)
                {
                    c[u] = ld_stream_i32(a.colidx + jj);
                    x[u] = ld_stream_f64(a.vals + jj);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
            {
                b[u].zero();
                if (kact && (j + u * NP) < je)
                    b[u].load(Bk + (long long)c[u] * a.ldb);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if ((j + u * NP) < je)
                    acc.fma(x[u], b[u]);
        }
        if constexpr (NP > 1)
        {
#pragma unroll
            for (int off = KL; off < T; off <<= 1)
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    acc.v[i] += __shfl_xor_sync(0xffffffffu, acc.v[i], off);
        }
        if (valid && g == 0 && kact)
            acc.store(a.C + (long long)(row - a.c_row0) * a.ldc + kcol);
    }
}

} // namespace spmm
