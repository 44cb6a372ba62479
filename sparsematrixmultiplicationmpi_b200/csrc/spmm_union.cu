// spmm_union.cu — row blocks over the union of their columns, B rows staged in a shared-memory window by TMA gather4.
//
// Replaces, for whole-matrix multiplies with k >= 32 on matrices whose neighbouring rows share columns, the inner loops of
//   /root/reference "Source Code/SparseMatrixFatVectorMultiply.cpp":17-28
// Why (profiles/r1_tiled.md, tools/microbench/lds_patterns.cu): the tiled kernel (spmm_tiled.cu) is bound by the SM's
// shared-memory data pipe — every non-zero reads its B row piece from shared memory (one wavefront per 128 bytes) and every
// byte staged costs a fill sector. This kernel cuts both:
//   * R consecutive rows (a block) are walked over the ascending union of their columns, so one B row piece read from
//     shared memory feeds R accumulators from registers (cop20k_A shape, R=2: 0.75 reads per non-zero);
//   * the k-tile is 32 columns (256-byte window rows), so the value/id stream is staged twice for k=64 instead of four times;
//   * the window holds single B rows in groups of four (cp.async.bulk.tensor tile::gather4): no box granularity, no pool.
// Work split (layout: spmm_union_build.h): an item = SL slots (one per team of 32/SL lanes) = the work of ONE consumer warp;
// consumer warp c takes items c, c+NCW, ... of the CTA's chunk; NPW producer warps walk all items, D items in flight.
// Every consumer passes every item's "landed" barrier in order (it may read window rows an earlier item loaded); barriers
// are indexed modulo 32 >= 2*D so that no barrier can run two phases ahead of a warp that still has to observe it.
// STATUS: experimental. Correct (tests/test_gpu_parity.py, tests/union_layout_emul.cpp) but 1.7x slower than the tiled kernel on
// the cop20k_A shape (profiles/r1_union.md: TMA issue cost per 1 KB gather, ~500 cycles of barrier hand-off per 2.5 KB item,
// too few items in flight for the consumers to hide their own latency). Never picked by AUTO.
// Per (row, column) the accumulation order is ascending column as in the reference; a block longer than the split length
// is cut into segments whose partial sums are folded in a fixed order (within the 1e-12 tolerance, like the merge kernel).
// Absent entries of the union are 0.0 values: B is assumed finite (0 * inf would leak into a row that does not hold
// that column) — the same assumption spmm_rowblock.cu documents.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "spmm_launch.cuh"
#include "spmm_union_build.h"

namespace spmm
{

namespace
{
constexpr int SMEM_CAP = 232448; // 227 KB opt-in limit per CTA on sm_100
constexpr unsigned RING_OFF = 1024;
constexpr int BAR_MAX = 32;

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// four rows of B (any row ids) x the box width into four consecutive smem rows
__device__ __forceinline__ void tma_gather4(unsigned dst, const CUtensorMap *map, int c0, int4 rows, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_box(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
// a group of four B rows: one box copy when the rows are consecutive, else a gather
__device__ __forceinline__ void tma_group(unsigned dst, const CUtensorMap *row_map, const CUtensorMap *box_map, int c0, int4 rows,
                                          unsigned bar, bool boxes)
{
    if (boxes && rows.y == rows.x + 1 && rows.z == rows.x + 2 && rows.w == rows.x + 3)
        tma_box(dst, box_map, c0, rows.x, bar);
    else
        tma_gather4(dst, row_map, c0, rows, bar);
}
__device__ __forceinline__ double2 lds128d(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64u(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds32u(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds16u(unsigned addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds8u(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

struct UnionArgs
{
    ExtraDst extra; // peer copies of C (element offsets from C)
    const unsigned char *blob;
    const UItem *items;
    const int *chunk_first;
    const int4 *gcols;
    const int *gslot;
    double *C;
    long long ldc;
    int n_rows, kc, nkt, D, maxg;
    long long *prof; // dbg & 64: per CTA 8 counters
    int dbg; // diagnostics (wrong results): 1 consumers skip the arithmetic, 2 no B rows are staged, 4 no blobs either
    unsigned slab_off; // offset of the window from the aligned smem base
};

template <int R, int KT, int SL, int NCW, int NPW>
__global__ void __launch_bounds__((NCW + NPW) * 32, 1)
    spmm_union_kernel(const UnionArgs a, const __grid_constant__ CUtensorMap row_map, const __grid_constant__ CUtensorMap box_map)
{
    constexpr int TLN = 32 / SL;       // lanes per team
    constexpr int NL = KT / (2 * TLN); // LDS.128 per lane and entry
    constexpr int U = 4;               // entries in flight per team (ids come four at a time)
    constexpr unsigned GB = 4u * KT * 8u; // bytes of a window group (4 B rows)
    constexpr unsigned HDR = (SL * 4 + 12 + 15) & ~15;
    static_assert(NL >= 1 && NL * 2 * TLN == KT, "team must cover the k-tile");
    static_assert(TLN == 8 || NL % 2 == 0, "4-lane teams swap 64-byte halves between neighbours");
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned s0 = (smem_u32(smem) + 1023u) & ~1023u;
    const unsigned full = s0, empty = s0 + 8 * BAR_MAX;
    const unsigned s_ring = s0 + RING_OFF, s_slab = s0 + a.slab_off;
    const int D = a.D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x / a.nkt, k0 = (blockIdx.x % a.nkt) * KT;
    const int first = a.chunk_first[chunk], n = a.chunk_first[chunk + 1] - first;
    if (threadIdx.x == 0)
    {
        for (int i = 0; i < BAR_MAX; ++i)
        {
            mbar_init(full + 8 * i, 1);  // the expect_tx arrive of producer warp 0
            mbar_init(empty + 8 * i, 1); // the item's consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (n <= 0)
        return;
    const UItem *items = a.items + first;

    if (warp >= NCW)
    {
        // ---------------- producers: every warp walks all items and issues its share of the gathers ----------------
        // Descriptors and gather lists come from global memory a batch of 32 items at a time: lane i of every producer
        // warp holds item base+i (its descriptor and this warp's gathers of it: groups pw, pw+NPW, ...) and issues that
        // item's copies itself when the warp gets there; the next batch is in flight while this one is walked. (A
        // register ring refilled every item stalls on the newest load: loads share scoreboards.)
        const int pw = warp - NCW;
        constexpr int GP = 6; // gathers per warp and item held in registers (further ones are read when needed)
        struct Batch
        {
            UItem d;
            int4 gc[GP];
            int gs[GP];
        };
        const int4 *gcols = a.gcols + (size_t)first * a.maxg;
        const int *gslot = a.gslot + (size_t)first * a.maxg;
        auto load_batch = [&](Batch &bt, int base) {
            const int idx = base + lane;
            bt.d.n_groups = 0;
            bt.d.drain = 0;
            if (idx < n)
            {
                bt.d = items[idx];
#pragma unroll
                for (int t = 0; t < GP; ++t)
                    if (pw + NPW * t < a.maxg)
                    {
                        bt.gc[t] = gcols[(size_t)idx * a.maxg + pw + NPW * t];
                        bt.gs[t] = gslot[(size_t)idx * a.maxg + pw + NPW * t];
                    }
            }
        };
        Batch cur, nxt;
        load_batch(cur, 0);
        load_batch(nxt, 32);
        int waited = 0; // items < waited have finished
        long long p_wait = 0, p_issue = 0, p_batch = 0;
        const long long p_t0 = clock64();
        for (int base = 0; base < n; base += 32)
        {
            const unsigned drains = __ballot_sync(0xFFFFFFFFu, cur.d.drain != 0);
            const int cnt = min(32, n - base);
            for (int i = 0; i < cnt; ++i)
            {
                const int w = base + i;
                // item w overwrites its ring piece and window groups: whatever it overwrites was last read by item w-D at
                // the latest (the builder's rule), or — drain — by any earlier item
                const int target = ((drains >> i) & 1u) ? w : w - D + 1;
                const long long c0 = clock64();
                for (; waited < target; ++waited)
                    mbar_wait(empty + 8 * (waited % BAR_MAX), (waited / BAR_MAX) & 1);
                const long long c1 = clock64();
                if (lane == i)
                {
                    const unsigned bar = full + 8 * (w % BAR_MAX);
                    if (pw == 0)
                    {
                        // (gathers of the other warps may complete first: the transaction count of an mbarrier may run
                        // negative, and the phase cannot end before this arrive)
                        mbar_expect_tx(bar, ((a.dbg & 4) ? 0u : cur.d.bytes) + ((a.dbg & 2) ? 0u : (unsigned)cur.d.n_groups * GB));
                        if (!(a.dbg & 4))
                            bulk_g2s(s_ring + cur.d.ring_off, a.blob + cur.d.blob_off, cur.d.bytes, bar);
                    }
                    if (!(a.dbg & 2))
                    {
#pragma unroll
                        for (int t = 0; t < GP; ++t)
                            if (pw + NPW * t < cur.d.n_groups)
                                tma_group(s_slab + (unsigned)cur.gs[t] * GB, &row_map, &box_map, k0, cur.gc[t], bar, !(a.dbg & 8));
                        for (int g = pw + NPW * GP; g < cur.d.n_groups; g += NPW) // long lists: rare
                            tma_group(s_slab + (unsigned)gslot[(size_t)w * a.maxg + g] * GB, &row_map, &box_map, k0,
                                      gcols[(size_t)w * a.maxg + g], bar, !(a.dbg & 8));
                    }
                }
                __syncwarp();
                const long long c2 = clock64();
                p_wait += c1 - c0;
                p_issue += c2 - c1;
            }
            const long long c3 = clock64();
            cur = nxt;
            load_batch(nxt, base + 64);
            // the blobs of the batch after the next one: into L2 now (they are read once, from HBM)
            if (pw == NPW - 1 && !(a.dbg & 16) && base + 64 + lane < n)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.blob + nxt.d.blob_off), "r"(nxt.d.bytes) : "memory");
            __syncwarp();
            p_batch += clock64() - c3;
        }
        if ((a.dbg & 64) && lane == 0)
        {
            long long *pr = a.prof + (size_t)blockIdx.x * 16 + pw * 4;
            if (pw < 2)
            {
                pr[0] = p_wait;
                pr[1] = p_issue;
                pr[2] = p_batch;
                pr[3] = clock64() - p_t0;
            }
        }
        return;
    }

    // ---------------- consumers: warp c owns items c, c+NCW, ... ----------------
    const int tw = lane / TLN, l = lane % TLN;
    const int tq = (TLN == 4) ? (tw & 1) : 0; // 4-lane teams: the two teams of a quarter-warp read opposite 64-byte halves
    int colo[NL];                              // column (in doubles) of accumulator x inside the k-tile
#pragma unroll
    for (int x = 0; x < NL; ++x)
        colo[x] = ((x ^ tq) * TLN + l) * 2;
    int seen = 0; // items whose "landed" barrier this warp has passed
    unsigned ring_next = warp < n ? items[warp].ring_off : 0u;
    long long q_wait = 0, q_work = 0;
    const long long q_t0 = clock64();
    for (int w = warp; w < n; w += NCW)
    {
        const long long c0 = clock64();
        const unsigned ring_off = ring_next;
        if (w + NCW < n)
            ring_next = items[w + NCW].ring_off; // in flight while this item is computed
        // every item up to w has landed (all lanes wait: one waiting lane + a warp barrier measured slower)
        for (; seen <= w; ++seen)
            mbar_wait(full + 8 * (seen % BAR_MAX), (seen / BAR_MAX) & 1);
        const long long c1 = clock64();
        q_wait += c1 - c0;
        if (a.dbg & 4) // diagnostics: protocol only
        {
            __syncwarp();
            if (lane == 0)
                mbar_arrive(empty + 8 * (w % BAR_MAX));
            continue;
        }
        const unsigned blob = s_ring + ring_off;
        const int len = (int)lds16u(blob + tw * 2);
        const unsigned blk = lds8u(blob + SL * 2 + tw), nseg = lds8u(blob + SL * 3 + tw);
        const int row0 = (int)lds32u(blob + SL * 4), steps = (int)lds32u(blob + SL * 4 + 4);
        const unsigned has_split = lds32u(blob + SL * 4 + 8);
        const unsigned ids_s = blob + HDR + tw * 8;
        const unsigned vals_s = blob + HDR + ((steps + 3) >> 2) * (SL * 8) + tw * (R * 8);
        double2 acc[R][NL];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int x = 0; x < NL; ++x)
                acc[r][x] = make_double2(0.0, 0.0);
        for (int s = 0; s < ((a.dbg & 1) ? 0 : steps); s += U)
        {
            unsigned id[U];
            {
                uint2 four = make_uint2(0u, 0u);
                if (s < len)
                    four = lds64u(ids_s + (s >> 2) * (SL * 8));
                id[0] = four.x & 0xFFFFu;
                id[1] = four.x >> 16;
                id[2] = four.y & 0xFFFFu;
                id[3] = four.y >> 16;
            }
            double v[U][R];
#pragma unroll
            for (int q = 0; q < U; ++q)
            {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    v[q][r] = 0.0;
                if (s + q < len)
                {
#pragma unroll
                    for (int r = 0; r < R; r += 2)
                    {
                        const double2 t = lds128d(vals_s + (unsigned)(s + q) * (SL * R * 8) + r * 8);
                        v[q][r] = t.x;
                        v[q][r + 1] = t.y;
                    }
                }
            }
            double2 b[U][NL];
#pragma unroll
            for (int q = 0; q < U; ++q)
#pragma unroll
                for (int x = 0; x < NL; ++x)
                {
                    b[q][x] = make_double2(0.0, 0.0);
                    if (s + q < len)
                        b[q][x] = lds128d(s_slab + id[q] * (KT * 8) + colo[x] * 8);
                }
            // (a software-pipelined version of this loop — next step's loads issued before these FMAs — measured
            // slower: 168 registers and the copies between the two stages; profiles/r1_union.md)
#pragma unroll
            for (int q = 0; q < U; ++q)
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int x = 0; x < NL; ++x)
                    {
                        acc[r][x].x = fma(v[q][r], b[q][x].x, acc[r][x].x);
                        acc[r][x].y = fma(v[q][r], b[q][x].y, acc[r][x].y);
                    }
        }
        // column-aligned view: accumulator x of an odd 4-lane team sits at the columns of accumulator x^1 of an even one
        double2 al[R][NL];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int x = 0; x < NL; ++x)
                al[r][x] = tq ? acc[r][(TLN == 4 && NL > 1) ? (x ^ 1) : x] : acc[r][x];
        if (has_split)
        {
            // the segments of a long block sit in consecutive slots: the head adds the partial sums of its continuations
            // in slot order (fixed order: deterministic)
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int x = 0; x < NL; ++x)
                {
                    double2 sum = al[r][x];
#pragma unroll
                    for (int j = 1; j < SL; ++j)
                    {
                        const double ox = __shfl_down_sync(0xFFFFFFFFu, al[r][x].x, j * TLN);
                        const double oy = __shfl_down_sync(0xFFFFFFFFu, al[r][x].y, j * TLN);
                        if (j < (int)nseg)
                        {
                            sum.x += ox;
                            sum.y += oy;
                        }
                    }
                    al[r][x] = sum;
                }
        }
        if (blk != 0xFFu && nseg != 0)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
            {
                const int row = row0 + (int)blk * R + r;
                if (row < a.n_rows)
                {
                    double *cr = a.C + (long long)row * a.ldc + k0;
#pragma unroll
                    for (int x = 0; x < NL; ++x)
                    {
                        const int col = (x * TLN + l) * 2;
                        if (k0 + col < a.kc)
                        {
                            st_c2(cr + col, al[r][x].x, al[r][x].y);
                            for (int d = 0; d < a.extra.n; ++d)
                                st_c2(cr + a.extra.off[d] + col, al[r][x].x, al[r][x].y);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(empty + 8 * (w % BAR_MAX)); // this item's ring piece is free; its window groups age normally
        q_work += clock64() - c1;
    }
    if ((a.dbg & 64) && lane == 0 && warp < 2)
    {
        long long *pr = a.prof + (size_t)blockIdx.x * 16 + 8 + warp * 4;
        pr[0] = q_wait;
        pr[1] = q_work;
        pr[2] = n;
        pr[3] = clock64() - q_t0;
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int encode_row_map(CUtensorMap *map, const double *d_B, long long ldb, int kc, int n_cols, int kt, int rows = 1)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc)
    {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SPMM_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)kc, (cuuint64_t)n_cols};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldb * 8};
    const cuuint32_t box[2] = {(cuuint32_t)kt, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(d_B), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)cr));
        return SPMM_ERR_CUDA;
    }
    return SPMM_OK;
}

template <int R, int KT, int SL, int NCW, int NPW>
int launch_union_t(const UnionDev *u, const double *d_B, long long ldb, double *d_C, long long ldc, int kc, int n_rows,
                   int n_cols, cudaStream_t stream, const ExtraDst &extra)
{
    auto kern = spmm_union_kernel<R, KT, SL, NCW, NPW>;
    static std::mutex mu;
    static bool configured = false;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!configured)
        {
            SPMM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
            configured = true;
        }
    }
    CUtensorMap row_map, box_map;
    int rc = encode_row_map(&row_map, d_B, ldb, kc, n_cols, KT);
    if (!rc)
        rc = encode_row_map(&box_map, d_B, ldb, kc, n_cols, KT, 4);
    if (rc)
        return rc;
    UnionArgs a = {};
    a.extra = extra;
    a.blob = u->d_blob;
    a.items = reinterpret_cast<const UItem *>(u->d_items);
    a.chunk_first = u->d_chunk_first;
    a.gcols = reinterpret_cast<const int4 *>(u->d_gcols);
    a.gslot = u->d_gslot;
    a.C = d_C;
    a.ldc = ldc;
    a.n_rows = n_rows;
    a.kc = kc;
    a.nkt = (kc + KT - 1) / KT;
    a.D = u->D;
    a.maxg = u->maxg;
    a.dbg = tuning().union_debug;
    a.slab_off = (unsigned)u->slab_off;
    const size_t smem = (size_t)u->slab_off + (size_t)u->NG * 4 * KT * 8 + 1024;
    static long long *d_prof = nullptr;
    static int prof_calls = 0;
    const int grid = u->n_chunks * a.nkt;
    if ((a.dbg & 64) && !d_prof)
        SPMM_CUDA(cudaMalloc(&d_prof, sizeof(long long) * 16 * 4096));
    a.prof = d_prof;
    kern<<<grid, (NCW + NPW) * 32, smem, stream>>>(a, row_map, box_map);
    SPMM_CUDA(cudaGetLastError());
    if ((a.dbg & 64) && ++prof_calls % 20 == 0 && grid <= 4096)
    {
        std::vector<long long> h((size_t)16 * grid);
        SPMM_CUDA(cudaStreamSynchronize(stream));
        SPMM_CUDA(cudaMemcpy(h.data(), d_prof, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost));
        double sm[16] = {0};
        long long mx = 0, mn = 1ll << 60, mxi = 0, mni = 1 << 30;
        for (int i = 0; i < grid; ++i)
        {
            for (int j = 0; j < 16; ++j)
                sm[j] += (double)h[(size_t)i * 16 + j] / grid;
            mx = std::max(mx, h[(size_t)i * 16 + 11]);
            mn = std::min(mn, h[(size_t)i * 16 + 11]);
            mxi = std::max(mxi, h[(size_t)i * 16 + 10]);
            mni = std::min(mni, h[(size_t)i * 16 + 10]);
        }
        fprintf(stderr, "[union prof] consumer0 total min %lld max %lld, items min %lld max %lld\n", mn, mx, mni, mxi);
        fprintf(stderr,
                "[union prof dbg=%d] per CTA (clk): producer0 wait %.0f issue %.0f batch %.0f total %.0f | producer1 wait %.0f issue "
                "%.0f batch %.0f total %.0f | consumer0 wait %.0f work %.0f items %.0f total %.0f | consumer1 wait %.0f work %.0f total %.0f\n",
                a.dbg, sm[0], sm[1], sm[2], sm[3], sm[4], sm[5], sm[6], sm[7], sm[8], sm[9], sm[10], sm[11], sm[12], sm[13], sm[15]);
    }
    return SPMM_OK;
}

} // namespace

void free_union(spmm_csr_s *A)
{
    UnionDev *u = A->un;
    if (!u)
        return;
    cudaFree(u->d_blob);
    cudaFree(u->d_items);
    cudaFree(u->d_chunk_first);
    cudaFree(u->d_gcols);
    cudaFree(u->d_gslot);
    delete u;
    A->un = nullptr;
}

bool union_shape_ok(const spmm_csr_s *A, const double *d_B, long long ldb, const double *d_C, long long ldc, int kc)
{
    const UnionDev *u = A->un;
    return u && kc >= 2 && kc % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0 && ((uintptr_t)d_B % 16 == 0) &&
           ((uintptr_t)d_C % 16 == 0) && (unsigned long long)ldb * 8ull < (1ull << 40);
}

int launch_union(const spmm_csr_s *A, const double *d_B, long long ldb, double *d_C, long long ldc, int kc,
                 cudaStream_t stream, const ExtraDst *extra)
{
    const UnionDev *u = A->un;
    const ExtraDst x = extra ? *extra : ExtraDst();
#define SPMM_UNION_CASE(RR, KK, SS, NN, PP)                                                              \
    if (u->R == RR && u->KT == KK && u->SL == SS && u->NCW == NN && u->NPW == PP)                        \
        return launch_union_t<RR, KK, SS, NN, PP>(u, d_B, ldb, d_C, ldc, kc, A->n_rows, A->n_cols, stream, x);
    SPMM_UNION_CASE(2, 32, 4, 4, 4)
    SPMM_UNION_CASE(2, 32, 4, 6, 4)
    SPMM_UNION_CASE(2, 32, 4, 8, 4)
    SPMM_UNION_CASE(2, 32, 4, 6, 8)
    SPMM_UNION_CASE(2, 32, 4, 8, 8)
    SPMM_UNION_CASE(2, 32, 8, 4, 4)
    SPMM_UNION_CASE(2, 32, 8, 4, 8)
    SPMM_UNION_CASE(4, 32, 4, 4, 4)
    SPMM_UNION_CASE(4, 32, 4, 4, 8)
    SPMM_UNION_CASE(2, 16, 4, 8, 4)
    SPMM_UNION_CASE(2, 16, 4, 12, 4)
#undef SPMM_UNION_CASE
    set_error("union kernel: no instantiation for this (rows per block, k-tile, slots, consumer warps)");
    return SPMM_ERR_INVALID;
}

} // namespace spmm

using namespace spmm;

extern "C"
{

int spmm_csr_build_union(spmm_csr_t A, int rows_per_block, int k)
{
    SPMM_REQUIRE(A != nullptr, "handle is NULL");
    SPMM_REQUIRE(rows_per_block == 0 || rows_per_block == -1 || rows_per_block == 2 || rows_per_block == 4,
                 "rows_per_block must be 0 (drop), -1 (auto), 2 or 4");
    SPMM_REQUIRE(k >= 0, "k is negative");
    SPMM_CUDA(cudaSetDevice(A->device));
    free_union(A);
    if (rows_per_block == 0 || A->n_rows == 0 || A->nnz == 0)
        return SPMM_OK;
    const Tuning &tn = tuning();
    UnionParams p;
    p.R = rows_per_block > 0 ? rows_per_block : 2;
    p.KT = tn.tiled_kt == 16 ? 16 : 32;
    p.slots = tn.union_slots == 8 ? 8 : 4;
    const int ncw = tn.tiled_ncw > 0 ? tn.tiled_ncw : 6;
    p.D = tn.tiled_depth > 0 ? tn.tiled_depth : ncw + 2;
    SPMM_REQUIRE(p.D >= ncw && 2 * p.D <= BAR_MAX, "union layout: depth must be at least the consumer warps and at most 16");
    p.split_len = tn.union_split > 0 ? tn.union_split : 48;
    const int nkt = std::max(1, (std::max(k, 1) + p.KT - 1) / p.KT);
    const int sms = device_props(A->device).sm_count;
    p.n_chunks = tn.tiled_chunk > 0 ? tn.tiled_chunk : std::max(1, sms / nkt);
    p.smem_bytes = SMEM_CAP - 3072; // barriers, alignment of the base and of the window
    p.max_groups = tn.tiled_ns;
    // the builder runs on the host: fetch the CSR arrays
    std::vector<int> rp((size_t)A->n_rows + 1), ci((size_t)A->nnz);
    std::vector<double> va((size_t)A->nnz);
    SPMM_CUDA(cudaMemcpy(rp.data(), A->d_rowptr, sizeof(int) * rp.size(), cudaMemcpyDeviceToHost));
    SPMM_CUDA(cudaMemcpy(ci.data(), A->d_colidx, sizeof(int) * ci.size(), cudaMemcpyDeviceToHost));
    SPMM_CUDA(cudaMemcpy(va.data(), A->d_vals, sizeof(double) * va.size(), cudaMemcpyDeviceToHost));
    UnionLayout L;
    if (build_union_layout(A->n_rows, A->n_cols, rp.data(), ci.data(), va.data(), p, &L))
    {
        set_error(L.error);
        return SPMM_ERR_UNSUPPORTED;
    }
    UnionDev *u = new UnionDev();
    A->un = u;
    u->R = p.R;
    u->KT = p.KT;
    u->SL = p.slots;
    u->NCW = ncw;
    u->NPW = tn.tiled_npw == 8 ? 8 : 4;
    u->maxg = L.maxg;
    u->D = p.D;
    u->NG = L.NG;
    u->n_chunks = L.p.n_chunks;
    u->n_items = L.n_items;
    u->nkt = nkt;
    u->ring_bytes = L.ring_bytes;
    u->slab_off = (int)((RING_OFF + (size_t)L.ring_bytes + 1023) & ~(size_t)1023);
    u->staged_rows = L.staged_rows;
    u->union_entries = L.union_entries;
    u->slot_steps = L.slot_steps;
    u->drains = L.drains;
    auto up = [&](void **dst, const void *src, size_t bytes) -> int {
        SPMM_CUDA(cudaMalloc(dst, std::max<size_t>(bytes, 16)));
        if (bytes)
            SPMM_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return SPMM_OK;
    };
    int rc = up((void **)&u->d_blob, L.blob.data(), L.blob.size());
    if (!rc) rc = up(&u->d_items, L.items.data(), sizeof(UItem) * L.items.size());
    if (!rc) rc = up((void **)&u->d_chunk_first, L.chunk_first.data(), sizeof(int) * L.chunk_first.size());
    if (!rc) rc = up(&u->d_gcols, L.gcols.data(), sizeof(int) * L.gcols.size());
    if (!rc) rc = up((void **)&u->d_gslot, L.gslot.data(), sizeof(int) * L.gslot.size());
    if (rc)
    {
        free_union(A);
        return rc;
    }
    if (getenv("SPMM_TILED_DEBUG"))
        fprintf(stderr,
                "[union build] R=%d KT=%d SL=%d NCW=%d D=%d chunks=%d items=%d NG=%d ring=%d union/nnz=%.3f slot-steps/union=%.3f "
                "staged/N=%.2f drains=%d blob=%.1f MB\n",
                u->R, u->KT, u->SL, u->NCW, u->D, u->n_chunks, u->n_items, u->NG, u->ring_bytes,
                (double)L.union_entries / (double)A->nnz, (double)L.slot_steps / (double)std::max(1ll, L.union_entries),
                (double)L.staged_rows / A->n_rows, L.drains, L.blob.size() / 1e6);
    return SPMM_OK;
}

int spmm_csr_union_info(spmm_csr_t A, int *rows_per_block, int *k_tile, int *window_rows, double *union_per_nnz,
                        double *padding, double *staged_per_row)
{
    SPMM_REQUIRE(A != nullptr, "handle is NULL");
    const UnionDev *u = A->un;
    if (rows_per_block) *rows_per_block = u ? u->R : 0;
    if (k_tile) *k_tile = u ? u->KT : 0;
    if (window_rows) *window_rows = u ? u->NG * 4 : 0;
    if (union_per_nnz) *union_per_nnz = (u && A->nnz) ? (double)u->union_entries / (double)A->nnz : 0.0;
    if (padding) *padding = (u && u->union_entries) ? (double)u->slot_steps / (double)u->union_entries : 0.0;
    if (staged_per_row) *staged_per_row = (u && A->n_rows) ? (double)u->staged_rows / (double)A->n_rows : 0.0;
    return SPMM_OK;
}

} // extern "C"
