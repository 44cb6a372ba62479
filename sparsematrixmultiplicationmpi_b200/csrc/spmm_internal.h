// spmm_internal.h — host-side structures shared by the translation units of libspmm_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "spmm_b200.h"

namespace spmm
{

void set_error(const std::string &msg);
void note_kernel(const char *name); // kernel family of the calling thread's last launch (spmm_last_kernel_name)
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SPMM_CUDA(call)                                                        \
    do                                                                         \
    {                                                                          \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess)                                                \
            return ::spmm::cuda_fail(e__, #call, __FILE__, __LINE__);          \
    } while (0)

#define SPMM_REQUIRE(cond, msg)                                                \
    do                                                                         \
    {                                                                          \
        if (!(cond))                                                           \
        {                                                                      \
            ::spmm::set_error(std::string("invalid argument: ") + (msg));     \
            return SPMM_ERR_INVALID;                                           \
        }                                                                      \
    } while (0)

// Row-length schedule of a CSR handle (built once on the device).
struct Schedule
{
    long long bins[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // 0, 1-2, 3-4, 5-8, 9-16, 17-32, 33-256, >256
    int max_len = 0;
    double mean_len = 0.0;
    int auto_kernel = SPMM_KERNEL_ROWS;
};

// Tuning knobs (0 = choose automatically). Set through spmm_tune_set for experiments.
struct Tuning
{
    int rows_kl = 0;        // lanes across the columns of a B row
    int rows_nv = 0;        // accesses per lane (1, 2, 4)
    int rows_np = 0;        // concurrent non-zeros of one row per team
    int rows_unroll = 0;    // steps in flight per team
    int rows_vec = 0;       // 1 forces 8-byte accesses
    int rows_ctas_per_sm = 0;
    int merge_items = 0;    // merge-path items per team
    int merge_wave = 0;     // merge-path: warps per SM the item count is shrunk for on small shards (0 = 64)
    int rows_sweep = 0;     // 1: one CTA per SM walks the column tiles itself (L1-resident window)
    int rows_threads = 0;   // sweep kernels: 512 or 1024 threads
    int rows_tile = 0;      // 0 auto, > 0 rows per round-robin tile, -1 never tile (one chunk per CTA)
    int rows_prefetch = -1; // -1 auto; bit0 A chunk, bit1 B share: TMA prefetch into L2 at kernel start
    int tiled = -1;         // -1 auto, 0 never use the tile layout, 1 always when built
    int tiled_kt = 0;       // tiled kernel: k-tile width (16, 32)
    int tiled_ncw = 0;      // consumer warps (8, 12, 16)
    int tiled_unroll = 0;   // records in flight per team (2, 4)
    int tiled_thr = 0;      // build: non-zeros a box needs in a tile to be loaded as a box (else single rows)
    int tiled_chunk = 0;    // build: tiles per chunk
    int tiled_depth = 0;    // build: work items in flight (2..8)
    int tiled_pool = 0;     // build: rows of the singles pool
    int tiled_ns = 0;       // build: cap on the window slots
    int tiled_ksplit = 0;   // build: CTAs that share one chunk, each taking a group of k-tiles (0/1 none)
    int tiled_npw = 0;      // launch: producer warps (4, 8)
    int host_slabs = 0;     // host-buffer multiply: k-slabs in the PCIe pipeline (0 auto, 1 none)
    int tiled_auto_after = 8; // AUTO builds the tile layout on the multiply after this many whole-matrix multiplies of a handle
                              // (the build costs 4-7 ms of wall clock on cfg2 at k = 16 and 14-25 ms at k = 64 — five device passes of 1.9 ms
                              // plus allocations, tools/build_tiles_probe.py — and saves ~50 us per multiply)
    int tiled_pdl = 1;      // launch: programmatic dependent launch of the tiled kernel (0 off)
    int tiled_group = 0;    // build: tiles one far-band apart walked in turn, this many bands per group (0 auto = 2, 1 off)
    int tiled_stride = 0;   // build: far-band distance in rows (0 = detect from the matrix)
    int tiled_prefetch = -1; // launch: tiles ahead whose blob is prefetched into L2 (-1 auto)
    int stream = -1;        // -1 auto, 0 AUTO never uses the stream kernel (k = 1, 2, 4, 8)
    int stream_auto_kmax = 1; // AUTO takes the stream kernel up to this k ...
    int stream_auto_min_nnz = 4 << 20; // ... from this many non-zeros (measured: equal to the row kernel at 2.6 M non-zeros, 43 against 56 us at 10.5 M; profiles/r1_stream.md)
    int stream_tile = 0;    // stream kernel: non-zeros per tile (0 = 4096 / k)
};
Tuning &tuning();

struct DeviceProps
{
    int sm_count = 0;
    long long l2_bytes = 0;
};
const DeviceProps &device_props(int device);
// fn(i) for i in [0, n) on the library's host threads, the caller taking part (spmm_host.cu)
void host_parallel_for(int n, const std::function<void(int)> &fn);

} // namespace spmm

struct spmm_csr_s
{
    int device = 0;
    int n_rows = 0, n_cols = 0;
    long long nnz = 0;
    int *d_rowptr = nullptr;
    int *d_colidx = nullptr;
    double *d_vals = nullptr;
    bool owns = false;
    spmm::Schedule sched;
    // staging of the host-buffer entry points (spmm_host.cu): pinned mirrors of B and C, their device images, streams
    std::mutex host_mu; // one host-buffer call per handle at a time
    double *d_B = nullptr, *d_C = nullptr;
    size_t d_B_elems = 0, d_C_elems = 0;
    cudaStream_t stream = nullptr; // compute
    cudaStream_t stream_up = nullptr, stream_down = nullptr; // the two PCIe directions
    std::vector<cudaEvent_t> events;
    // B-staged row tiles (spmm_tiled.cu), optional
    int tl_T = 0, tl_BR = 0, tl_tiles = 0, tl_NS = 0, tl_POOL = 0, tl_max_recs = 0, tl_max_blob = 0, tl_chunk = 0, tl_kt = 0, tl_depth = 0, tl_drains = 0, tl_ksplit = 0;
    long long tl_box_rows_loaded = 0, tl_single_rows = 0; // B rows staged per pass over the matrix
    int auto_calls = 0;                                   // whole-matrix AUTO multiplies seen before the lazy build
    bool tl_tried = false;                                // AUTO already attempted the lazy build
    bool tl_auto = false;                                 // the layout was built by AUTO (it may rebuild it for another k)
    unsigned char *d_tblob = nullptr;
    void *d_tdesc = nullptr, *d_tloads = nullptr;
    int *d_tsingles = nullptr;
    int *d_torder = nullptr;          // walking order of the tiles (nullptr: as they lie)
    int tl_stride = 0, tl_group = 0;  // detected far-band distance in rows, planes interleaved per super-group
    // precomputed CTA cuts per (kind, grid size): kind 0 = rows of the CSR
    mutable std::map<long long, int *> bounds;
    // merge-path scratch (carry rows), grown on demand
    double *d_carry = nullptr;
    int *d_carry_row = nullptr;
    size_t carry_elems = 0, carry_rows = 0;
};

namespace spmm
{
// launchers implemented in the kernel translation units
// rows [row_begin,row_end), each clipped to the non-zero range [nnz_lo,nnz_hi); row c_row0 is stored at d_C[0]
int launch_rows(const spmm_csr_s *A, int row_begin, int row_end, long long nnz_lo, long long nnz_hi, int c_row0,
                const double *d_B, long long ldb, double *d_C, long long ldc, int kc, int derived,
                cudaStream_t stream, const struct ExtraDst *extra = nullptr); // derived: 0 CSR row kernel only, 1 best available, 6 tiled
void set_b_window(bool on); // the next launch_rows of this thread gets a window of B rows (no whole-B prefetch)
int launch_merge(spmm_csr_s *A, int row_begin, int row_end, long long nnz_lo, long long nnz_hi, int c_row0,
                 const double *d_B, long long ldb, double *d_C, long long ldc, int kc, cudaStream_t stream,
                 const struct ExtraDst *extra = nullptr);
bool tiled_shape_ok(const spmm_csr_s *A, const double *d_B, long long ldb, const double *d_C, long long ldc, int kc);
int launch_tiled(const spmm_csr_s *A, const double *d_B, long long ldb, double *d_C, long long ldc, int kc,
                 cudaStream_t stream, const struct ExtraDst *extra = nullptr);
void free_tiles(spmm_csr_s *A);
int build_tiles(spmm_csr_s *A, int rows_per_tile, int box_rows, int kt_want, int ksplit_want);
// what AUTO cuts the tile layout for: 64-byte window rows for k <= 8; chunks shared by 2 / 4 CTAs (one k-tile group each) from k = 32 / 64
inline int tiles_kt_for(int k) { return k <= 8 ? 8 : 16; }
inline int tiles_ksplit_for(int k) { return k >= 64 ? 4 : (k >= 32 ? 2 : 1); }
bool stream_shape_ok(const spmm_csr_s *A, const double *d_B, long long ldb, const double *d_C, long long ldc, int kc);
int launch_stream(const spmm_csr_s *A, const double *d_B, long long ldb, double *d_C, long long ldc, int kc,
                  cudaStream_t stream);
// CTA cuts cached on the handle; `fill` launches the kernel that computes grid+1 cuts into its argument
int cached_bounds(const spmm_csr_s *A, int kind, int grid, cudaStream_t stream, const int **out,
                  const std::function<void(int *)> &fill);
void drop_bounds(spmm_csr_s *A, int kind); // kind < 0: all
// handle plumbing shared by spmm_capi.cu and csr_build.cu
int make_handle(int device, int n_rows, int n_cols, long long nnz, spmm_csr_s **out);
int alloc_arrays(spmm_csr_s *A);
int build_schedule(spmm_csr_s *A, cudaStream_t stream);
} // namespace spmm
