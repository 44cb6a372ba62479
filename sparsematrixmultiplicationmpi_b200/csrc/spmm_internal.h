// spmm_internal.h — host-side structures shared by the translation units of libspmm_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "spmm_b200.h"

namespace spmm
{

void set_error(const std::string &msg);
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
    int rows_np = 0;        // concurrent non-zeros per team
    int rows_unroll = 0;    // B-row loads in flight per team step
    int rows_vec = 0;       // doubles per lane (1, 2, 4)
    int rows_ctas_per_sm = 0;
    int rows_threads = 0;   // 128 / 256 / 512
    int merge_items = 0;    // merge-path items per team
};
Tuning &tuning();

struct DeviceProps
{
    int sm_count = 0;
    long long l2_bytes = 0;
};
const DeviceProps &device_props(int device);

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
    // staging for the host-buffer entry point
    double *h_stage = nullptr; // pinned
    size_t h_stage_elems = 0;
    double *d_B = nullptr, *d_C = nullptr;
    size_t d_B_elems = 0, d_C_elems = 0;
    cudaStream_t stream = nullptr; // owned, for host-buffer calls
    // merge-path scratch (carry rows), grown on demand
    double *d_carry = nullptr;
    int *d_carry_row = nullptr;
    size_t carry_elems = 0, carry_rows = 0;
};

namespace spmm
{
// launchers implemented in the kernel translation units
int launch_rows(const spmm_csr_s *A, int row_begin, int row_end, int c_row0, const double *d_B, long long ldb,
                double *d_C, long long ldc, int kc, cudaStream_t stream);
int launch_merge(spmm_csr_s *A, long long nnz_begin, long long nnz_end, int c_row0, const double *d_B,
                 long long ldb, double *d_C, long long ldc, int kc, bool range_mode, cudaStream_t stream);
int build_schedule(spmm_csr_s *A, cudaStream_t stream);
} // namespace spmm
