// spmm_w1.cu — instantiations with 8-byte B/C accesses (odd k, odd leading dimension or unaligned slabs).
#include "spmm_launch.cuh"
namespace spmm
{
int launch_rows_w1(const spmm_csr_s *A, int kl, int nv, int np, int u, const SpmmArgs &a, int tiles, int dev, cudaStream_t s)
{
    return launch_rows_shape<1>(A, kl, nv, np, u, a, tiles, dev, s);
}
int launch_merge_w1(int kl, int nv, int u, const SpmmArgs &a, int tiles, cudaStream_t s)
{
    return launch_merge_shape<1>(kl, nv, u, a, tiles, s);
}
} // namespace spmm
