// lds_patterns.cu — shared-memory load wavefronts per warp instruction for the broadcast patterns the SpMM kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu ; ./lds_patterns
// Each of 8 warps per CTA issues ITER x 16 independent loads of one pattern; cycles per warp instruction at saturation
// = wavefronts the data pipe spends on it.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2000;

template <int WIDTH> // bytes per lane: 4, 8, 16
__device__ __forceinline__ unsigned ld(unsigned addr)
{
    unsigned a, b = 0, c = 0, d = 0;
    if constexpr (WIDTH == 16)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
    else if constexpr (WIDTH == 8)
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
    else
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(addr));
    return a ^ b ^ c ^ d;
}

template <int WIDTH>
__global__ void probe(const int *lane_off, long long *cycles, double *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 32768 / 8; i += blockDim.x)
        reinterpret_cast<double *>(sm)[i] = 1.0;
    __syncthreads();
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)lane_off[threadIdx.x & 31];
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it)
    {
        const unsigned a = base + ((it & 7) << 11); // 2 KB steps: same banks, different rows
#pragma unroll
        for (int j = 0; j < 16; ++j)
            acc ^= ld<WIDTH>(a + j * 1024);
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0)
        cycles[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc;
}

template <int WIDTH>
void run(const char *name, const int *h_off, int warps)
{
    int *d_off;
    long long *d_cyc;
    double *d_sink;
    cudaMalloc(&d_off, 32 * sizeof(int));
    cudaMalloc(&d_cyc, 148 * 32 * sizeof(long long));
    cudaMalloc(&d_sink, 148 * 1024 * sizeof(double));
    cudaMemcpy(d_off, h_off, 32 * sizeof(int), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe<WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<WIDTH><<<148, warps * 32, 65536>>>(d_off, d_cyc, d_sink);
    cudaDeviceSynchronize();
    probe<WIDTH><<<148, warps * 32, 65536>>>(d_off, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148 * 32];
    cudaMemcpy(h, d_cyc, 148 * warps * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < 148 * warps; ++i)
        mx = h[i] > mx ? (double)h[i] : mx;
    // all warps of a CTA run concurrently: pipe cycles per warp instruction = max cycles / (ITER*16*warps)
    printf("%-58s width %2d warps %2d: %.2f cycles per warp instruction (%s)\n", name, WIDTH, warps,
           mx / (ITER * 16.0 * warps), cudaGetErrorString(e));
    cudaFree(d_off);
    cudaFree(d_cyc);
    cudaFree(d_sink);
}

int main()
{
    int off[32];
    for (int warps : {8, 16})
    {
        for (int l = 0; l < 32; ++l) off[l] = l * 16;
        run<16>("LDS.128 32 distinct addresses (512 B contiguous)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 2) * 16;
        run<16>("LDS.128 16 pairs of lanes (256 B contiguous)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 16;
        run<16>("LDS.128 8 teams of 4 lanes (128 B contiguous)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 8) * 16;
        run<16>("LDS.128 4 teams of 8 lanes (64 B contiguous)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 8) * 32;
        run<16>("LDS.128 4 teams of 8 lanes (32 B apart)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = 0;
        run<16>("LDS.128 one address", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 16 + ((l / 4) * 5 % 8) * 256;
        run<16>("LDS.128 8 teams of 4 lanes, distinct banks, different rows", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 256;
        run<16>("LDS.128 8 teams of 4 lanes, same banks (8-way conflict)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 256 + (l % 4) * 16 + ((l / 4) & 1) * 64;
        run<16>("LDS.128 team rows: 8 rows x 64 B, odd teams other half", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = l * 8;
        run<8>("LDS.64 32 distinct addresses (256 B contiguous)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 8;
        run<8>("LDS.64 8 teams of 4 lanes (64 B contiguous)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 8 + ((l / 4) * 3 % 8) * 128;
        run<8>("LDS.64 8 teams of 4 lanes, distinct banks, different rows", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 128;
        run<8>("LDS.64 8 teams of 4 lanes, same banks", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 8) * 8;
        run<8>("LDS.64 4 teams of 8 lanes (32 B contiguous)", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = 0;
        run<8>("LDS.64 one address", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = l * 4;
        run<4>("LDS.32 32 distinct addresses", off, warps);
        for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 4;
        run<4>("LDS.32 8 teams of 4 lanes", off, warps);
    }
    return 0;
}
