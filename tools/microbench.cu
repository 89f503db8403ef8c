// Standalone bandwidth probes for the B200 box (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints: HBM read, HBM copy, L2-resident read bandwidth for a few grid shapes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ILP>
__global__ void k_read(const uint4* __restrict__ p, size_t n, unsigned* sink) {
    unsigned acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (ILP - 1) * stride < n; i += ILP * stride) {
        uint4 v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p + i + k * stride));
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    for (; i < n; i += stride) { uint4 v = p[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

template <int ILP>
__global__ void k_copy(const uint4* __restrict__ p, uint4* __restrict__ q, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (ILP - 1) * stride < n; i += ILP * stride) {
        uint4 v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p + i + k * stride));
#pragma unroll
        for (int k = 0; k < ILP; ++k) q[i + k * stride] = v[k];
    }
    for (; i < n; i += stride) q[i] = p[i];
}

// warp-per-row pattern: each warp reads whole 7168-byte rows (the kernels' access shape)
__global__ void k_rows(const uint4* __restrict__ p, int rows, int vec_per_row, unsigned* sink) {
    unsigned acc = 0;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        const uint4* row = p + (size_t)r * vec_per_row;
        uint4 v[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) v[k] = (k * 32 + lane < vec_per_row) ? __ldg(row + k * 32 + lane) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 14; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
float time_ms(F f, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s SMs %d L2 %d MB clock %d MHz\n", pr.name, pr.multiProcessorCount, pr.l2CacheSize >> 20, pr.clockRate / 1000);
    const size_t big = (size_t)1 << 30;       // 1 GiB
    uint4 *a, *b; unsigned* sink;
    CK(cudaMalloc(&a, big)); CK(cudaMalloc(&b, big)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, big)); CK(cudaMemset(b, 2, big));
    const int sms = pr.multiProcessorCount;
    for (int bpsm : {2, 4, 8}) for (int threads : {256, 512}) {
        int grid = sms * bpsm;
        float ms = time_ms([&] { k_read<8><<<grid, threads>>>(a, big / 16, sink); }, 10);
        printf("hbm_read  1GiB grid=%4d x %3d ILP8: %7.1f GB/s\n", grid, threads, big / ms / 1e6);
    }
    for (int bpsm : {2, 4, 8}) {
        int grid = sms * bpsm;
        float ms = time_ms([&] { k_copy<8><<<grid, 256>>>(a, b, big / 16); }, 10);
        printf("hbm_copy  1GiB->1GiB grid=%4d x 256 ILP8: %7.1f GB/s (read+write)\n", grid, 2.0 * big / ms / 1e6);
    }
    {
        float ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, big, cudaMemcpyDeviceToDevice)); }, 10);
        printf("cudaMemcpy D2D 1GiB: %7.1f GB/s (read+write)\n", 2.0 * big / ms / 1e6);
    }
    for (size_t mb : {16, 32, 64, 96}) {
        size_t bytes = mb << 20;
        for (int bpsm : {4, 8}) {
            int grid = sms * bpsm;
            float ms = time_ms([&] { k_read<8><<<grid, 512>>>(a, bytes / 16, sink); }, 50);
            printf("l2_read  %3zu MiB grid=%4d x 512 ILP8: %7.1f GB/s\n", mb, grid, bytes / ms / 1e6);
        }
    }
    {
        const int rows = 36898, vpr = 448;   // 64x576 rows of 3584 bf16
        for (int bpsm : {2, 4, 8}) {
            int grid = sms * bpsm;
            float ms = time_ms([&] { k_rows<<<grid, 256>>>(a, rows, vpr, sink); }, 20);
            printf("row_read 36898 x 7168B grid=%4d x 256: %7.1f GB/s\n", grid, (double)rows * vpr * 16 / ms / 1e6);
        }
        float ms = time_ms([&] { k_rows<<<(rows + 7) / 8, 256>>>(a, rows, vpr, sink); }, 20);
        printf("row_read 36898 x 7168B grid=rows/8 x 256 (one row per warp): %7.1f GB/s\n", (double)rows * vpr * 16 / ms / 1e6);
    }
    // launch latency of an empty dependent kernel chain
    {
        float ms = time_ms([&] { for (int i = 0; i < 10; ++i) k_read<1><<<1, 32>>>(a, 0, sink); }, 20);
        printf("10 dependent tiny launches: %.2f us each\n", ms * 100);
    }
    return 0;
}
