// Probe (not part of the product library): read pattern of the similarity pass.  Warp per row; each warp reads its row
// and a "previous" row at distance D rows; rows are visited in sequence order (stride 1) or chain-major (by-patch) order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sim_probe tools/sim_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int VPR = 448, P = 576, F = 64;

template <int NC> __device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 v;
    if (NC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else v = __ldg(p);
    return v;
}

// ORDER 0: warp w -> row w (sequence order), previous row = w - D.   ORDER 1: warp w -> chain-major: p = w / F, t = w % F,
// row = t * P + p, previous = row - P (what k_similarity does on the first call of a prefill)
template <int NC, int ORDER>
__global__ void __launch_bounds__(256) k_S(const uint4* __restrict__ in, int rows, int D, unsigned* sink) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= rows) return;
    int r, pr;
    if (ORDER == 0) { r = w; pr = w - D; }
    else { const int p = w / F, t = w % F; r = t * P + p; pr = r - P; }
    if (pr < 0) pr = r;
    const uint4* a = in + (size_t)pr * VPR;
    const uint4* b = in + (size_t)r * VPR;
    unsigned acc = 0;
    for (int v0 = lane; v0 < VPR; v0 += 128) {
        uint4 x[4], y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) { x[q] = ld<NC>(a + v0 + 32 * q); y[q] = ld<NC>(b + v0 + 32 * q); }
#pragma unroll
        for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) acc += (x[q].x ^ y[q].x) + (x[q].y ^ y[q].y) + (x[q].z ^ y[q].z) + (x[q].w ^ y[q].w);
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void k_flush(uint4* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) p[i] = make_uint4(1, 2, 3, 4);
}

int main() {
    const int rows = P * F;
    uint4 *in, *junk; unsigned* sink;
    const size_t junk_n = (size_t)256 << 20 >> 4;
    CK(cudaMalloc(&in, (size_t)rows * VPR * 16)); CK(cudaMalloc(&junk, junk_n * 16)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, (size_t)rows * VPR * 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct V { const char* name; int nc, order, D; } vs[] = {
        {"sequence order, prev = row-1,   no_allocate", 1, 0, 1}, {"sequence order, prev = row-1,   __ldg", 0, 0, 1},
        {"sequence order, prev = row-576, no_allocate", 1, 0, 576}, {"sequence order, prev = row-576, __ldg", 0, 0, 576},
        {"chain-major (by-patch) order,   no_allocate", 1, 1, 0}, {"chain-major (by-patch) order,   __ldg", 0, 1, 0}};
    for (auto& v : vs) {
        float best = 1e9f;
        for (int it = 0; it < 8; ++it) {
            k_flush<<<148 * 8, 256>>>(junk, junk_n);
            CK(cudaEventRecord(e0));
            const int g = (rows + 7) / 8;
            if (v.nc && v.order == 0) k_S<1, 0><<<g, 256>>>(in, rows, v.D, sink);
            else if (!v.nc && v.order == 0) k_S<0, 0><<<g, 256>>>(in, rows, v.D, sink);
            else if (v.nc) k_S<1, 1><<<g, 256>>>(in, rows, v.D, sink);
            else k_S<0, 1><<<g, 256>>>(in, rows, v.D, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 2 && ms < best) best = ms;
        }
        printf("%-48s %7.1f us  (%.0f GB/s of one pass)\n", v.name, best * 1e3f, rows * 7168e-3 / (best * 1e3f));
    }
    return 0;
}
