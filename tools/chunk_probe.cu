// Probe for a chunked two-pass pipeline (not part of the product library): does a second pass over a chunk that a first
// pass has just read hit the L2, and how long does the whole pipeline take?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chunk_probe tools/chunk_probe.cu
// Shapes: 36 864 rows of 7 168 bytes (C2).  "S" = read every row of a chunk (similarity-like, warp per row, adjacent warps also
// read the neighbour row); "G" = read every row of the chunk again, write 60 % of them compacted (gather-like).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int VPR = 448;          // 16-byte vectors per row

__device__ __forceinline__ uint4 ldnc(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stcs(uint4* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// rows [r0, r1): warp per row, reads the row and the row before it (the neighbour warp's row: L1/L2 hit)
__global__ void __launch_bounds__(256) k_S(const uint4* __restrict__ in, int r0, int r1, unsigned* sink) {
    const int lane = threadIdx.x & 31, r = r0 + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= r1) return;
    const uint4* a = in + (size_t)(r > 0 ? r - 1 : r) * VPR;
    const uint4* b = in + (size_t)r * VPR;
    unsigned acc = 0;
    for (int v0 = lane; v0 < VPR; v0 += 128) {
        uint4 x[4], y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) { x[q] = ldnc(a + v0 + 32 * q); y[q] = ldnc(b + v0 + 32 * q); }
#pragma unroll
        for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) acc += (x[q].x ^ y[q].x) + (x[q].y ^ y[q].y) + (x[q].z ^ y[q].z) + (x[q].w ^ y[q].w);
    }
    if (acc == 0x12345678u) *sink = acc;
}

// rows [r0, r1): warp per row; 3 of 5 rows are copied to out (compacted), 1 of those also adds the next row
__global__ void __launch_bounds__(128, 8) k_G(const uint4* __restrict__ in, uint4* __restrict__ out, int r0, int r1) {
    const int lane = threadIdx.x & 31, r = r0 + blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= r1) return;
    const int m = r % 5;
    if (m >= 3) return;
    const int d = (r / 5) * 3 + m;
    const uint4* a = in + (size_t)r * VPR;
    uint4* o = out + (size_t)d * VPR;
    if (m == 0) {
        const uint4* b = in + (size_t)(r + 3) * VPR;
        for (int v0 = lane; v0 < VPR; v0 += 128) {
            uint4 x[4], y[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) { x[q] = ldnc(a + v0 + 32 * q); y[q] = ldnc(b + v0 + 32 * q); }
#pragma unroll
            for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) { x[q].x += y[q].x; x[q].y += y[q].y; x[q].z += y[q].z; x[q].w += y[q].w; stcs(o + v0 + 32 * q, x[q]); }
        }
    } else {
        for (int v0 = lane; v0 < VPR; v0 += 256) {
            uint4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) if (v0 + 32 * q < VPR) x[q] = ldnc(a + v0 + 32 * q);
#pragma unroll
            for (int q = 0; q < 8; ++q) if (v0 + 32 * q < VPR) stcs(o + v0 + 32 * q, x[q]);
        }
    }
}

int main(int argc, char** argv) {
    const int rows = 36864;
    uint4 *in, *out; unsigned* sink;
    CK(cudaMalloc(&in, (size_t)rows * VPR * 16 + 8 * VPR * 16)); CK(cudaMalloc(&out, (size_t)rows * VPR * 16)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, (size_t)rows * VPR * 16)); CK(cudaMemset(out, 0, (size_t)rows * VPR * 16));
    cudaStream_t A, B, C; CK(cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&C, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, ev[64], fork, join; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (auto& evc : ev) CK(cudaEventCreateWithFlags(&evc, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    auto S = [&](int r0, int r1, cudaStream_t st) { k_S<<<(r1 - r0 + 7) / 8, 256, 0, st>>>(in, r0, r1, sink); };
    auto G = [&](int r0, int r1, cudaStream_t st) { k_G<<<(r1 - r0 + 3) / 4, 128, 0, st>>>(in, out, r0, r1); };
    for (int mode = 0; mode < 8; ++mode) {
        const int chunk_frames[8] = {64, 64, 8, 8, 4, 12, 16, 8};
        const int cf = chunk_frames[mode], cr = cf * 576, n = (rows + cr - 1) / cr;
        const char* name = mode == 0 ? "two-pass, whole sequence" : mode == 1 ? "S only" : mode == 2 ? "chunked 8f, one stream" : mode == 7 ? "chunked 8f, G only after all S (cold)" : "chunked, two streams";
        float best = 1e9f;
        for (int it = 0; it < 12; ++it) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, A));
            if (mode == 0) { S(0, rows, A); G(0, rows, A); }
            else if (mode == 1) { S(0, rows, A); }
            else if (mode == 2) { for (int c = 0; c < n; ++c) { int r0 = c * cr, r1 = min(rows, r0 + cr); S(r0, r1, A); G(r0, r1, A); } }
            else if (mode == 7) { for (int c = 0; c < n; ++c) { int r0 = c * cr, r1 = min(rows, r0 + cr); S(r0, r1, A); } for (int c = 0; c < n; ++c) { int r0 = c * cr, r1 = min(rows, r0 + cr); G(r0, r1, A); } }
            else {
                CK(cudaEventRecord(fork, A)); CK(cudaStreamWaitEvent(B, fork, 0));
                for (int c = 0; c < n; ++c) {
                    int r0 = c * cr, r1 = min(rows, r0 + cr);
                    S(r0, r1, A); CK(cudaEventRecord(ev[c], A));
                    CK(cudaStreamWaitEvent(B, ev[c], 0)); G(r0, r1, B);
                }
                CK(cudaEventRecord(join, B)); CK(cudaStreamWaitEvent(A, join, 0));
            }
            CK(cudaEventRecord(e1, A));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 2 && ms < best) best = ms;
        }
        printf("mode %d  %-40s chunk %2d frames x %2d: %7.1f us\n", mode, name, cf, n, best * 1e3f);
    }
    return 0;
}
