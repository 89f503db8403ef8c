// Bulk-copy (cp.async.bulk) read bandwidth of persistent CTAs, one per SM, each streaming its own slice of a buffer
// through a ring of shared-memory stages — the load side of k_frame_merge without anything else.
//   tools/tma_probe [MB]   prints GB/s for a sweep of (copy bytes, copies per stage, stages, access pattern)
// Measurement aid: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// pattern 0: CTA c reads a contiguous slice.  pattern 1: frame-major like k_frame_merge — step f of CTA c reads
// bytes [f * G * chunk + c * chunk, + chunk).
__global__ void __launch_bounds__(128, 1)
k_probe(const char* __restrict__ src, long long total, int chunk, int pieces, int stages, int steps, int pattern, int readback,
        unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = (unsigned long long*)smem;          // [stages]
    volatile int* freed = (volatile int*)(smem + 256);              // consumer progress
    const uint32_t ring = smem_u32(smem + 512);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&full[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *freed = 0;
    }
    __syncthreads();
    const int G = gridDim.x, c = blockIdx.x;
    if (threadIdx.x == 0) {
        for (int f = 0; f < steps; ++f) {
            const int st = f % stages;
            while (f >= stages && *freed < f - stages + 1) {}
            const uint32_t bar = smem_u32(&full[st]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(chunk) : "memory");
            const long long off = pattern == 0 ? ((long long)c * steps + f) * chunk : ((long long)f * G + c) * chunk;
            const int pb = chunk / pieces;
            for (int q = 0; q < pieces; ++q)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(ring + (uint32_t)st * chunk + q * pb), "l"(src + off + (long long)q * pb), "r"(pb), "r"(bar) : "memory");
        }
    } else if (threadIdx.x >= 32 && threadIdx.x < 64) {
        unsigned long long acc = 0;
        for (int f = 0; f < steps; ++f) {
            const int st = f % stages;
            const uint32_t bar = smem_u32(&full[st]), parity = (f / stages) & 1;
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (readback) {
                const uint4* p = (const uint4*)(smem + 512 + (size_t)st * chunk);
                for (int v = threadIdx.x - 32; v < chunk / 16; v += 32) { uint4 x = p[v]; acc += x.x ^ x.y ^ x.z ^ x.w; }
            }
            __syncwarp();
            if (threadIdx.x == 32) *freed = f + 1;
        }
        if (acc == 0x1234567ull) sink[0] = acc;
    }
}

int main(int argc, char** argv) {
    const long long MB = argc > 1 ? atoll(argv[1]) : 512;
    const long long total = MB << 20;
    char* buf;
    unsigned long long* sink;
    cudaMalloc(&buf, total);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, total);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int chunks[] = {7168, 14336, 28672, 57344};
    for (int grid : {sms, 144})
    for (int pattern = 0; pattern < 2; ++pattern)
        for (int ci = 0; ci < 4; ++ci)
            for (int pieces : {1, 4})
                for (int stages : {2, 3, 4, 7}) {
                    const int chunk = chunks[ci];
                    if ((long long)stages * chunk + 512 > 227 * 1024) continue;
                    if (grid == 144 && (pattern == 0 || chunk != 28672)) continue;
                    const int steps = (int)(total / ((long long)grid * chunk));
                    const size_t smem = 512 + (size_t)stages * chunk;
                    float best = 1e9f;
                    for (int it = 0; it < 4; ++it) {
                        cudaEventRecord(e0);
                        k_probe<<<grid, 128, smem>>>(buf, total, chunk, pieces, stages, steps, pattern, 0, sink);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        if (it && ms < best) best = ms;
                    }
                    cudaError_t err = cudaGetLastError();
                    const double bytes = (double)steps * grid * chunk;
                    printf("grid %3d pattern %d chunk %5d x%d pieces stages %d: %7.1f us  %6.0f GB/s  (%.0f KB in flight / SM)%s\n", grid, pattern, chunk, pieces, stages,
                           best * 1e3, bytes / best / 1e6, stages * chunk / 1024.0, err ? cudaGetErrorString(err) : "");
                }
    return 0;
}
