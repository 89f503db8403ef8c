// Probe (not part of the product library): what does the ORDER in which a record-driven row gather walks its rows cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu
// Uniform video 64 frames x 576 patches x 7 168-byte rows, 39 % of the rows merged into the previous kept row of their
// chain (runs), records (src row, dst row, by-patch position, run length) like k_keep_scan writes them.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int VPR = 448, P = 576, F = 64;

__device__ __forceinline__ uint4 ldnc(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stcs(uint4* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// member rows by arithmetic (MEM=0) or through an order[] array like the real kernel (MEM=1)
template <int MEM>
__global__ void __launch_bounds__(128, 8) k_gather(const uint4* __restrict__ in, uint4* __restrict__ out, const int4* __restrict__ rec,
                                                   const int* __restrict__ order, int n) {
    const int lane = threadIdx.x & 31, u = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (u >= n) return;
    const int4 r = __ldg(rec + u);
    const uint4* a = in + (size_t)r.x * VPR;
    uint4* o = out + (size_t)r.y * VPR;
    const int L = r.w;
    if (L == 0) {
        for (int v0 = lane; v0 < VPR; v0 += 256) {
            uint4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) if (v0 + 32 * q < VPR) x[q] = ldnc(a + v0 + 32 * q);
#pragma unroll
            for (int q = 0; q < 8; ++q) if (v0 + 32 * q < VPR) stcs(o + v0 + 32 * q, x[q]);
        }
    } else {
        for (int v0 = lane; v0 < VPR; v0 += 128) {
            uint4 x[4], y[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) x[q] = ldnc(a + v0 + 32 * q);
            for (int m = 1; m <= L; ++m) {
                const int row = MEM ? order[r.z + m] : r.x + m * P;
                const uint4* b = in + (size_t)row * VPR;
#pragma unroll
                for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) y[q] = ldnc(b + v0 + 32 * q);
#pragma unroll
                for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) { x[q].x += y[q].x; x[q].y += y[q].y; x[q].z += y[q].z; x[q].w += y[q].w; }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) if (v0 + 32 * q < VPR) stcs(o + v0 + 32 * q, x[q]);
        }
    }
}

__global__ void k_flush(uint4* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) p[i] = make_uint4(1, 2, 3, 4);
}

int main() {
    const int rows = P * F;
    std::vector<char> flag(rows, 0);
    srand(1);
    for (int p = 0; p < P; ++p) for (int t = 1; t < F; ++t) flag[t * P + p] = (rand() % 100) < 39;
    std::vector<int> dst(rows, -1), order(rows);
    int nk = 0;
    for (int i = 0; i < rows; ++i) if (!flag[i]) dst[i] = nk++;
    for (int p = 0; p < P; ++p) for (int t = 0; t < F; ++t) order[p * F + t] = t * P + p;
    std::vector<int4> rec_bp;            // by-patch ascending
    for (int p = 0; p < P; ++p) for (int t = 0; t < F; ++t) {
        const int i = t * P + p;
        if (flag[i]) continue;
        int L = 0; while (t + 1 + L < F && flag[(t + 1 + L) * P + p]) ++L;
        rec_bp.push_back(make_int4(i, dst[i], p * F + t, L));
    }
    std::vector<int4> rec_bpd(rec_bp.rbegin(), rec_bp.rend());                 // by-patch descending (what the library does)
    std::vector<int4> rec_seq(rec_bp);                                           // destination rows ascending
    std::sort(rec_seq.begin(), rec_seq.end(), [](const int4& a, const int4& b) { return a.y < b.y; });
    std::vector<int4> rec_seqd(rec_seq.rbegin(), rec_seq.rend());               // destination rows descending
    printf("rows %d kept %d (%.1f MB read, %.1f MB written)\n", rows, nk, rows * 7168e-6, nk * 7168e-6);
    uint4 *in, *out, *junk; int4* d_rec; int* d_order;
    const size_t junk_n = (size_t)256 << 20 >> 4;
    CK(cudaMalloc(&in, (size_t)rows * VPR * 16)); CK(cudaMalloc(&out, (size_t)rows * VPR * 16)); CK(cudaMalloc(&junk, junk_n * 16));
    CK(cudaMalloc(&d_rec, rec_bp.size() * 16)); CK(cudaMalloc(&d_order, rows * 4));
    CK(cudaMemset(in, 1, (size_t)rows * VPR * 16));
    CK(cudaMemcpy(d_order, order.data(), rows * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct V { const char* name; std::vector<int4>* r; int mem; } vs[] = {
        {"by-patch descending, members via order[]", &rec_bpd, 1}, {"by-patch descending, members by arithmetic", &rec_bpd, 0},
        {"by-patch ascending, members via order[]", &rec_bp, 1}, {"destination ascending, members via order[]", &rec_seq, 1},
        {"destination ascending, members by arithmetic", &rec_seq, 0}, {"destination descending, members via order[]", &rec_seqd, 1}};
    for (auto& v : vs) {
        CK(cudaMemcpy(d_rec, v.r->data(), v.r->size() * 16, cudaMemcpyHostToDevice));
        float best = 1e9f;
        for (int it = 0; it < 8; ++it) {
            k_flush<<<148 * 8, 256>>>(junk, junk_n);                              // cold L2
            CK(cudaEventRecord(e0));
            const int n = (int)v.r->size();
            if (v.mem) k_gather<1><<<(n + 3) / 4, 128>>>(in, out, d_rec, d_order, n);
            else k_gather<0><<<(n + 3) / 4, 128>>>(in, out, d_rec, d_order, n);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 2 && ms < best) best = ms;
        }
        printf("%-48s %7.1f us  (%.0f GB/s read+write)\n", v.name, best * 1e3f, (rows + nk) * 7168e-3 / (best * 1e3f));
    }
    return 0;
}
