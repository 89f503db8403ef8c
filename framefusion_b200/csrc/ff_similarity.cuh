// Adjacent-frame cosine similarity in by-patch order (main.py:216-238, 345-349), generic path:
// one warp per by-patch position j reads rows order[j-1] and order[j] with 16-byte vector loads
// (the predecessor row is an L2 hit: it was row j-1's "current" row a few thousand rows earlier),
// reduces the three row sums with warp shuffles and finishes the rounding chain of ff_common.cuh.
// Also emits the threshold flag sim >= T(similarity_lower_bound) (main.py:113) and its count.
#pragma once
#include "ff_common.cuh"

namespace ff {

template <int DT, bool VEC>
__global__ void __launch_bounds__(256)
k_similarity(const void* __restrict__ hidden, int H, int S, const int* __restrict__ order, const int* __restrict__ chain,
             const int64_t* __restrict__ counters_in, float thr, float* __restrict__ sim, uint8_t* __restrict__ flag,
             int64_t* counters) {
    pdl_enter();
    __shared__ int s_cnt[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int j = blockIdx.x * 8 + wid;
    // the links of position j are requested before N (the count they are checked against) has arrived: the arrays
    // hold S entries and the grid covers no more than S positions, so the reads are in bounds either way
    const int jc = min(j, S - 1), jp = max(jc - 1, 0);
    const int c_cur = chain[jc], c_prev = chain[jp], o_cur = order[jc], o_prev = order[jp];
    const int N = (int)counters_in[C_N];
    int hit = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[C_TICKET2] = 0;   // slot counter of the scan kernels' records
    if (j < N) {
        float s = -2.0f;                                   // IGNORE_TOKEN at chain heads (main.py:225-238)
        if (j > 0 && c_cur == c_prev) {
            const char* ra = (const char*)hidden + (int64_t)o_prev * H * sizeof(typename Num<DT>::store_t);
            const char* rb = (const char*)hidden + (int64_t)o_cur * H * sizeof(typename Num<DT>::store_t);
            float dot = 0.f, na = 0.f, nb = 0.f;
            if (VEC) {
                const int nvec = H / Num<DT>::EPV;
                int v = lane;
                for (; v + 96 < nvec; v += 128) {          // 4 x 2 independent 16-byte loads in flight per lane
                    uint4 a0 = ldg16(ra + (int64_t)v * 16), b0 = ldg16(rb + (int64_t)v * 16);
                    uint4 a1 = ldg16(ra + (int64_t)(v + 32) * 16), b1 = ldg16(rb + (int64_t)(v + 32) * 16);
                    uint4 a2 = ldg16(ra + (int64_t)(v + 64) * 16), b2 = ldg16(rb + (int64_t)(v + 64) * 16);
                    uint4 a3 = ldg16(ra + (int64_t)(v + 96) * 16), b3 = ldg16(rb + (int64_t)(v + 96) * 16);
                    acc_pair<DT>(a0, b0, dot, na, nb);
                    acc_pair<DT>(a1, b1, dot, na, nb);
                    acc_pair<DT>(a2, b2, dot, na, nb);
                    acc_pair<DT>(a3, b3, dot, na, nb);
                }
                for (; v < nvec; v += 32) {
                    uint4 a0 = ldg16(ra + (int64_t)v * 16), b0 = ldg16(rb + (int64_t)v * 16);
                    acc_pair<DT>(a0, b0, dot, na, nb);
                }
            } else {
                for (int e = lane; e < H; e += 32)
                    acc_pair_scalar<DT>(Num<DT>::load(ra, e), Num<DT>::load(rb, e), dot, na, nb);
            }
            dot = warp_sum(dot);
            na = warp_sum(na);
            nb = warp_sum(nb);
            s = finish_cosine<DT>(dot, na, nb);
        }
        hit = (s >= thr);                                  // NaN compares false
        if (lane == 0) {
            sim[j] = s;
            flag[j] = (uint8_t)hit;
        }
    }
    if (lane == 0) s_cnt[wid] = hit;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_cnt[w];
        if (t) atomicAdd((unsigned long long*)&counters[C_COUNT], (unsigned long long)t);
    }
}

}  // namespace ff
