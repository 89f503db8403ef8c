// Importance signal: attention probabilities of the last `num` queries against all keys
// (/root/reference/framefusion/utils.py:27-57).  The reference materialises repeat_kv(K) (7x the KV bytes)
// and runs a [num x D] x [D x S] matmul per query head; here one thread owns one (kv head, key position),
// reads that 256-byte K row once and serves every query head of the GQA group from shared memory.
//   logits = T( T( T(q.k) * scale ) + bias ),  bias = -inf above the causal diagonal (triu(diagonal=S-L+1))
//   probs  = T( exp(x - max) / sum exp(x - max) )   with float32 softmax internals
#pragma once
#include <cooperative_groups.h>

#include "ff_common.cuh"

namespace ff {

constexpr int IMP_THREADS = 128;

template <int DT, bool VEC>
__global__ void __launch_bounds__(IMP_THREADS)
k_importance_logits(const void* __restrict__ q, const void* __restrict__ k, int n_q_heads, int n_kv_heads, int S, int D,
                    int num, int64_t q_hs, int64_t q_ss, int64_t k_hs, int64_t k_ss, int is_causal, float scale,
                    float* __restrict__ logits /* [Hq, num, S] float32 holding T values */) {
    pdl_enter();
    extern __shared__ __align__(16) float s_q[];          // [group * num][D]
    typedef typename Num<DT>::store_t st;
    const int hk = blockIdx.y;
    const int group = n_q_heads / n_kv_heads;
    const int L = num;
    for (int idx = threadIdx.x; idx < group * L * D; idx += blockDim.x) {
        const int d = idx % D, r = (idx / D) % L, g = idx / (D * L);
        const int h = hk * group + g;
        s_q[idx] = Num<DT>::load(q, (int64_t)h * q_hs + (int64_t)(S - L + r) * q_ss + d);
    }
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const st* krow = (const st*)k + (int64_t)hk * k_hs + (int64_t)s * k_ss;
    const int GL = group * L;
    if (VEC) {
        // eight (query head, query row) dot products advance together over one pass of the K row (16-byte loads); more
        // pairs than eight (Qwen2-VL: 7 heads x 4 query rows per KV head) take further passes, which the L1 serves
        for (int gb = 0; gb < GL; gb += 8) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int v = 0; v < D / Num<DT>::EPV; ++v) {
                float kf[Num<DT>::EPV];
                Num<DT>::unpack(ldg16((const char*)krow + (int64_t)v * 16), kf);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (gb + g < GL) {
                        // 16-byte shared-memory reads: D is a multiple of the vector length on this path
                        const float4* qv = reinterpret_cast<const float4*>(s_q + (gb + g) * D + v * Num<DT>::EPV);
#pragma unroll
                        for (int f = 0; f < Num<DT>::EPV / 4; ++f) {
                            const float4 qq = qv[f];
                            acc[g] = fmaf(kf[4 * f + 0], qq.x, acc[g]);
                            acc[g] = fmaf(kf[4 * f + 1], qq.y, acc[g]);
                            acc[g] = fmaf(kf[4 * f + 2], qq.z, acc[g]);
                            acc[g] = fmaf(kf[4 * f + 3], qq.w, acc[g]);
                        }
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                if (gb + g < GL) {
                    const int r = (gb + g) % L, h = hk * group + (gb + g) / L;
                    float x = Num<DT>::rnd(acc[g]);            // matmul output in T
                    x = Num<DT>::rnd(x * scale);               // * scale_factor
                    const float bias = (is_causal && s > S - L + r) ? -INFINITY : 0.f;
                    x = Num<DT>::rnd(x + bias);                // += attn_bias
                    logits[((int64_t)h * L + r) * S + s] = x;
                }
            }
        }
        return;
    }
    for (int gr = 0; gr < GL; ++gr) {                      // rows that are not 16-byte aligned: element by element
        const float* qv = s_q + gr * D;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc = fmaf(Num<DT>::load(krow, d), qv[d], acc);
        const int r = gr % L, h = hk * group + gr / L;
        float x = Num<DT>::rnd(acc);                       // matmul output in T
        x = Num<DT>::rnd(x * scale);                       // * scale_factor
        const float bias = (is_causal && s > S - L + r) ? -INFINITY : 0.f;
        x = Num<DT>::rnd(x + bias);                        // += attn_bias
        logits[((int64_t)h * L + r) * S + s] = x;
    }
}

// One thread-block CLUSTER of four CTAs per (head, query) row: each CTA owns a quarter of the row, keeps it in registers
// between the passes when it fits, and the four partial maxima / sums meet through distributed shared memory (two
// cluster barriers instead of a second kernel or 28 lonely blocks on 148 SMs).  The partial sums are added in rank order,
// so the result does not depend on timing.
constexpr int SOFTMAX_CLUSTER = 4;
constexpr int SOFTMAX_CACHE = 10;                         // values per thread: rows up to 4 * 10 * 1024 = 40 960 keys

template <int DT>
__global__ void __cluster_dims__(SOFTMAX_CLUSTER, 1, 1) __launch_bounds__(1024)
k_softmax_rows(const float* __restrict__ logits, int S, void* __restrict__ probs) {
    pdl_enter();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float s_red[32];
    __shared__ float s_part[2];                            // this CTA's maximum and sum, read by the other three
    const int rank = (int)cluster.block_rank();
    const int row = blockIdx.x / SOFTMAX_CLUSTER;
    const float* x = logits + (int64_t)row * S;
    const int seg = (S + SOFTMAX_CLUSTER - 1) / SOFTMAX_CLUSTER;
    const int lo = min(rank * seg, S), hi = min(lo + seg, S);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool cached = seg <= SOFTMAX_CACHE * (int)blockDim.x;
    float c[SOFTMAX_CACHE];
    float m = -INFINITY;
    if (cached) {
#pragma unroll
        for (int i = 0; i < SOFTMAX_CACHE; ++i) {
            const int s = lo + threadIdx.x + i * blockDim.x;
            c[i] = s < hi ? __ldg(x + s) : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < SOFTMAX_CACHE; ++i) m = fmaxf(m, c[i]);
    } else {
        for (int s = lo + threadIdx.x; s < hi; s += blockDim.x) m = fmaxf(m, x[s]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if (lane == 0) s_red[wid] = m;
    __syncthreads();
    if (wid == 0) {
        float v = lane < nw ? s_red[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
        if (lane == 0) s_part[0] = v;
    }
    cluster.sync();
    m = -INFINITY;
#pragma unroll
    for (int r = 0; r < SOFTMAX_CLUSTER; ++r) m = fmaxf(m, *cluster.map_shared_rank(&s_part[0], r));
    float sum = 0.f;
    if (cached) {
#pragma unroll
        for (int i = 0; i < SOFTMAX_CACHE; ++i) {
            const int s = lo + threadIdx.x + i * blockDim.x;
            if (s < hi) {
                c[i] = expf(c[i] - m);
                sum += c[i];
            }
        }
    } else {
        for (int s = lo + threadIdx.x; s < hi; s += blockDim.x) sum += expf(x[s] - m);
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[wid] = sum;
    __syncthreads();
    if (wid == 0) {
        float v = lane < nw ? s_red[lane] : 0.f;
        v = warp_sum(v);
        if (lane == 0) s_part[1] = v;
    }
    cluster.sync();
    sum = 0.f;
#pragma unroll
    for (int r = 0; r < SOFTMAX_CLUSTER; ++r) sum += *cluster.map_shared_rank(&s_part[1], r);
    cluster.sync();                                        // nobody leaves while its shared memory may still be read
    if (cached) {
#pragma unroll
        for (int i = 0; i < SOFTMAX_CACHE; ++i) {
            const int s = lo + threadIdx.x + i * blockDim.x;
            if (s < hi) Num<DT>::store(probs, (int64_t)row * S + s, c[i] / sum);
        }
    } else {
        for (int s = lo + threadIdx.x; s < hi; s += blockDim.x)
            Num<DT>::store(probs, (int64_t)row * S + s, expf(x[s] - m) / sum);
    }
}

}  // namespace ff
