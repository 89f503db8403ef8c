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

constexpr int IMP_THREADS = 256;
constexpr int IMP_LANES = 4;                              // lanes that share one key row
constexpr int IMP_KEYS = IMP_THREADS / IMP_LANES;         // key rows per block
constexpr int IMP_PAD = 8;                                // floats between the lanes' slices of q in shared memory (bank spread)

// 16-byte vector of T -> float2 pairs (even element, odd element)
template <int DT>
__device__ __forceinline__ void unpack_pairs(const uint4& v, float2* p) {
    float f[Num<DT>::EPV];
    Num<DT>::unpack(v, f);
#pragma unroll
    for (int e = 0; e < Num<DT>::EPV / 2; ++e) p[e] = make_float2(f[2 * e], f[2 * e + 1]);
}

// VEC (rows of 16-byte vectors, head_dim a multiple of IMP_LANES vectors): four lanes share a key row — each reads a
// contiguous quarter of it (a warp reads eight whole rows: 2 KB in one piece) and carries the partial dot products of every
// (query head, query row) pair of the GQA group over its quarter, as packed float32 pairs; two shuffles add the quarters up.
// One thread per key with the whole row (the first version) was a chain of ~3 000 dependent instructions at half a wave.
template <int DT, bool VEC>
__global__ void __launch_bounds__(IMP_THREADS)
k_importance_logits(const void* __restrict__ q, const void* __restrict__ k, int n_q_heads, int n_kv_heads, int S, int D,
                    int num, int64_t q_hs, int64_t q_ss, int64_t k_hs, int64_t k_ss, int is_causal, float scale,
                    float* __restrict__ logits /* [Hq, num, S] float32 holding T values */) {
    pdl_enter();
    extern __shared__ __align__(16) float s_q[];
    typedef typename Num<DT>::store_t st;
    const int hk = blockIdx.y;
    const int group = n_q_heads / n_kv_heads;
    const int L = num;
    const int GL = group * L;
    if (VEC) {
        // q in shared memory as [lane slice][pair][D / 4] with IMP_PAD floats between the slices
        const int DQ = D / IMP_LANES, slice = GL * DQ + IMP_PAD;
        for (int idx = threadIdx.x; idx < GL * D; idx += blockDim.x) {
            const int d = idx % D, gr = idx / D;
            const int r = gr % L, g = gr / L;
            const int h = hk * group + g;
            s_q[(d / DQ) * slice + gr * DQ + (d % DQ)] = Num<DT>::load(q, (int64_t)h * q_hs + (int64_t)(S - L + r) * q_ss + d);
        }
        __syncthreads();
        const int sub = threadIdx.x & (IMP_LANES - 1);
        const int s = blockIdx.x * IMP_KEYS + (threadIdx.x >> 2);
        const int sc = min(s, S - 1);                       // every lane of a warp stays for the shuffles
        const char* kq = (const char*)((const st*)k + (int64_t)hk * k_hs + (int64_t)sc * k_ss) + (int64_t)sub * DQ * sizeof(st);
        const float* qs = s_q + sub * slice;
        const int nv = DQ / Num<DT>::EPV;                   // 16-byte vectors of this lane's quarter
        for (int gb = 0; gb < GL; gb += 8) {
            float2 acc[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) acc[g] = make_float2(0.f, 0.f);
            for (int v = 0; v < nv; ++v) {
                float2 kp[Num<DT>::EPV / 2];
                unpack_pairs<DT>(ldg16(kq + (int64_t)v * 16), kp);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (gb + g < GL) {
                        const float2* qv = reinterpret_cast<const float2*>(qs + (gb + g) * DQ + v * Num<DT>::EPV);
#pragma unroll
                        for (int e = 0; e < Num<DT>::EPV / 2; ++e) acc[g] = __ffma2_rn(kp[e], qv[e], acc[g]);
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float x = acc[g].x + acc[g].y;
                x += __shfl_xor_sync(FULL, x, 1);
                x += __shfl_xor_sync(FULL, x, 2);
                // lane `sub` of the four writes the pairs gb + g == sub (mod 4)
                if (gb + g < GL && s < S && ((gb + g) & (IMP_LANES - 1)) == sub) {
                    const int r = (gb + g) % L, h = hk * group + (gb + g) / L;
                    x = Num<DT>::rnd(x);                    // matmul output in T
                    x = Num<DT>::rnd(x * scale);            // * scale_factor
                    const float bias = (is_causal && s > S - L + r) ? -INFINITY : 0.f;
                    x = Num<DT>::rnd(x + bias);             // += attn_bias
                    logits[((int64_t)h * L + r) * S + s] = x;
                }
            }
        }
        return;
    }
    for (int idx = threadIdx.x; idx < GL * D; idx += blockDim.x) {
        const int d = idx % D, r = (idx / D) % L, g = idx / (D * L);
        const int h = hk * group + g;
        s_q[idx] = Num<DT>::load(q, (int64_t)h * q_hs + (int64_t)(S - L + r) * q_ss + d);
    }
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const st* krow = (const st*)k + (int64_t)hk * k_hs + (int64_t)s * k_ss;
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

// The same logits on the tensor cores (bf16 / f16, head_dim a multiple of 32 up to IMP_MMA_MAX_D): per KV head the logits are
// K [S x D] times the group's queries [D x group*num] — "a bf16 matmul with float32 accumulation, rounded once to T", which
// is what the reference's matmul is.  A warp owns 16 keys: every lane fetches its two rows' pieces with 16-byte loads, all in
// flight at once (the whole 4 KB tile is one round trip), and feeds them to mma.sync m16n8k16 as they lie: the k index of a
// dot product may be permuted freely as long as both operands agree, so the eight consecutive elements a lane holds serve as
// the fragment columns {2c, 2c+1, 2c+8, 2c+9} of two k-steps, and the queries are read from shared memory in the same order.
// Replaces a kernel whose duration was the dependent load / FMA chain of one warp (25 us at S = 22 290 for 23 MB).
constexpr int IMP_MMA_KEYS = 16 * (IMP_THREADS / 32);     // keys per block
constexpr int IMP_MMA_MAX_D = 256;
constexpr int IMP_MMA_PAD = 8;                            // elements between query rows in shared memory (bank spread)

template <int DT>
__device__ __forceinline__ void mma_16x8x16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    if (DT == FF_BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int DT, int NJ>                                 // NJ = D / 32
__global__ void __launch_bounds__(IMP_THREADS)
k_importance_logits_mma(const void* __restrict__ q, const void* __restrict__ k, int n_q_heads, int n_kv_heads, int S, int num,
                        int64_t q_hs, int64_t q_ss, int64_t k_hs, int64_t k_ss, int is_causal, float scale,
                        float* __restrict__ logits /* [Hq, num, S] float32 holding T values */) {
    pdl_enter();
    extern __shared__ __align__(16) uint16_t s_qh[];       // [GLpad][D + IMP_MMA_PAD] raw 16-bit elements
    constexpr int D = NJ * 32, ROW = D + IMP_MMA_PAD;
    const int hk = blockIdx.y, group = n_q_heads / n_kv_heads, L = num, GL = group * L, GLpad = (GL + 7) / 8 * 8;
    // the K tile first: its 2 * NJ loads per lane travel while the queries are staged
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const int key0 = (blockIdx.x * (IMP_THREADS / 32) + (threadIdx.x >> 5)) * 16;
    const int s0 = key0 + g, s1 = s0 + 8;
    const char* r0 = (const char*)((const uint16_t*)k + (int64_t)hk * k_hs + (int64_t)min(s0, S - 1) * k_ss) + c * 16;
    const char* r1 = (const char*)((const uint16_t*)k + (int64_t)hk * k_hs + (int64_t)min(s1, S - 1) * k_ss) + c * 16;
    uint4 lo[NJ], hi[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        lo[j] = ldg16(r0 + j * 64);
        hi[j] = ldg16(r1 + j * 64);
    }
    const uint16_t* q16 = (const uint16_t*)q;
    for (int idx = threadIdx.x; idx < GLpad * D; idx += blockDim.x) {
        const int d = idx % D, gr = idx / D;
        uint16_t v = 0;
        if (gr < GL) v = q16[(int64_t)(hk * group + gr / L) * q_hs + (int64_t)(S - L + gr % L) * q_ss + d];
        s_qh[gr * ROW + d] = v;
    }
    __syncthreads();
    if (key0 >= S) return;
#pragma unroll 1
    for (int nt = 0; nt < GLpad; nt += 8) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const uint16_t* qrow = s_qh + (nt + g) * ROW + c * 8;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const uint4 b = *reinterpret_cast<const uint4*>(qrow + j * 32);
            mma_16x8x16<DT>(acc, lo[j].x, hi[j].x, lo[j].y, hi[j].y, b.x, b.y);
            mma_16x8x16<DT>(acc, lo[j].z, hi[j].z, lo[j].w, hi[j].w, b.z, b.w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int n = nt + 2 * c + (e & 1), s = (e & 2) ? s1 : s0;
            if (n < GL && s < S) {
                const int r = n % L, h = hk * group + n / L;
                float x = Num<DT>::rnd(acc[e]);             // matmul output in T
                x = Num<DT>::rnd(x * scale);                // * scale_factor
                const float bias = (is_causal && s > S - L + r) ? -INFINITY : 0.f;
                x = Num<DT>::rnd(x + bias);                 // += attn_bias
                logits[((int64_t)h * L + r) * S + s] = x;
            }
        }
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
