// Shared device helpers for the FrameFusion token-reduction kernels (sm_100a).
//
// Numerics contract (SURVEY.md §8 a-N; every rounding below is one the reference performs because
// it materialises a tensor in the hidden dtype T — /root/reference/framefusion/main.py:345-349, 304-317):
//   prod_T(a,b)  = T(a*b)                    the elementwise product tensor of cosine_similarity
//   sums are float32 (ATen accumulates bf16/f16 reductions in float32), then rounded to T once
//   norm = T(sqrt(sum_f32 a*a)),  den = T(n1*n2),  sim = T(dot/den)
//   merge: acc = T(acc + member) one member at a time in chain order, then T(acc / T(L+1))
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/framefusion_b200.h"

namespace ff {

constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

// device-side counters kept at the head of the workspace (int64 slots)
enum Counter {
    C_N = 0,        // by-patch chain tokens
    C_NVIS = 1,     // patch_type != -1
    C_COUNT = 2,    // sim >= thr
    C_TICKET = 3,   // last-block-done ticket
    C_SKEEP = 4,
    C_BRANCH = 5,
    C_K = 6,
    C_ERR = 7,
    C_NMERGED = 8,
    C_NNEXT = 9,    // N of the links written for the next call
    C_TICKET2 = 10, // records handed out to the rows outside the chains (k_keep_scan / k_decide_scan)
    C_FIRSTINV = 11,// S - (first chain row), 0 if there is none (k_links_seq; order of the read-once kernel's S units)
    C_SPANS = 12,   // spans of consecutive chain rows in the sequence (k_links_seq, first call of a prefill)
    C_NONUNI = 13,  // != 0: inside a span the patch ids do not run 0, 1, .., P-1, 0, 1, .. (k_links_seq)
    C_SLOTS = 32
};

// ------------------------------------------------------------------------------------------------
// dtype traits.  Values travel as float32 that are exactly representable in T.
// ------------------------------------------------------------------------------------------------
template <int DT> struct Num;

template <> struct Num<FF_BF16> {
    typedef uint16_t store_t;
    static constexpr int EPV = 8;   // elements per 16-byte vector
    static __device__ __forceinline__ float load(const void* p, int64_t i) {
        return __uint_as_float(((uint32_t)((const uint16_t*)p)[i]) << 16);
    }
    static __device__ __forceinline__ float rnd(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
    static __device__ __forceinline__ void store(void* p, int64_t i, float x) {
        ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(x);
    }
    static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
        f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
        f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
        f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
    }
    static __device__ __forceinline__ uint32_t pack2(float a, float b) {   // rounds to T
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ uint4 pack(const float* f) {
        return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
    // T(a + b) on a 16-byte vector: the packed add rounds the exact sum once, which equals the float32 add followed
    // by the rounding to T (24 bits hold the sum of two 8-bit significands without a second rounding that matters)
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
        __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    static __device__ __forceinline__ uint4 add_vec(const uint4& a, const uint4& b) {
        return make_uint4(add2(a.x, b.x), add2(a.y, b.y), add2(a.z, b.z), add2(a.w, b.w));
    }
    // T(x * 2^-k): exact scaling (or one rounding into the subnormals), the same value x / 2^k rounds to
    static __device__ __forceinline__ uint4 scale_vec(const uint4& a, float s) {
        const __nv_bfloat162 m = __float2bfloat162_rn(s);
        uint4 o;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) po[i] = __hmul2(pa[i], m);
        return o;
    }
    // T(a*b) for two packed pairs, returned as two floats
    static __device__ __forceinline__ void prod2(uint32_t a, uint32_t b, float& lo, float& hi) {
        __nv_bfloat162 p = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        uint32_t w = *reinterpret_cast<uint32_t*>(&p);
        lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
    }
};

template <> struct Num<FF_F16> {
    typedef uint16_t store_t;
    static constexpr int EPV = 8;
    static __device__ __forceinline__ float load(const void* p, int64_t i) {
        return __half2float(((const __half*)p)[i]);
    }
    static __device__ __forceinline__ float rnd(float x) { return __half2float(__float2half_rn(x)); }
    static __device__ __forceinline__ void store(void* p, int64_t i, float x) { ((__half*)p)[i] = __float2half_rn(x); }
    static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    static __device__ __forceinline__ uint32_t pack2(float a, float b) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ uint4 pack(const float* f) {
        return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {       // see Num<FF_BF16>::add2 (11-bit significands)
        __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    static __device__ __forceinline__ uint4 add_vec(const uint4& a, const uint4& b) {
        return make_uint4(add2(a.x, b.x), add2(a.y, b.y), add2(a.z, b.z), add2(a.w, b.w));
    }
    static __device__ __forceinline__ uint4 scale_vec(const uint4& a, float s) {
        const __half2 m = __float2half2_rn(s);
        uint4 o;
        const __half2* pa = reinterpret_cast<const __half2*>(&a);
        __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) po[i] = __hmul2(pa[i], m);
        return o;
    }
    static __device__ __forceinline__ void prod2(uint32_t a, uint32_t b, float& lo, float& hi) {
        float2 fa = __half22float2(*reinterpret_cast<__half2*>(&a));
        float2 fb = __half22float2(*reinterpret_cast<__half2*>(&b));
        lo = rnd(fa.x * fb.x); hi = rnd(fa.y * fb.y);      // float product of two f16 is exact; one rounding
    }
};

template <> struct Num<FF_F32> {
    typedef float store_t;
    static constexpr int EPV = 4;
    static __device__ __forceinline__ float load(const void* p, int64_t i) { return ((const float*)p)[i]; }
    static __device__ __forceinline__ float rnd(float x) { return x; }
    static __device__ __forceinline__ void store(void* p, int64_t i, float x) { ((float*)p)[i] = x; }
    static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
    static __device__ __forceinline__ uint4 pack(const float* f) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    static __device__ __forceinline__ uint32_t addf(uint32_t a, uint32_t b) { return __float_as_uint(__fadd_rn(__uint_as_float(a), __uint_as_float(b))); }
    static __device__ __forceinline__ uint4 add_vec(const uint4& a, const uint4& b) {
        return make_uint4(addf(a.x, b.x), addf(a.y, b.y), addf(a.z, b.z), addf(a.w, b.w));
    }
    static __device__ __forceinline__ uint4 scale_vec(const uint4& a, float s) {
        return make_uint4(__float_as_uint(__uint_as_float(a.x) * s), __float_as_uint(__uint_as_float(a.y) * s),
                          __float_as_uint(__uint_as_float(a.z) * s), __float_as_uint(__uint_as_float(a.w) * s));
    }
};

// Accumulates the three row sums of one 16-byte vector pair: dot += T(a*b), na += a*a, nb += b*b.
template <int DT>
__device__ __forceinline__ void acc_pair(const uint4& va, const uint4& vb, float& dot, float& na, float& nb) {
    float a[Num<DT>::EPV], b[Num<DT>::EPV];
    Num<DT>::unpack(va, a);
    Num<DT>::unpack(vb, b);
    if (DT == FF_F32) {
#pragma unroll
        for (int e = 0; e < Num<DT>::EPV; ++e) {
            dot += __fmul_rn(a[e], b[e]);          // the product tensor is materialised before the sum
            na = fmaf(a[e], a[e], na);
            nb = fmaf(b[e], b[e], nb);
        }
    } else {
#pragma unroll
        for (int e = 0; e < Num<DT>::EPV; ++e) {
            dot += Num<DT>::rnd(a[e] * b[e]);      // a*b is exact in float32 for 16-bit inputs: one rounding to T
            na = fmaf(a[e], a[e], na);             // exact product, float32 accumulate
            nb = fmaf(b[e], b[e], nb);
        }
    }
}

// Row sums of one 16-byte vector pair with two interleaved float32 accumulators per sum (even / odd elements):
// dot += T(a*b), na += a*a, nb += b*b.  bf16: the product tensor element T(a*b) is one mul.rn.bf16x2 (exact product,
// one rounding — what the reference's bf16 multiply does); the sums run on the packed float32x2 pipe of sm_100.
template <int DT>
__device__ __forceinline__ void acc_pair2(const uint4& va, const uint4& vb, float2& dot, float2& na, float2& nb) {
    if (DT == FF_BF16) {
        const uint32_t aw[4] = {va.x, va.y, va.z, va.w}, bw[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __nv_bfloat162 pa = *reinterpret_cast<const __nv_bfloat162*>(&aw[q]);
            __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&bw[q]);
            __nv_bfloat162 pp = __hmul2(pa, pb);
            const uint32_t pw = *reinterpret_cast<uint32_t*>(&pp);
            dot = __fadd2_rn(dot, make_float2(__uint_as_float(pw << 16), __uint_as_float(pw & 0xffff0000u)));
            const float2 af = make_float2(__uint_as_float(aw[q] << 16), __uint_as_float(aw[q] & 0xffff0000u));
            const float2 bf = make_float2(__uint_as_float(bw[q] << 16), __uint_as_float(bw[q] & 0xffff0000u));
            na = __ffma2_rn(af, af, na);
            nb = __ffma2_rn(bf, bf, nb);
        }
    } else {
        float a[Num<DT>::EPV], b[Num<DT>::EPV];
        Num<DT>::unpack(va, a);
        Num<DT>::unpack(vb, b);
#pragma unroll
        for (int e = 0; e < Num<DT>::EPV; e += 2) {
            if (DT == FF_F32) { dot.x += __fmul_rn(a[e], b[e]); dot.y += __fmul_rn(a[e + 1], b[e + 1]); }
            else { dot.x += Num<DT>::rnd(a[e] * b[e]); dot.y += Num<DT>::rnd(a[e + 1] * b[e + 1]); }
            na.x = fmaf(a[e], a[e], na.x);
            na.y = fmaf(a[e + 1], a[e + 1], na.y);
            nb.x = fmaf(b[e], b[e], nb.x);
            nb.y = fmaf(b[e + 1], b[e + 1], nb.y);
        }
    }
}

template <int DT>
__device__ __forceinline__ void acc_pair_scalar(float a, float b, float& dot, float& na, float& nb) {
    if (DT == FF_F32) dot += __fmul_rn(a, b); else dot += Num<DT>::rnd(a * b);
    na = fmaf(a, a, na);
    nb = fmaf(b, b, nb);
}

// sim = T( T(dot) / T( T(sqrt(na)) * T(sqrt(nb)) ) )
template <int DT>
__device__ __forceinline__ float finish_cosine(float dot, float na, float nb) {
    float d = Num<DT>::rnd(dot);
    float n1 = Num<DT>::rnd(sqrtf(na));
    float n2 = Num<DT>::rnd(sqrtf(nb));
    float den = Num<DT>::rnd(n1 * n2);
    return Num<DT>::rnd(d / den);
}

// T(x / div) elementwise, div = T(L + 1) (main.py:314-317: one true division, rounded to T).
//  * a power-of-two run (2, 4, 8 ... rows: most runs) is an exact scaling, done on the packed vector;
//  * bf16, any run up to 256 rows: T(x * RN(1/div)) equals T(x / div) for EVERY finite bf16 x (checked exhaustively,
//    65 280 values x 256 divisors; a bf16 quotient by a small integer is never within float32 error of a rounding
//    boundary of T), so the reciprocal multiply is exact and nothing needs the IEEE division;
//  * everything else (f16 / f32 with other divisors — f16 does have exceptions in its subnormals —, longer runs) divides.
__device__ __noinline__ float ieee_div(float x, float d) { return x / d; }

template <int DT>
struct Divider {
    float div, rcp;
    bool pow2, by_rcp;
    __device__ __forceinline__ explicit Divider(int n) {
        div = Num<DT>::rnd((float)n);
        rcp = 1.0f / div;
        pow2 = (n & (n - 1)) == 0 && n <= 256;             // 2^-k is then a normal number in every T
        by_rcp = pow2 || (DT == FF_BF16 && n <= 256);
    }
    __device__ __forceinline__ float one(float x) const { return by_rcp ? x * rcp : ieee_div(x, div); }
    // by_rcp only: x * rcp in float32, rounded to T
    __device__ __forceinline__ uint4 vec_rcp(const uint4& a) const {
        float x[Num<DT>::EPV];
        Num<DT>::unpack(a, x);
#pragma unroll
        for (int e = 0; e < Num<DT>::EPV; ++e) x[e] *= rcp;
        return Num<DT>::pack(x);
    }
    __device__ __forceinline__ uint4 vec_fast(const uint4& a) const {
        if (pow2) return Num<DT>::scale_vec(a, rcp);
        if (by_rcp) return vec_rcp(a);
        return vec(a);
    }
    __device__ __noinline__ uint4 vec(const uint4 a) const {      // out of line: rare
        float x[Num<DT>::EPV];
        Num<DT>::unpack(a, x);
#pragma unroll
        for (int e = 0; e < Num<DT>::EPV; ++e) x[e] = one(x[e]);
        return Num<DT>::pack(x);
    }
};

// Programmatic dependent launch (ff_api.cu: launch_pdl).  pdl_wait() returns once the previous kernel of the stream
// has completed and its writes are visible; pdl_trigger() lets the next kernel's blocks become resident early (they
// stop at their own pdl_wait()).  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_wait();
    pdl_trigger();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// streaming 16-byte load / store (read once / written once: keep them out of L1)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total gets the block sum.  `smem` needs 33 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                       // protect smem reuse across calls
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? smem[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += t;
        }
        smem[lane] = wi - w;               // exclusive warp offsets
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    int excl = smem[wid] + incl - v;
    *total = smem[32];
    return excl;
}

// order-preserving key of a float: larger float -> larger key; every NaN -> the largest key
// (torch.topk ranks NaN above everything).
__device__ __forceinline__ uint32_t float_key(float x) {
    if (x != x) return 0xffffffffu;
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int DT>
__device__ __forceinline__ float load_T(const void* p, int64_t i) { return Num<DT>::load(p, i); }

// ---- shared by the single-launch kernel (ff_frame.cuh) and the gather ---------------------------------------------------
// the aux tensors, one entry per (tensor, plane): rows of at most 512 bytes in 16- or 8-byte pieces, one piece per lane
struct AuxFlat {
    int n;                                         // entries; -1: the tensors do not fit this form (gather_aux_rows instead)
    int row_bytes[8];
    int piece[8];                                  // 16 or 8: bytes per lane
    const char* src[8];
    char* dst[8];
};

// relaxed device-scope accesses to the flag words the CTAs of a grid exchange
// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// ---- row movers: 16-byte vectors; a row is cut into pieces of N vectors per lane (N = MAXN ... 1) and a last partial
// piece, so that every piece issues its N (or 2 N) loads back to back with no predicate in between — predicated loads
// are not batched by ptxas, and a warp with one or two loads in flight is latency bound.
template <int N>
__device__ __forceinline__ void copy_piece(const char* __restrict__ src, char* __restrict__ dst) {
    uint4 x[N];
#pragma unroll
    for (int q = 0; q < N; ++q) x[q] = ld_stream16(src + q * 512);
#pragma unroll
    for (int q = 0; q < N; ++q) st_stream16(dst + q * 512, x[q]);
}

// src row -> dst row (read for the last time, written once: streaming both ways)
__device__ __forceinline__ void copy_row(const char* __restrict__ src, char* __restrict__ dst, int nvec, int lane) {
    int v = 0;                                              // vectors done (warp-uniform)
    src += lane * 16;
    dst += lane * 16;
#pragma unroll 1
    for (; v + 256 <= nvec; v += 256) copy_piece<8>(src + (int64_t)v * 16, dst + (int64_t)v * 16);
    if (v + 128 <= nvec) { copy_piece<4>(src + (int64_t)v * 16, dst + (int64_t)v * 16); v += 128; }
    if (v + 64 <= nvec) { copy_piece<2>(src + (int64_t)v * 16, dst + (int64_t)v * 16); v += 64; }
    if (v + 32 <= nvec) { copy_piece<1>(src + (int64_t)v * 16, dst + (int64_t)v * 16); v += 32; }
    if (v + lane < nvec) copy_piece<1>(src + (int64_t)v * 16, dst + (int64_t)v * 16);
}

}  // namespace ff
