// Frame-pipelined merge kernel: similarity + threshold select + run merge + compaction of hidden_states and of the aux
// tensors in ONE launch in which every row of hidden_states travels HBM -> shared memory -> HBM exactly once
// (main.py:104-138, threshold branch, first merge call of a uniform video: F frames x P patches in one span of the
// sequence, chains = patch ids — the layout every embed-stage patch of the reference builds, interface.py:140-166).
//
// Why this shape.  A token is compared with the token of the SAME patch one frame earlier (main.py:216-238) and merged
// rows are summed along that chain (main.py:285-317), while the output is compacted in SEQUENCE order (main.py:132-138):
// the destination of a row is the number of kept rows before it, i.e. it needs the flags of every patch of its frame.
// So the chains are OWNED: CTA c holds patches [cR, cR + R) (R = ceil(P / #SMs): 4 at P = 576) and walks the frames in
// order, one persistent CTA per SM, all co-resident.  Per frame a CTA moves R contiguous rows (28 KB at C2):
//
//   producer thread   one bulk copy (cp.async.bulk, mbarrier complete_tx) of the CTA's R rows of frame f into a ring of
//                     shared-memory stages, as far ahead as the ring allows — this is what keeps HBM busy;
//   S warps (one per chain)   cosine of row (f, p) with row (f-1, p) — both in shared memory, the norm of the previous
//                     row carried in a register — with the reference's rounding chain (main.py:345-349); the flag goes
//                     to shared memory and, with ONE relaxed 64-bit reduction, to the global flag words
//                     ((reported mask << 32) | kept mask per 32 sequence rows);
//   prefix warp       polls the flag words of frame f until every patch has reported, popcounts: destination rows of
//                     the CTA's own rows of that frame, and the running base of the next frame;
//   G warps (one per chain)   a kept row whose successor is kept goes out as it is (bulk copy shared -> global); a kept
//                     row whose successor merges opens a run: members are added into the chain's accumulator row in
//                     shared memory, one rounding to T per add in chain order, and the closing add divides by T(L+1)
//                     (main.py:304-317) and sends the row out.  Nothing is read twice, not even from the L2;
//   aux warps (one per chain)   cos / sin / patch_type / position-id rows of the kept rows, dst[], and — after one grid
//                     barrier at the very end — the by-patch arrays of the next call (order / chain / rank), so that the
//                     multi-kernel path can serve the following calls.  They also move the rows outside the chains.
//
// Nothing waits for a CTA that is behind by less than the ring; the only grid-wide coupling is the flag words, polled a
// couple of frames behind the similarity front.  Every wait is bounded (FF_ST_INTERNAL).  The kernel speculates on the
// threshold branch (main.py:114-116) and verifies on the device that the layout is the uniform one; otherwise it reports
// FF_ST_ERROR = 3 and the host redoes the call on the multi-kernel path (the input is never modified).
#pragma once
#include "ff_common.cuh"
#include "ff_fused.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FR_MAXR = 8;                         // chains per CTA
constexpr int FR_MAXSTAGES = 16;                   // ring stages (frames in shared memory)
constexpr int FR_NQ = 32;                          // frames the flag / destination rings in shared memory hold (> stages + 2)
constexpr int FR_META = 2048;                      // bytes of shared memory in front of the stages
constexpr long long FR_TIMEOUT_NS = 400ll * 1000 * 1000;   // a launch older than this gives up at its next wait
constexpr int FR_TRACE_K = 8;                      // time stamps per (CTA, frame) of a traced launch

struct FrameArgs {
    AuxFlat auxf;
    const char* hidden;
    char* out;
    int S, P, R, nvec, row_bytes, n_stages;
    unsigned* gbar;                                // [2] zero on entry: arrivals at the final barrier, abort flag
    unsigned long long* words;                     // [S / 32 + 1] zero on entry: (reported mask << 32) | kept mask of rows 32 i ..
    float* sim;                                    // [N] by by-patch position
    uint8_t* flag;                                 // [N] merged-away flags by by-patch position
    int* dst;                                      // [S] destination row or -1
    int* keptdst;                                  // [N] scratch: destination by by-patch position
    int* len_next;                                 // [P] scratch: kept rows per chain
    int* order_next;
    int* chain_next;
    int* rank_next;
    int64_t* counters;
    int64_t* counters_next;
    int64_t* status;
    long long* trace;                              // [grid][trace_frames][FR_TRACE_K] globaltimer stamps, or null
    int trace_frames;
    float thr;
    double bound;
    long long seq;
};

struct FrameShared {
    unsigned long long full[FR_MAXSTAGES];         // mbarriers: stage loaded
    volatile int s_done[FR_MAXR];                  // frames whose similarity S warp w has finished
    volatile int g_free[FR_MAXR];                  // G warp w no longer needs the stages of frames below this
    volatile int a_done[FR_MAXR];                  // frames aux warp w has finished
    volatile int p_done;                           // frames whose destinations are in dstv[]
    volatile int abort;
    volatile int base_total;                       // kept rows before the first row behind the span
    int pad;
    volatile int dstv[FR_NQ][FR_MAXR];
    volatile unsigned char kept[FR_NQ][FR_MAXR];
};
static_assert(sizeof(FrameShared) <= FR_META, "FR_META too small");

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fr_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void fr_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool fr_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fr_tma_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fr_tma_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fr_tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fr_tma_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fr_tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fr_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 fr_lds16(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void fr_sts16(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ long long fr_time() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct FrameCtx {
    FrameShared* sh;
    unsigned* gbar;
    long long t_end;                                        // %globaltimer value after which waits give up
    __device__ __forceinline__ void give_up() const {
        sh->abort = 1;
        st_relaxed32(gbar + 1, 1u);
    }
    // called from inside a wait loop with its poll count: true once the launch is too old (checked every 256 polls)
    __device__ __forceinline__ bool expired(int spins) const {
        if ((spins & 255) != 255 || fr_time() < t_end) return false;
        give_up();
        return true;
    }
    // whole warp: until *ctr >= need; false once the kernel is giving up
    __device__ __forceinline__ bool wait_ge(const volatile int* ctr, int need) const {
        int spins = 0;
        while (*ctr < need) {
            if (sh->abort || expired(++spins)) return false;
            __nanosleep(20);
        }
        return true;
    }
};

// dot += T(p * c), nb += c * c over one 16-byte vector pair, two interleaved float32 accumulators each
template <int DT>
__device__ __forceinline__ void fr_dot_nb(const uint4& vp, const uint4& vc, float2& dot, float2& nb) {
    if (DT == FF_BF16) {
        const uint32_t pw[4] = {vp.x, vp.y, vp.z, vp.w}, cw[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __nv_bfloat162 pa = *reinterpret_cast<const __nv_bfloat162*>(&pw[q]);
            __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&cw[q]);
            __nv_bfloat162 pp = __hmul2(pa, pb);
            const uint32_t w = *reinterpret_cast<uint32_t*>(&pp);
            dot = __fadd2_rn(dot, make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)));
            const float2 cf = make_float2(__uint_as_float(cw[q] << 16), __uint_as_float(cw[q] & 0xffff0000u));
            nb = __ffma2_rn(cf, cf, nb);
        }
    } else {
        float p[Num<DT>::EPV], c[Num<DT>::EPV];
        Num<DT>::unpack(vp, p);
        Num<DT>::unpack(vc, c);
#pragma unroll
        for (int e = 0; e < Num<DT>::EPV; e += 2) {
            if (DT == FF_F32) { dot.x += __fmul_rn(p[e], c[e]); dot.y += __fmul_rn(p[e + 1], c[e + 1]); }
            else { dot.x += Num<DT>::rnd(p[e] * c[e]); dot.y += Num<DT>::rnd(p[e + 1] * c[e + 1]); }
            nb.x = fmaf(c[e], c[e], nb.x);
            nb.y = fmaf(c[e + 1], c[e + 1], nb.y);
        }
    }
}

template <int DT>
__device__ __forceinline__ void fr_nb(const uint4& vc, float2& nb) {
    float c[Num<DT>::EPV];
    Num<DT>::unpack(vc, c);
#pragma unroll
    for (int e = 0; e < Num<DT>::EPV; e += 2) {
        nb.x = fmaf(c[e], c[e], nb.x);
        nb.y = fmaf(c[e + 1], c[e + 1], nb.y);
    }
}

// the aux rows of sequence row r -> destination row d (one piece per lane and entry when the tensors have that form)
__device__ __forceinline__ void frame_aux(const AuxFlat& f, const AuxPack& aux, int r, int d, int lane) {
    if (f.n < 0) { gather_aux_rows(aux, r, d, lane); return; }
    uint4 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        v[e] = make_uint4(0, 0, 0, 0);
        if (e < f.n) {
            const int rb = f.row_bytes[e];
            const char* s = f.src[e] + (int64_t)r * rb;
            if (f.piece[e] == 8) { if (lane * 8 < rb) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(s) + lane); v[e].x = t.x; v[e].y = t.y; } }
            else if (lane * 16 < rb) v[e] = __ldg(reinterpret_cast<const uint4*>(s) + lane);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (e < f.n) {
            const int rb = f.row_bytes[e];
            char* o = f.dst[e] + (int64_t)d * rb;
            if (f.piece[e] == 8) { if (lane * 8 < rb) reinterpret_cast<uint2*>(o)[lane] = make_uint2(v[e].x, v[e].y); }
            else if (lane * 16 < rb) reinterpret_cast<uint4*>(o)[lane] = v[e];
        }
}

// the call is decided: sizes, the speculated branch, the counters of the next call (main.py:112-120)
__device__ __forceinline__ void frame_finish(const FrameArgs& a, long long N, long long n_vis, long long n_merged, int ec, int internal) {
    const long long s_keep = a.S - n_merged;
    a.counters[C_COUNT] = n_merged;
    a.counters[C_NNEXT] = N - n_merged;
    a.counters[C_SKEEP] = s_keep;
    a.counters[C_BRANCH] = 0;
    a.counters[C_K] = 0;
    a.counters[C_NMERGED] = n_merged;
    a.counters_next[C_N] = N - n_merged;
    a.counters_next[C_NVIS] = n_vis - n_merged;
    a.counters_next[C_COUNT] = 0;
    a.counters_next[C_TICKET] = 0;
    a.counters_next[C_TICKET2] = 0;
    a.status[FF_ST_SEQ_KEEP] = s_keep;
    a.status[FF_ST_COUNT] = n_merged;
    a.status[FF_ST_NVIS] = n_vis;
    a.status[FF_ST_NCHAIN] = N;
    a.status[FF_ST_BRANCH] = 0;
    a.status[FF_ST_TOPK] = 0;
    a.status[FF_ST_ERROR] = ec;
    a.status[FF_ST_NMERGED] = n_merged;
    a.status[FF_ST_FUSED] = 2;
    a.status[FF_ST_INTERNAL] = internal;
    __threadfence_system();
    *(volatile int64_t*)&a.status[FF_ST_SEQ] = a.seq;
}

#define FR_STAMP(f, k) do { if (a.trace && (f) < a.trace_frames) a.trace[((int64_t)blockIdx.x * a.trace_frames + (f)) * FR_TRACE_K + (k)] = fr_time(); } while (0)

template <int DT, int MAXR>
__global__ void __launch_bounds__(32 * (2 + 3 * MAXR), 1)
k_frame_merge(const __grid_constant__ FrameArgs a, const __grid_constant__ AuxPack aux) {
    extern __shared__ __align__(128) unsigned char fr_smem[];
    pdl_wait();
    FrameShared* sh = reinterpret_cast<FrameShared*>(fr_smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int R = a.R, P = a.P, S = a.S;

    // ---- is this the layout the kernel is built for?  (every CTA reads the same counters and decides alike)
    const long long N = a.counters[C_N], n_vis = a.counters[C_NVIS], inv = a.counters[C_FIRSTINV];
    const bool uniform = N > 0 && N == n_vis && a.counters[C_SPANS] == 1 && a.counters[C_NONUNI] == 0 && N % P == 0 &&
                         inv > 0 && (long long)S - inv + N <= S;
    if (!uniform) {
        if (blockIdx.x == 0 && threadIdx.x == 0) frame_finish(a, N, n_vis, 0, 3, 0);
        return;
    }
    const int F = (int)(N / P), first = (int)(S - inv);
    const int p0 = blockIdx.x * R;
    const int Rc = min(R, P - p0);                          // chains of this CTA (>= 1 by the grid size)
    const int NS = a.n_stages;
    const int64_t rb = a.row_bytes;
    const uint32_t stage_bytes = (uint32_t)(R * rb);
    const uint32_t stages0 = fr_smem_u32(fr_smem + FR_META);
    const uint32_t acc0 = stages0 + (uint32_t)NS * stage_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) fr_mbar_init(fr_smem_u32(&sh->full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sh->p_done = 0;
        sh->abort = 0;
        sh->base_total = 0;
    }
    if (threadIdx.x < FR_MAXR) {
        sh->s_done[threadIdx.x] = 0;
        sh->g_free[threadIdx.x] = 0;
        sh->a_done[threadIdx.x] = 0;
    }
    __syncthreads();
    FrameCtx cx;
    cx.sh = sh;
    cx.gbar = a.gbar;
    cx.t_end = fr_time() + FR_TIMEOUT_NS;

    const int w_s = wid - 2, w_g = wid - 2 - R, w_a = wid - 2 - 2 * R;    // index inside the role

    if (wid == 0) {
        // ================================ producer ================================
        if (lane == 0) {
            for (int f = 0; f < F; ++f) {
                const int st = f % NS;
                if (f >= NS) {
                    bool ok = true;
                    for (int w = 0; w < Rc && ok; ++w) ok = cx.wait_ge(&sh->g_free[w], f - NS + 1) && cx.wait_ge(&sh->a_done[w], f - NS + 1);
                    if (!ok) break;
                }
                const uint32_t bar = fr_smem_u32(&sh->full[st]);
                const uint32_t bytes = (uint32_t)(Rc * rb);
                fr_mbar_expect_tx(bar, bytes);
                fr_tma_load(stages0 + (uint32_t)st * stage_bytes, a.hidden + ((int64_t)first + (int64_t)f * P + p0) * rb, bytes, bar);
                FR_STAMP(f, 0);
            }
        }
    } else if (wid == 1) {
        // ================================ prefix ================================
        int base = first;                                   // rows in front of the span are all kept
        for (int f = 0; f < F; ++f) {
            // nobody has all the flags of frame f before this CTA's own S warps are through it
            bool ok = true;
            for (int w = 0; w < Rc && ok; ++w) ok = cx.wait_ge(&sh->s_done[w], f + 1);
            if (!ok) break;
            const int lo = first + f * P, hi = lo + P;     // rows of the frame
            const int w_lo = lo >> 5, w_hi = (hi - 1) >> 5, nW = w_hi - w_lo + 1;
            int running = base;
            for (int c0 = 0; c0 < nW && ok; c0 += 32) {
                const int wi = w_lo + c0 + lane;
                const bool have = c0 + lane < nW;
                unsigned need = 0u;
                if (have) {
                    need = 0xffffffffu;
                    if (wi == w_lo) need &= 0xffffffffu << (lo & 31);
                    if (wi == w_hi) need &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                }
                unsigned long long v = 0ull;
                int spins = 0;
                while (true) {
                    if (have) v = ld_relaxed64(a.words + wi);
                    const bool done = ((unsigned)(v >> 32) & need) == need;
                    if (__all_sync(FULL, done)) break;
                    ++spins;
                    if (sh->abort || ((spins & 63) == 0 && ld_relaxed32(a.gbar + 1) != 0u) || cx.expired(spins)) { ok = false; break; }
                    __nanosleep(64);
                }
                if (!ok) break;
                const unsigned km = (unsigned)v & need;
                const int k = __popc(km);
                int incl = k;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                for (int w = 0; w < Rc; ++w) {
                    const int r = lo + p0 + w;
                    if (have && (r >> 5) == wi) sh->dstv[f % FR_NQ][w] = running + incl - k + __popc(km & ((1u << (r & 31)) - 1u));
                }
                running += __shfl_sync(FULL, incl, 31);
            }
            if (!ok) { sh->abort = 1; break; }
            base = running;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                if (f == F - 1) sh->base_total = base;
                __threadfence_block();
                sh->p_done = f + 1;
                FR_STAMP(f, 2);
            }
        }
    } else if (w_s < R) {
        // ================================ similarity, chain p0 + w_s ================================
        if (w_s < Rc) {
            const int w = w_s, p = p0 + w;
            float na_prev = 0.f;
            for (int f = 0; f < F; ++f) {
                const int st = f % NS;
                const uint32_t bar = fr_smem_u32(&sh->full[st]);
                const uint32_t parity = (uint32_t)((f / NS) & 1);
                int spins = 0;
                bool ok = true;
                while (!fr_mbar_try_wait(bar, parity)) {
                    if (sh->abort || cx.expired(++spins)) { ok = false; break; }
                }
                if (!ok) break;
                const uint32_t cur = stages0 + (uint32_t)st * stage_bytes + (uint32_t)(w * rb) + lane * 16;
                const uint32_t prv = stages0 + (uint32_t)((f + NS - 1) % NS) * stage_bytes + (uint32_t)(w * rb) + lane * 16;
                float2 d0 = make_float2(0.f, 0.f), d1 = d0, b0 = d0, b1 = d0;
                const int nvec = a.nvec;
                if (f > 0) {
#pragma unroll 2
                    for (int v = lane; v < nvec; v += 64) {
                        const uint4 pa = fr_lds16(prv + (v - lane) * 16), ca = fr_lds16(cur + (v - lane) * 16);
                        uint4 pb = make_uint4(0, 0, 0, 0), cb = pb;
                        const bool two = v + 32 < nvec;
                        if (two) { pb = fr_lds16(prv + (v - lane + 32) * 16); cb = fr_lds16(cur + (v - lane + 32) * 16); }
                        fr_dot_nb<DT>(pa, ca, d0, b0);
                        if (two) fr_dot_nb<DT>(pb, cb, d1, b1);
                    }
                } else {
                    for (int v = lane; v < nvec; v += 64) {
                        fr_nb<DT>(fr_lds16(cur + (v - lane) * 16), b0);
                        if (v + 32 < nvec) fr_nb<DT>(fr_lds16(cur + (v - lane + 32) * 16), b1);
                    }
                }
                const float dot = warp_sum((d0.x + d0.y) + (d1.x + d1.y));
                const float nb = warp_sum((b0.x + b0.y) + (b1.x + b1.y));
                float s = -2.0f;                            // IGNORE_TOKEN at chain heads (main.py:225-238)
                if (f > 0) s = finish_cosine<DT>(dot, na_prev, nb);
                na_prev = nb;
                const unsigned kept = !(f > 0 && s >= a.thr);   // NaN compares false: kept
                if (lane == 0) {
                    const int r = first + f * P + p;
                    const int64_t j = (int64_t)p * F + f;
                    sh->kept[f % FR_NQ][w] = (unsigned char)kept;
                    red_add64(a.words + (r >> 5), (1ull << (32 + (r & 31))) | ((unsigned long long)kept << (r & 31)));
                    a.sim[j] = s;
                    a.flag[j] = (uint8_t)(kept ^ 1u);
                    __threadfence_block();
                    sh->s_done[w] = f + 1;
                    if (w == 0) FR_STAMP(f, 1);
                }
                __syncwarp();
            }
        }
    } else if (w_g < R) {
        // ================================ merge + rows out, chain p0 + w_g ================================
        if (w_g < Rc) {
            const int w = w_g;
            const uint32_t accp = acc0 + (uint32_t)(w * rb);
            const int nvec = a.nvec;
            int L = 0, anchor_d = -1;
            for (int f = 0; f < F; ++f) {
                if (!cx.wait_ge(&sh->s_done[w], min(f + 2, F)) || !cx.wait_ge(&sh->p_done, f + 1)) break;
                __threadfence_block();
                const bool kept = sh->kept[f % FR_NQ][w] != 0;
                const bool nxt_kept = f + 1 < F ? sh->kept[(f + 1) % FR_NQ][w] != 0 : true;
                const uint32_t cur = stages0 + (uint32_t)(f % NS) * stage_bytes + (uint32_t)(w * rb);
                if (kept) {
                    const int d = sh->dstv[f % FR_NQ][w];
                    if (nxt_kept) {
                        if (lane == 0) fr_tma_store(a.out + (int64_t)d * rb, cur, (uint32_t)rb);
                    } else {
                        anchor_d = d;
                        L = 0;
                    }
                } else {
                    const uint32_t prv = stages0 + (uint32_t)((f + NS - 1) % NS) * stage_bytes + (uint32_t)(w * rb);
                    const uint32_t src = L == 0 ? prv : accp;     // the anchor itself, or the running sum
                    ++L;
                    if (nxt_kept) {
                        const Divider<DT> dv(L + 1);
#pragma unroll 2
                        for (int v = lane; v < nvec; v += 32) {
                            const uint4 y = fr_lds16(src + v * 16), x = fr_lds16(cur + v * 16);
                            fr_sts16(accp + v * 16, dv.vec_fast(Num<DT>::add_vec(y, x)));    // T(T(acc + member) / T(L + 1))
                        }
                        fr_fence_async();
                        __syncwarp();
                        if (lane == 0) fr_tma_store(a.out + (int64_t)anchor_d * rb, accp, (uint32_t)rb);
                        L = 0;
                    } else {
#pragma unroll 2
                        for (int v = lane; v < nvec; v += 32) {
                            const uint4 y = fr_lds16(src + v * 16), x = fr_lds16(cur + v * 16);
                            fr_sts16(accp + v * 16, Num<DT>::add_vec(y, x));                  // T(acc + member), main.py:304
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    fr_tma_commit();
                    fr_tma_wait_read_1();                   // the stores of frame f - 1 have left shared memory
                    sh->g_free[w] = f;
                    if (w == 0) FR_STAMP(f, 3);
                }
                __syncwarp();
            }
            if (lane == 0) fr_tma_wait_all();
        }
    } else if (w_a < R) {
        // ================================ aux rows, dst[], rows outside the chains ================================
        const int w = w_a;
        const int unit = blockIdx.x * R + w, n_units = gridDim.x * R;
        const int nvec = a.nvec;
        for (int t = unit; t < first; t += n_units) {       // rows in front of the span keep their place
            copy_row(a.hidden + (int64_t)t * rb, a.out + (int64_t)t * rb, nvec, lane);
            if (aux.n) frame_aux(a.auxf, aux, t, t, lane);
            if (lane == 0) { a.dst[t] = t; a.rank_next[t] = -1; }
        }
        bool ok = true;
        if (w < Rc) {
            const int p = p0 + w;
            int nk = 0;
            for (int f = 0; f < F; ++f) {
                if (!cx.wait_ge(&sh->p_done, f + 1)) { ok = false; break; }
                __threadfence_block();
                const bool kept = sh->kept[f % FR_NQ][w] != 0;
                const int d = kept ? sh->dstv[f % FR_NQ][w] : -1;
                const int r = first + f * P + p;
                if (kept && aux.n) frame_aux(a.auxf, aux, r, d, lane);
                if (lane == 0) {
                    a.dst[r] = d;
                    a.keptdst[(int64_t)p * F + f] = d;
                    sh->a_done[w] = f + 1;
                    if (w == 0) FR_STAMP(f, 4);
                }
                nk += kept;
                __syncwarp();
            }
            if (lane == 0) a.len_next[p] = nk;
        }
        if (ok && cx.wait_ge(&sh->p_done, F)) {             // rows behind the span move up by the merged rows
            __threadfence_block();
            const int bt = sh->base_total, n_post = S - first - (int)N;
            for (int t = unit; t < n_post; t += n_units) {
                const int r = first + (int)N + t, d = bt + t;
                copy_row(a.hidden + (int64_t)r * rb, a.out + (int64_t)d * rb, nvec, lane);
                if (aux.n) frame_aux(a.auxf, aux, r, d, lane);
                if (lane == 0) { a.dst[r] = d; a.rank_next[d] = -1; }
            }
        }
    }

    // ---- every chain is through: one barrier over the grid, then the by-patch arrays of the next call
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.gbar, 1u);
        int spins = 0;
        while (ld_relaxed32(a.gbar) < gridDim.x) {
            if (ld_relaxed32(a.gbar + 1) != 0u || cx.expired(++spins)) break;
            __nanosleep(100);
        }
        if (ld_relaxed32(a.gbar + 1) != 0u) sh->abort = 1;
        __threadfence();
    }
    __syncthreads();
    const bool failed = sh->abort != 0;
    if (!failed && w_a >= 0 && w_a < Rc) {
        const int p = p0 + w_a;
        int ex = 0;
        for (int q = lane; q < p; q += 32) ex += __ldcg(a.len_next + q);
        ex = warp_sum_int(ex);                              // kept chain rows of the chains in front of this one
        for (int f0 = 0; f0 < F; f0 += 32) {
            const int f = f0 + lane;
            const int d = f < F ? __ldcg(a.keptdst + (int64_t)p * F + f) : -1;
            const unsigned m = __ballot_sync(FULL, d >= 0);
            if (d >= 0) {
                const int e = ex + __popc(m & ((1u << lane) - 1u));
                a.order_next[e] = d;
                a.chain_next[e] = p;
                a.rank_next[d] = e;
            }
            ex += __popc(m);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long kept_vis = (long long)sh->base_total - first;
        const long long n_merged = failed ? 0 : N - kept_vis;
        int ec = 0;
        if (failed || !((double)n_merged / (double)n_vis < a.bound)) ec = 3;    // top-k branch (or a wait gave up): the host redoes the call
        frame_finish(a, N, n_vis, n_merged, ec, failed ? 1 : 0);
    }
    pdl_trigger();
}

}  // namespace ff
