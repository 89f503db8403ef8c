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
//                     row whose successor merges opens a run: every member is added IN PLACE — the running sum lives in
//                     the stage slot of the run's newest member, whose original bytes nobody needs once the S warp is two
//                     frames further (so there are no accumulator rows and the ring is one stage deeper) — one rounding to
//                     T per add in chain order, and the closing add divides by T(L+1) (main.py:304-317) and sends the row
//                     out.  Nothing is read twice, not even from the L2;
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
#include "ff_merge.cuh"

namespace ff {

constexpr int FR_MAXR = 8;                         // chains per CTA
constexpr int FR_MAXSTAGES = 16;                   // ring stages (frames in shared memory)
constexpr int FR_NQ = 16;                          // frames the flag / destination rings in shared memory hold (> stages + 2)
constexpr int FR_META = 2560;                      // bytes of shared memory in front of the stages
constexpr long long FR_TIMEOUT_NS = 400ll * 1000 * 1000;   // a launch older than this gives up at its next wait
constexpr int FR_PW = 4;                           // 32-word chunks of flag words one poll of the prefix warp can fetch
#ifndef FR_G_EARLY_ADD
#define FR_G_EARLY_ADD 1                           // G warps add run members before the destination of the anchor is known
#endif
#ifndef FR_AUX_BATCH
#define FR_AUX_BATCH 1                             // frames an aux warp moves per round (1 or 2)
#endif
#ifndef FR_INFLIGHT
#define FR_INFLIGHT 16                             // bulk loads a CTA keeps outstanding (16: as many as the ring allows)
#endif
#ifndef FR_SW
#define FR_SW 1                                    // similarity warps per chain in the build for up to four chains per CTA
#endif
#ifndef FR_S_FENCE
#define FR_S_FENCE 0
#endif
#ifndef FR_FINISHER
#define FR_FINISHER 0                              // 1: a finisher warp closes the rows (rounding chain, flags, global writes) for the S warps
#endif
#ifndef FR_S_WIDE
#define FR_S_WIDE 2                                // 1: the S warps' loop keeps four vector pairs per lane in flight
#endif
#ifndef FR_EARLY_STATUS
#define FR_EARLY_STATUS 1                          // 1: aux warp 0 of CTA 0 sends the status block as soon as the last frame is decided
#endif
#ifndef FR_G_WIDE
#define FR_G_WIDE 0                                // 1: the G warps' adds keep four vector pairs per lane in flight
#endif
#ifndef FR_ACC_INPLACE
#define FR_ACC_INPLACE 1                           // 1: a run's sum lives in the stage slot of its newest member (no accumulator rows: one more stage)
#endif
#ifndef FR_TRACE
#define FR_TRACE 0                                 // 1: the time stamps / cycle counters of ff_debug_frame_trace are compiled in (tools/build_variant.sh trace -DFR_TRACE=1)
#endif
#ifndef FR_NPW
#define FR_NPW 2                                   // prefix warps (they take the frames in turn)
#endif
#ifndef FR_PF
#define FR_PF 1                                    // frames one poll looks at
#endif
#ifndef FR_POLL_NS
#define FR_POLL_NS 64                              // pause after a poll that retired nothing
#endif
constexpr int FR_TRACE_K = 8;                      // time stamps per (CTA, frame) of a traced launch
constexpr int FR_DK = 4;                           // 32-frame chunks of its chain's destinations an aux warp keeps in registers for the tail
struct FrameKept { int v0, v1, v2, v3; };          // lane l: destination of the chain's row of frame 32 c + l, or -1

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
    unsigned long long empty[FR_MAXSTAGES];        // mbarriers: the G warps are through the stage (count = chains of the CTA)
    unsigned long long qbar[FR_NQ];                // mbarriers: the row sums of frame f (slot f % FR_NQ) are in psum[] for every chain
    unsigned long long sbar[FR_NQ];                // mbarriers: the merge flags of frame f are in kept[] for every chain
    unsigned long long pbar[FR_NQ];                // mbarriers: the destinations of frame f are in dstv[]
    volatile int a_done[FR_MAXR];                  // frames aux warp w has finished
    volatile int abort;
    volatile int base_total;                       // kept rows before the first row behind the span
    volatile int base_next;                        // kept rows before the next frame to be published (prefix warps)
    volatile int published;                        // CTA 0: the status block went out before the end of the kernel
    volatile int dstv[FR_NQ][FR_MAXR];
    volatile unsigned char kept[FR_NQ][FR_MAXR];
    volatile float psum[FR_NQ][FR_MAXR][2];        // float32 row sums of frame f, chain w: T(a*b) products, squares of row f
};
static_assert(sizeof(FrameShared) <= FR_META, "FR_META too small");

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fr_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void fr_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void fr_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
#ifndef FR_SUSPEND_NS
#define FR_SUSPEND_NS 0                            // suspend-time hint of a try_wait: the warp sleeps until the phase completes or this long
#endif
__device__ __forceinline__ bool fr_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if FR_SUSPEND_NS > 0
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t)FR_SUSPEND_NS) : "memory");
#else
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
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
__device__ __forceinline__ void fr_tma_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
    // Hand-overs between the warps of a CTA go through mbarriers: a waiting warp is suspended by the hardware (try_wait)
    // instead of polling — polling warps took 40 % of the issue slots of the warps that work.
    __device__ __forceinline__ bool wait_bar(unsigned long long* bar, int use) const {
        const uint32_t b = fr_smem_u32(bar), parity = (uint32_t)(use & 1);
        int spins = 0;
        while (!fr_mbar_try_wait(b, parity)) {
            if (sh->abort || expired(++spins)) return false;
        }
        return true;
    }
    // until the merge flags of frames 0 .. f are there for every chain of the CTA (frames are finished in order)
    __device__ __forceinline__ bool wait_s(int f) const { return wait_bar(&sh->sbar[f % FR_NQ], f / FR_NQ); }
    __device__ __forceinline__ bool wait_p(int f) const { return wait_bar(&sh->pbar[f % FR_NQ], f / FR_NQ); }
    // whole warp: until *ctr >= need (the one polled counter left: the aux warps' progress, which nobody is close to)
    __device__ __forceinline__ bool wait_ge(const volatile int* ctr, int need, unsigned sleep_ns = 200) const {
        int spins = 0;
        while (*ctr < need) {
            if (sh->abort || expired(++spins)) return false;
            __nanosleep(sleep_ns);
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

// the aux rows of sequence row r -> destination row d, one 16- or 8-byte piece per lane and entry: all loads, then all stores
__device__ __forceinline__ void frame_aux_load(const AuxFlat& f, int r, int lane, uint4* v) {
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
}
__device__ __forceinline__ void frame_aux_store(const AuxFlat& f, int d, int lane, const uint4* v) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (e < f.n) {
            const int rb = f.row_bytes[e];
            char* o = f.dst[e] + (int64_t)d * rb;
            if (f.piece[e] == 8) { if (lane * 8 < rb) reinterpret_cast<uint2*>(o)[lane] = make_uint2(v[e].x, v[e].y); }
            else if (lane * 16 < rb) reinterpret_cast<uint4*>(o)[lane] = v[e];
        }
}
__device__ __forceinline__ void frame_aux(const AuxFlat& f, const AuxPack& aux, int r, int d, int lane) {
    if (f.n < 0) { gather_aux_rows(aux, r, d, lane); return; }
    uint4 v[8];
    frame_aux_load(f, r, lane, v);
    frame_aux_store(f, d, lane, v);
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

// (the shipped build carries no tracing code: every role's loop is on the critical path, and code added to the S warps' role
// has cost 8 us twice, profiles/r02_frame_kernel.md)
#if FR_TRACE
#define FR_NOTE(f, k, val) do { if (a.trace && (f) < a.trace_frames) a.trace[((int64_t)blockIdx.x * a.trace_frames + (f)) * FR_TRACE_K + (k)] = (val); } while (0)
#define FR_STAMP(f, k) do { if (a.trace && (f) < a.trace_frames) a.trace[((int64_t)blockIdx.x * a.trace_frames + (f)) * FR_TRACE_K + (k)] = fr_time(); } while (0)
#else
#define FR_NOTE(f, k, val) do { } while (0)
#define FR_STAMP(f, k) do { } while (0)
#endif

// what every role of a CTA knows (the roles are inlined: as separate functions they read the kernel arguments through
// memory and ran 40 % slower)
struct FrameGeo {
    FrameShared* sh;
    FrameCtx cx;
    int F, first, p0, Rc, NS, R, P, S, lane;
    long long N;
    int64_t rb;
    uint32_t stage_bytes, stages0, acc0;
};

__device__ __forceinline__ void fr_role_producer(const FrameArgs& a, const AuxPack& aux, const FrameGeo& g, int w_role) {
    FrameShared* const sh = g.sh;
    const FrameCtx cx = g.cx;
    const int F = g.F, first = g.first, p0 = g.p0, Rc = g.Rc, NS = g.NS, R = g.R, P = g.P, S = g.S, lane = g.lane;
    const long long N = g.N;
    const int64_t rb = g.rb;
    const uint32_t stage_bytes = g.stage_bytes, stages0 = g.stages0, acc0 = g.acc0;
    (void)F; (void)first; (void)p0; (void)Rc; (void)NS; (void)R; (void)P; (void)S; (void)lane; (void)N; (void)rb; (void)stage_bytes; (void)stages0; (void)acc0; (void)sh;
    // ================================ producer ================================
    if (lane == 0) {
        // (the stage of a frame and the phase of its barriers are counted along: NS is a run-time value, and f % NS, f / NS are
        // some twenty dependent instructions each — on the S and G warps' paths they were paid several times per frame)
        int st = 0, ph = 0;
        for (int f = 0; f < F; ++f, st = st + 1 == NS ? 0 : st + 1, ph ^= (st == 0)) {
            if (f >= NS) {
                // the stage is free once the G warps are through frame f - NS; the aux warps touch no stage, they only
                // have to stay within the flag / destination rings (FR_NQ frames)
                bool ok = true;
                ok = cx.wait_bar(&sh->empty[st], ph ^ 1);
                for (int w = 0; w < Rc && ok; ++w) ok = cx.wait_ge(&sh->a_done[w], f - (FR_NQ - 4));
                if (!ok) break;
            }
            if (FR_INFLIGHT < FR_MAXSTAGES && f >= FR_INFLIGHT) {
                // no more than FR_INFLIGHT loads outstanding: everything this SM asks of the memory system — the polls
                // of the prefix warp above all — waits behind the bulk data already requested
                const int fo = f - FR_INFLIGHT;
                const uint32_t bo = fr_smem_u32(&sh->full[fo % NS]), po = (uint32_t)((fo / NS) & 1);
                int spins = 0;
                bool ok = true;
                while (!fr_mbar_try_wait(bo, po)) {
                    if (sh->abort || cx.expired(++spins)) { ok = false; break; }
                }
                if (!ok) break;
            }
            const uint32_t bar = fr_smem_u32(&sh->full[st]);
            const uint32_t bytes = (uint32_t)(Rc * rb);
            fr_mbar_expect_tx(bar, bytes);
            fr_tma_load(stages0 + (uint32_t)st * stage_bytes, a.hidden + ((int64_t)first + (int64_t)f * P + p0) * rb, bytes, bar);
            FR_STAMP(f, 0);
        }
    }
}

template <int NPW>
__device__ __forceinline__ void fr_role_prefix(const FrameArgs& a, const AuxPack& aux, const FrameGeo& g, int w_role) {
    FrameShared* const sh = g.sh;
    const FrameCtx cx = g.cx;
    const int F = g.F, first = g.first, p0 = g.p0, Rc = g.Rc, NS = g.NS, R = g.R, P = g.P, S = g.S, lane = g.lane;
    const long long N = g.N;
    const int64_t rb = g.rb;
    const uint32_t stage_bytes = g.stage_bytes, stages0 = g.stages0, acc0 = g.acc0;
    (void)F; (void)first; (void)p0; (void)Rc; (void)NS; (void)R; (void)P; (void)S; (void)lane; (void)N; (void)rb; (void)stage_bytes; (void)stages0; (void)acc0; (void)sh;
    // ================================ prefix: FR_NPW warps take the frames in turn ================================
    // A poll of the flag words is a round trip through a memory system busy with bulk copies (1 - 2 us), and a frame
    // needs at least one: ONE warp doing the frames one after the other made that round trip the period of the whole
    // pipeline.  So warp k polls the frames k, k + FR_NPW, ..: it counts the kept rows of its frame and the prefixes of
    // the CTA's own rows as soon as every patch has reported, then takes the running base from the warp of the previous
    // frame through shared memory (p_done / base_next) and publishes.
    const int kw = w_role;
    int polls = 0, last_done = -1;
    for (int f = kw; f < F; f += NPW) {
        // nobody has all the flags of a frame before this CTA's own S warps are through it
        bool ok = true;
        ok = cx.wait_s(f);
        if (!ok) break;
        const int lo = first + f * P, hi = lo + P;      // rows of the frame
        const int w_lo = lo >> 5, w_hi = (hi - 1) >> 5;
        int total = 0;                                  // kept rows of the frame
        int mine[FR_MAXR];                              // kept rows of the frame in front of the CTA's row w (lane 0 .. whoever holds the word)
#pragma unroll
        for (int w = 0; w < FR_MAXR; ++w) mine[w] = 0;
        for (int c0 = w_lo; c0 <= w_hi && ok; c0 += 32) {
            const int wi = c0 + lane;
            unsigned need = 0u;
            if (wi <= w_hi) {
                need = 0xffffffffu;
                if (wi == w_lo) need &= 0xffffffffu << (lo & 31);
                if (wi == w_hi) need &= 0xffffffffu >> (31 - ((hi - 1) & 31));
            }
            unsigned long long v = 0ull;
            while (true) {
                if (need) v = ld_relaxed64(a.words + wi);
                if (__all_sync(FULL, ((unsigned)(v >> 32) & need) == need)) break;
                ++polls;
                if (sh->abort || ((polls & 63) == 0 && ld_relaxed32(a.gbar + 1) != 0u) || cx.expired(polls)) { ok = false; break; }
                __nanosleep(FR_POLL_NS);
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
#pragma unroll
            for (int w = 0; w < FR_MAXR; ++w) {
                const int r = lo + p0 + w;
                const int here = (w < Rc && (r >> 5) == wi) ? total + incl - k + __popc(km & ((1u << (r & 31)) - 1u)) : 0;
                mine[w] += __shfl_sync(FULL, here, (r >> 5) - c0 < 32 && (r >> 5) >= c0 ? (r >> 5) - c0 : 0);
            }
            total += __shfl_sync(FULL, incl, 31);
        }
        if (!ok) { sh->abort = 1; break; }
        // the frames in front of this one are published by the other prefix warps
        if (f > 0 && !cx.wait_p(f - 1)) break;
        if (lane == 0) {
            const int base = f == 0 ? first : sh->base_next;   // rows in front of the span are all kept
#pragma unroll
            for (int w = 0; w < FR_MAXR; ++w)
                if (w < Rc) sh->dstv[f % FR_NQ][w] = base + mine[w];
            sh->base_next = base + total;
            if (f == F - 1) sh->base_total = base + total;
            fr_mbar_arrive(fr_smem_u32(&sh->pbar[f % FR_NQ]));
            FR_STAMP(f, 2);
        }
        __syncwarp();
        last_done = f;
    }
#if FR_EARLY_STATUS
    if (blockIdx.x == 0 && lane == 0 && last_done == F - 1 && !sh->abort) {
        // This warp has just published the last frame: every flag of the sequence is in, the sizes and the branch are
        // decided, and the host is told NOW — the last rows, the barrier over the grid and the arrays of the next call
        // take a few more microseconds, which its work for the next call overlaps (ff_status_wait).  A wait that gives up
        // after this is still reported (FF_ST_INTERNAL): the host looks again at its next call.
        const long long n_merged = N - ((long long)sh->base_total - first);
        const long long n_vis = a.counters[C_NVIS];
        frame_finish(a, N, n_vis, n_merged, (double)n_merged / (double)n_vis < a.bound ? 0 : 3, 0);
        sh->published = 1;
    }
#endif
}

template <int DT>
__device__ __forceinline__ void fr_role_sim(const FrameArgs& a, const AuxPack& aux, const FrameGeo& g, int w_role) {
    FrameShared* const sh = g.sh;
    const FrameCtx cx = g.cx;
    const int F = g.F, first = g.first, p0 = g.p0, Rc = g.Rc, NS = g.NS, R = g.R, P = g.P, S = g.S, lane = g.lane;
    const long long N = g.N;
    const int64_t rb = g.rb;
    const uint32_t stage_bytes = g.stage_bytes, stages0 = g.stages0, acc0 = g.acc0;
    (void)F; (void)first; (void)p0; (void)Rc; (void)NS; (void)R; (void)P; (void)S; (void)lane; (void)N; (void)rb; (void)stage_bytes; (void)stages0; (void)acc0; (void)sh;
    // ================================ row sums, chain p0 + w ================================
    // The S warp of a chain only SUMS: T(a*b) products and the squares of the new row over the 448 vectors of a row
    // pair, two warp reductions, two floats into shared memory.  What follows — the rounding chain with its square roots
    // and its division, the flag, the global writes — is a few hundred DEPENDENT instructions per row and belongs to the
    // finisher warp, which does the four chains of a frame in four lanes at once: left here it was half of this warp's
    // time per row, and this warp's time per row is the period of the whole pipeline.
    const int w = w_role;
    if (w < Rc) {
        long long c_wait = 0, c_busy = 0, c_loop = 0, c_red = 0, c0 = clock64();   // (traced launches: where this warp's cycles go)
        float na_prev = 0.f;                                // (no finisher: the squares of the previous row)
        (void)na_prev;
        int st = 0, st_prev = NS - 1, ph = 0;               // stage of frame f, of frame f - 1, parity of f / NS (counted along, see the producer)
        for (int f = 0; f < F; ++f, st_prev = st, st = st + 1 == NS ? 0 : st + 1, ph ^= (st == 0)) {
            if (!cx.wait_bar(&sh->full[st], ph)) break;
            if (w == 0 && lane == 0) FR_STAMP(f, 5);    // the frame's rows are in shared memory
            if (FR_TRACE && a.trace) { const long long c1 = clock64(); c_wait += c1 - c0; c0 = c1; }
            const uint32_t cur = stages0 + (uint32_t)st * stage_bytes + (uint32_t)(w * rb) + lane * 16;
            const uint32_t prv = stages0 + (uint32_t)st_prev * stage_bytes + (uint32_t)(w * rb) + lane * 16;
            float2 d0 = make_float2(0.f, 0.f), d1 = d0, b0 = d0, b1 = d0;
#ifdef FR_DIAG_S_DIV
            const int nvec = a.nvec / FR_DIAG_S_DIV;     // diagnostic build: wrong similarities, a fraction of the arithmetic
#else
            const int nvec = a.nvec;
#endif
            if (f > 0) {
#if FR_S_WIDE == 2
                // groups of two vector pairs per lane, the next group's four loads requested before the arithmetic of the
                // current one (two register sets in turn): no load latency is exposed after the first group
                const int ng = nvec >> 6;
                float2 d2 = d0, d3 = d0, b2 = d0, b3 = d0;
                uint4 pa0, ca0, pa1, ca1, pb0, cb0, pb1, cb1;
                pa0 = ca0 = pa1 = ca1 = pb0 = cb0 = pb1 = cb1 = make_uint4(0, 0, 0, 0);
                if (ng > 0) { pa0 = fr_lds16(prv); ca0 = fr_lds16(cur); pa1 = fr_lds16(prv + 512); ca1 = fr_lds16(cur + 512); }
                int gi = 0;
#pragma unroll 1
                for (; gi + 1 < ng; gi += 2) {
                    const uint32_t o1 = (uint32_t)(gi + 1) * 1024u;
                    pb0 = fr_lds16(prv + o1); cb0 = fr_lds16(cur + o1); pb1 = fr_lds16(prv + o1 + 512); cb1 = fr_lds16(cur + o1 + 512);
                    fr_dot_nb<DT>(pa0, ca0, d0, b0);
                    fr_dot_nb<DT>(pa1, ca1, d1, b1);
                    if (gi + 2 < ng) {
                        const uint32_t o2 = (uint32_t)(gi + 2) * 1024u;
                        pa0 = fr_lds16(prv + o2); ca0 = fr_lds16(cur + o2); pa1 = fr_lds16(prv + o2 + 512); ca1 = fr_lds16(cur + o2 + 512);
                    }
                    fr_dot_nb<DT>(pb0, cb0, d2, b2);
                    fr_dot_nb<DT>(pb1, cb1, d3, b3);
                }
                if (gi < ng) {
                    fr_dot_nb<DT>(pa0, ca0, d0, b0);
                    fr_dot_nb<DT>(pa1, ca1, d1, b1);
                }
                for (int v = (ng << 6) + lane; v < nvec; v += 32) fr_dot_nb<DT>(fr_lds16(prv + (v - lane) * 16), fr_lds16(cur + (v - lane) * 16), d0, b0);
                d0 = make_float2(d0.x + d2.x, d0.y + d2.y); d1 = make_float2(d1.x + d3.x, d1.y + d3.y);
                b0 = make_float2(b0.x + b2.x, b0.y + b2.y); b1 = make_float2(b1.x + b3.x, b1.y + b3.y);
#elif FR_S_WIDE
                // four vector pairs per lane in flight, whole groups first (no predicate between the loads), then the rest
                int v = lane;
                float2 d2 = d0, d3 = d0, b2 = d0, b3 = d0;
#pragma unroll 1
                for (; v + 96 < nvec; v += 128) {
                    const uint32_t po = prv + (v - lane) * 16, co = cur + (v - lane) * 16;
                    const uint4 p0v = fr_lds16(po), c0v = fr_lds16(co), p1v = fr_lds16(po + 512), c1v = fr_lds16(co + 512);
                    const uint4 p2v = fr_lds16(po + 1024), c2v = fr_lds16(co + 1024), p3v = fr_lds16(po + 1536), c3v = fr_lds16(co + 1536);
                    fr_dot_nb<DT>(p0v, c0v, d0, b0);
                    fr_dot_nb<DT>(p1v, c1v, d1, b1);
                    fr_dot_nb<DT>(p2v, c2v, d2, b2);
                    fr_dot_nb<DT>(p3v, c3v, d3, b3);
                }
                for (; v < nvec; v += 32) fr_dot_nb<DT>(fr_lds16(prv + (v - lane) * 16), fr_lds16(cur + (v - lane) * 16), d0, b0);
                d0 = make_float2(d0.x + d2.x, d0.y + d2.y); d1 = make_float2(d1.x + d3.x, d1.y + d3.y);
                b0 = make_float2(b0.x + b2.x, b0.y + b2.y); b1 = make_float2(b1.x + b3.x, b1.y + b3.y);
#else
#pragma unroll 2
                for (int v = lane; v < nvec; v += 64) {
                    const uint4 pa = fr_lds16(prv + (v - lane) * 16), ca = fr_lds16(cur + (v - lane) * 16);
                    uint4 pb = make_uint4(0, 0, 0, 0), cb = pb;
                    const bool two = v + 32 < nvec;
                    if (two) { pb = fr_lds16(prv + (v - lane + 32) * 16); cb = fr_lds16(cur + (v - lane + 32) * 16); }
                    fr_dot_nb<DT>(pa, ca, d0, b0);
                    if (two) fr_dot_nb<DT>(pb, cb, d1, b1);
                }
#endif
            } else {
                for (int v = lane; v < nvec; v += 64) {
                    fr_nb<DT>(fr_lds16(cur + (v - lane) * 16), b0);
                    if (v + 32 < nvec) fr_nb<DT>(fr_lds16(cur + (v - lane + 32) * 16), b1);
                }
            }
            long long c_l = 0, c_r = 0;
            if (FR_TRACE && a.trace) c_l = clock64();
            const float dot = warp_sum((d0.x + d0.y) + (d1.x + d1.y));
            const float nb = warp_sum((b0.x + b0.y) + (b1.x + b1.y));
            if (FR_TRACE && a.trace) { c_r = clock64(); c_loop += c_l - c0; c_red += c_r - c_l; }
#if FR_FINISHER
            if (lane == 0) {
                sh->psum[f % FR_NQ][w][0] = dot;
                sh->psum[f % FR_NQ][w][1] = nb;
                fr_mbar_arrive(fr_smem_u32(&sh->qbar[f % FR_NQ]));
            }
#else
            float s = -2.0f;                            // IGNORE_TOKEN at chain heads (main.py:225-238)
            if (f > 0) s = finish_cosine<DT>(dot, na_prev, nb);
            na_prev = nb;
            const unsigned kept = !(f > 0 && s >= a.thr);   // NaN compares false: kept
            if (lane == 0) {
                const int r = first + f * P + p0 + w;
                const int64_t j = (int64_t)(p0 + w) * F + f;
                sh->kept[f % FR_NQ][w] = (unsigned char)kept;
                fr_mbar_arrive(fr_smem_u32(&sh->sbar[f % FR_NQ]));   // (global writes after the hand-over: nothing in the CTA waits for them)
                red_add64(a.words + (r >> 5), (1ull << (32 + (r & 31))) | ((unsigned long long)kept << (r & 31)));
                a.sim[j] = s;
                a.flag[j] = (uint8_t)(kept ^ 1u);
                if (w == 0) FR_STAMP(f, 1);
            }
#endif
            __syncwarp();
            if (FR_TRACE && a.trace) { const long long c1 = clock64(); c_busy += c1 - c0; c0 = c1; }
        }
        if (w == 0 && lane == 0) { FR_NOTE(0, 7, c_busy); FR_NOTE(1, 7, c_wait); FR_NOTE(4, 7, c_loop); FR_NOTE(5, 7, c_red); }
    }
}

// ================================ finisher: lane w closes chain p0 + w of every frame ================================
template <int DT>
__device__ __forceinline__ void fr_role_finish(const FrameArgs& a, const AuxPack& aux, const FrameGeo& g, int w_role) {
    FrameShared* const sh = g.sh;
    const FrameCtx cx = g.cx;
    const int F = g.F, first = g.first, p0 = g.p0, Rc = g.Rc, P = g.P, lane = g.lane;
    const int w = lane, p = p0 + lane;
    float na_prev = 0.f;                                    // the squares of the previous row of this lane's chain
    for (int f = 0; f < F; ++f) {
        if (!cx.wait_bar(&sh->qbar[f % FR_NQ], f / FR_NQ)) break;
        float s = -2.0f;                                    // IGNORE_TOKEN at chain heads (main.py:225-238)
        unsigned kept = 1u;
        if (w < Rc) {
            const float dot = sh->psum[f % FR_NQ][w][0], nb = sh->psum[f % FR_NQ][w][1];
            if (f > 0) s = finish_cosine<DT>(dot, na_prev, nb);
            na_prev = nb;
            kept = !(f > 0 && s >= a.thr);                  // NaN compares false: kept
            sh->kept[f % FR_NQ][w] = (unsigned char)kept;
        }
        __syncwarp();                                       // (orders the lanes' flags before lane 0's arrival)
        if (lane == 0) {
            fr_mbar_arrive(fr_smem_u32(&sh->sbar[f % FR_NQ]));
            FR_STAMP(f, 1);
        }
        if (w < Rc) {
            // (global writes after the hand-over: nothing in the CTA waits for them)
            const int r = first + f * P + p;
            const int64_t j = (int64_t)p * F + f;
            red_add64(a.words + (r >> 5), (1ull << (32 + (r & 31))) | ((unsigned long long)kept << (r & 31)));
            a.sim[j] = s;
            a.flag[j] = (uint8_t)(kept ^ 1u);
        }
    }
}

template <int DT>
__device__ __forceinline__ void fr_role_merge(const FrameArgs& a, const AuxPack& aux, const FrameGeo& g, int w_role) {
    FrameShared* const sh = g.sh;
    const FrameCtx cx = g.cx;
    const int F = g.F, first = g.first, p0 = g.p0, Rc = g.Rc, NS = g.NS, R = g.R, P = g.P, S = g.S, lane = g.lane;
    const long long N = g.N;
    const int64_t rb = g.rb;
    const uint32_t stage_bytes = g.stage_bytes, stages0 = g.stages0, acc0 = g.acc0;
    (void)F; (void)first; (void)p0; (void)Rc; (void)NS; (void)R; (void)P; (void)S; (void)lane; (void)N; (void)rb; (void)stage_bytes; (void)stages0; (void)acc0; (void)sh;
    const int w_g = w_role;
    // ================================ merge + rows out, chain p0 + w_g ================================
    // Step f looks one row ahead: if row f + 1 merges, it is added NOW — into the raw row f if that one is kept (it
    // opens the run), else into the running sum — and if row f + 2 is kept the same pass divides and the row goes
    // out.  A kept row with a kept successor goes out as it is.  So step f is the last reader of stage f.
    if (w_g < Rc) {
        const int w = w_g;
#if !FR_ACC_INPLACE
        const uint32_t accp = acc0 + (uint32_t)(w * rb);
#endif
        const int nvec = a.nvec;
        int L = 0, anchor_d = -1;
        long long c_wait = 0, c_busy = 0, c0 = clock64();
        int st = 0, st_nxt = 1, st_prv = NS - 1;            // stages of the frames f, f + 1, f - 1 (counted along, see the producer)
        for (int f = 0; f < F; ++f, st_prv = st, st = st_nxt, st_nxt = st_nxt + 1 == NS ? 0 : st_nxt + 1) {
            if (!cx.wait_s(min(f + 3, F) - 1)) break;
            if (FR_TRACE && a.trace) { const long long c1 = clock64(); c_wait += c1 - c0; c0 = c1; }
            const bool kept = sh->kept[f % FR_NQ][w] != 0;
            const bool nxt_kept = f + 1 < F ? sh->kept[(f + 1) % FR_NQ][w] != 0 : true;
            const bool nn_kept = f + 2 < F ? sh->kept[(f + 2) % FR_NQ][w] != 0 : true;
            const uint32_t cur = stages0 + (uint32_t)st * stage_bytes + (uint32_t)(w * rb);
            const uint32_t nxt = stages0 + (uint32_t)st_nxt * stage_bytes + (uint32_t)(w * rb);
#if FR_ACC_INPLACE
            // The running sum of a run lives in the slot of its NEWEST member: this step adds row f + 1 into what slot f holds
            // (the anchor itself, or the sum so far) and leaves the result in slot f + 1.  Nobody else wants that slot's
            // original bytes any more: its last reader is the S warp's pass over frame f + 2 (as the previous row), and this
            // step has waited for the flags of frame f + 2.  So no accumulator rows, and the ring is one stage deeper.
            const uint32_t accw = nxt;
#else
            const uint32_t accw = accp;
#endif
            if (kept) L = 0;
#if !FR_G_EARLY_ADD
            if (!cx.wait_p(f)) break;
#endif
            if (!nxt_kept) {
                // the arithmetic needs no destination: it runs ahead of the prefix warp
#if FR_ACC_INPLACE
                const uint32_t src = cur;                     // the anchor itself, or the running sum step f - 1 left there
#else
                const uint32_t src = kept ? cur : accp;       // the anchor itself, or the running sum
#endif
                ++L;
                if (nn_kept) {
                    const Divider<DT> dv(L + 1);
#if FR_G_WIDE
                    for (int v0 = lane; v0 < nvec; v0 += 128) {         // four vector pairs per lane in flight
                        uint4 y[4], x[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (v0 + 32 * q < nvec) { y[q] = fr_lds16(src + (v0 + 32 * q) * 16); x[q] = fr_lds16(nxt + (v0 + 32 * q) * 16); }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (v0 + 32 * q < nvec) fr_sts16(accw + (v0 + 32 * q) * 16, dv.vec_fast(Num<DT>::add_vec(y[q], x[q])));   // T(T(acc + member) / T(L + 1))
                    }
#else
#pragma unroll 2
                    for (int v = lane; v < nvec; v += 32) {
                        const uint4 y = fr_lds16(src + v * 16), x = fr_lds16(nxt + v * 16);
                        fr_sts16(accw + v * 16, dv.vec_fast(Num<DT>::add_vec(y, x)));    // T(T(acc + member) / T(L + 1))
                    }
#endif
                    fr_fence_async();
                } else {
#if FR_G_WIDE
                    for (int v0 = lane; v0 < nvec; v0 += 128) {
                        uint4 y[4], x[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (v0 + 32 * q < nvec) { y[q] = fr_lds16(src + (v0 + 32 * q) * 16); x[q] = fr_lds16(nxt + (v0 + 32 * q) * 16); }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (v0 + 32 * q < nvec) fr_sts16(accw + (v0 + 32 * q) * 16, Num<DT>::add_vec(y[q], x[q]));   // T(acc + member), main.py:304
                    }
#else
#pragma unroll 2
                    for (int v = lane; v < nvec; v += 32) {
                        const uint4 y = fr_lds16(src + v * 16), x = fr_lds16(nxt + v * 16);
                        fr_sts16(accw + v * 16, Num<DT>::add_vec(y, x));                  // T(acc + member), main.py:304
                    }
#endif
                }
            }
            if (kept) {
                if (FR_TRACE && a.trace) { const long long c1 = clock64(); c_busy += c1 - c0; c0 = c1; }
                if (!cx.wait_p(f)) break;
                if (FR_TRACE && a.trace) { const long long c1 = clock64(); c_wait += c1 - c0; c0 = c1; }
                anchor_d = sh->dstv[f % FR_NQ][w];
            }
            __syncwarp();
            if (lane == 0) {
                if (nxt_kept) { if (kept) fr_tma_store(a.out + (int64_t)anchor_d * rb, cur, (uint32_t)rb); }
                else if (nn_kept) fr_tma_store(a.out + (int64_t)anchor_d * rb, accw, (uint32_t)rb);
            }
            __syncwarp();
            if (lane == 0) {
                // the store of this step drains while the next step runs: what is known to have left shared memory now
                // is the store of the step before, and with it stage f - 1 (and the accumulator row) can be reused
                fr_tma_commit();
                fr_tma_wait_read_1();
                if (f > 0) fr_mbar_arrive(fr_smem_u32(&sh->empty[st_prv]));
                if (w == 0) FR_STAMP(f, 3);
            }
            __syncwarp();
            if (FR_TRACE && a.trace) { const long long c1 = clock64(); c_busy += c1 - c0; c0 = c1; }
        }
        if (w == 0 && lane == 0) { FR_NOTE(2, 7, c_busy); FR_NOTE(3, 7, c_wait); }
        if (lane == 0) fr_tma_wait_all();
    }
}

__device__ __forceinline__ void fr_role_aux(const FrameArgs& a, const AuxPack& aux, const FrameGeo& g, int w_role, FrameKept& dk) {
    FrameShared* const sh = g.sh;
    const FrameCtx cx = g.cx;
    const int F = g.F, first = g.first, p0 = g.p0, Rc = g.Rc, NS = g.NS, R = g.R, P = g.P, S = g.S, lane = g.lane;
    const long long N = g.N;
    const int64_t rb = g.rb;
    const uint32_t stage_bytes = g.stage_bytes, stages0 = g.stages0, acc0 = g.acc0;
    (void)F; (void)first; (void)p0; (void)Rc; (void)NS; (void)R; (void)P; (void)S; (void)lane; (void)N; (void)rb; (void)stage_bytes; (void)stages0; (void)acc0; (void)sh;
    const int w_a = w_role;
    // ================================ aux rows, dst[], rows outside the chains ================================
    const int w = w_a;
    dk.v0 = dk.v1 = dk.v2 = dk.v3 = -1;                    // (read by the tail of the kernel)
    const int unit = blockIdx.x * R + w, n_units = gridDim.x * R;
    const int nvec = a.nvec;
    for (int t = unit; t < first; t += n_units) {       // rows in front of the span keep their place
        copy_row(a.hidden + (int64_t)t * rb, a.out + (int64_t)t * rb, nvec, lane);
        if (aux.n) frame_aux(a.auxf, aux, t, t, lane);
        if (lane == 0) { a.dst[t] = t; a.rank_next[t] = -1; }
    }
    bool ok = true;
    if (w < Rc) {
        // two frames per round: the loads of both rows are requested before either is stored (a round is one trip
        // through the memory system, and one frame per trip would be too slow)
        const int p = p0 + w;
        const bool flat = a.auxf.n >= 0;
        int nk = 0;
        for (int f = 0; f < F;) {
            if (!cx.wait_p(f)) { ok = false; break; }
            const int nb = 1;
            const bool k0 = sh->kept[f % FR_NQ][w] != 0, k1 = nb > 1 && sh->kept[(f + 1) % FR_NQ][w] != 0;
            const int d0 = k0 ? sh->dstv[f % FR_NQ][w] : -1, d1 = k1 ? sh->dstv[(f + 1) % FR_NQ][w] : -1;
            const int r0 = first + f * P + p, r1 = r0 + P;
            if (aux.n) {
                if (flat) {
                    uint4 v0[8], v1[8];
                    if (k0) frame_aux_load(a.auxf, r0, lane, v0);
                    if (k1) frame_aux_load(a.auxf, r1, lane, v1);
                    if (k0) frame_aux_store(a.auxf, d0, lane, v0);
                    if (k1) frame_aux_store(a.auxf, d1, lane, v1);
                } else {
                    if (k0) gather_aux_rows(aux, r0, d0, lane);
                    if (k1) gather_aux_rows(aux, r1, d1, lane);
                }
            }
            if ((f & 31) == lane) {
                const int c = f >> 5;
                dk.v0 = c == 0 ? d0 : dk.v0;
                dk.v1 = c == 1 ? d0 : dk.v1;
                dk.v2 = c == 2 ? d0 : dk.v2;
                dk.v3 = c == 3 ? d0 : dk.v3;
            }
            if (lane == 0) {
                a.dst[r0] = d0;
                a.keptdst[(int64_t)p * F + f] = d0;
                if (nb > 1) {
                    a.dst[r1] = d1;
                    a.keptdst[(int64_t)p * F + f + 1] = d1;
                }
                sh->a_done[w] = f + nb;
                if (w == 0) FR_STAMP(f, 4);
            }
            nk += (int)k0 + (int)k1;
            f += nb;
            __syncwarp();
        }
        if (lane == 0) a.len_next[p] = nk;
    }
    if (ok && cx.wait_p(F - 1)) {                      // rows behind the span move up by the merged rows
        const int bt = sh->base_total, n_post = S - first - (int)N;
        for (int t = unit; t < n_post; t += n_units) {
            const int r = first + (int)N + t, d = bt + t;
            copy_row(a.hidden + (int64_t)r * rb, a.out + (int64_t)d * rb, nvec, lane);
            if (aux.n) frame_aux(a.auxf, aux, r, d, lane);
            if (lane == 0) { a.dst[r] = d; a.rank_next[d] = -1; }
        }
    }
}

template <int DT, int MAXR, int NPW>
__global__ void __launch_bounds__(32 * (2 + NPW + 3 * MAXR), 1)
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
    const uint32_t acc0 = stages0 + (uint32_t)NS * stage_bytes;      // (FR_ACC_INPLACE = 0: the accumulator rows behind the stages)

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            fr_mbar_init(fr_smem_u32(&sh->full[s]), 1);
            fr_mbar_init(fr_smem_u32(&sh->empty[s]), Rc);
        }
        for (int q = 0; q < FR_NQ; ++q) {
            fr_mbar_init(fr_smem_u32(&sh->qbar[q]), Rc);
            fr_mbar_init(fr_smem_u32(&sh->sbar[q]), FR_FINISHER ? 1 : Rc);
            fr_mbar_init(fr_smem_u32(&sh->pbar[q]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sh->base_next = 0;
        sh->published = 0;
        sh->abort = 0;
        sh->base_total = 0;
    }
    if (threadIdx.x < FR_MAXR) sh->a_done[threadIdx.x] = 0;
    __syncthreads();
    FrameCtx cx;
    cx.sh = sh;
    cx.gbar = a.gbar;
    cx.t_end = fr_time() + FR_TIMEOUT_NS;
    if (threadIdx.x == 0) FR_STAMP(0, 6);                   // kernel under way

    const int w_s = wid - 2 - NPW, w_g = w_s - R, w_a = w_s - 2 * R;    // index inside the role

    FrameGeo g;
    g.sh = sh; g.cx = cx;
    g.F = F; g.first = first; g.p0 = p0; g.Rc = Rc; g.NS = NS; g.R = R; g.P = P; g.S = S; g.lane = lane;
    g.N = N; g.rb = rb; g.stage_bytes = stage_bytes; g.stages0 = stages0; g.acc0 = acc0;
    FrameKept dk;                                           // (aux warps only: set in their role, read in the tail)
    if (wid == 0) fr_role_producer(a, aux, g, 0);
    else if (wid <= NPW) fr_role_prefix<NPW>(a, aux, g, wid - 1);
    else if (wid == NPW + 1) { if (FR_FINISHER) fr_role_finish<DT>(a, aux, g, 0); }
    else if (w_s < R) fr_role_sim<DT>(a, aux, g, w_s);
    else if (w_g < R) fr_role_merge<DT>(a, aux, g, w_g);
    else if (w_a < R) fr_role_aux(a, aux, g, w_a, dk);

    // ---- every chain is through: one barrier over the grid, then the by-patch arrays of the next call
    __syncthreads();
    if (threadIdx.x == 0) {
        FR_STAMP(1, 6);                                     // this CTA's chains are through
        __threadfence();
        atomicAdd(a.gbar, 1u);
        int spins = 0;
        while (ld_relaxed32(a.gbar) < gridDim.x) {
            if (ld_relaxed32(a.gbar + 1) != 0u || cx.expired(++spins)) break;
            __nanosleep(100);
        }
        if (ld_relaxed32(a.gbar + 1) != 0u) sh->abort = 1;
        __threadfence();
        FR_STAMP(2, 6);                                     // the grid is through
    }
    __syncthreads();
    const bool failed = sh->abort != 0;
    if (!failed && w_a >= 0 && w_a < Rc) {
        // The tail is latency: everything it reads comes in as few round trips as possible.  The kept rows of the chains in
        // front of this one: 20 loads per lane in flight at once (unconditional — predicated loads are not batched; 640 chains
        // per trip), the destinations of the chain's own rows: out of this warp's registers for the first 32 FR_DK frames.
        const int p = p0 + w_a;
        int ex = 0;
        for (int q0 = 0; q0 < p; q0 += 32 * 20) {
            int v[20];
#pragma unroll
            for (int u = 0; u < 20; ++u) v[u] = __ldcg(a.len_next + min(q0 + 32 * u + lane, P - 1));
#pragma unroll
            for (int u = 0; u < 20; ++u) ex += q0 + 32 * u + lane < p ? v[u] : 0;
        }
        ex = warp_sum_int(ex);                              // kept chain rows of the chains in front of this one
        auto emit = [&](int d) {
            const unsigned m = __ballot_sync(FULL, d >= 0);
            if (d >= 0) {
                const int e = ex + __popc(m & ((1u << lane) - 1u));
                a.order_next[e] = d;
                a.chain_next[e] = p;
                a.rank_next[d] = e;
            }
            ex += __popc(m);
        };
        emit(dk.v0);
        if (F > 32) emit(dk.v1);
        if (F > 64) emit(dk.v2);
        if (F > 96) emit(dk.v3);
        for (int f0 = FR_DK * 32; f0 < F; f0 += 32) {
            const int f = f0 + lane;
            emit(f < F ? __ldcg(a.keptdst + (int64_t)p * F + f) : -1);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (!sh->published) {
            const long long kept_vis = (long long)sh->base_total - first;
            const long long n_merged = failed ? 0 : N - kept_vis;
            int ec = 0;
            if (failed || !((double)n_merged / (double)n_vis < a.bound)) ec = 3;    // top-k branch (or a wait gave up): the host redoes the call
            frame_finish(a, N, n_vis, n_merged, ec, failed ? 1 : 0);
        } else if (failed) {
            a.status[FF_ST_INTERNAL] = 1;                   // after the status block went out
            a.status[FF_ST_ERROR] = 3;
            __threadfence_system();
        }
        FR_STAMP(3, 6);                                     // kernel done
    }
    pdl_trigger();
}

}  // namespace ff
