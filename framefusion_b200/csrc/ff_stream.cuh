// Single-pass streaming merge kernel: similarity + threshold select + run merge + compaction of hidden_states,
// cos / sin / patch_type / position ids in ONE sweep in which every row of hidden_states is read from HBM exactly
// once and never again (main.py:104-138, threshold branch — every merge call of a prefill but possibly the last).
//
// Shape of the problem.  Similarities couple a token only with the previous surviving token of the SAME patch id
// (a "chain", main.py:216-238), while the output is compacted in SEQUENCE order (main.py:132-138).  So:
//
//   * chains are owned: a persistent grid of <= one CTA per SM, each CTA owns `cpc` consecutive patch ids, and
//     inside the CTA a TEAM of two warps owns one chain for the whole kernel.  A team walks its chain front to
//     back.  Rows arrive in shared memory by TMA (cp.async.bulk + mbarrier, a few rows ahead), the previous row
//     of the chain is still there, so sim(prev, cur) costs no second read — not from HBM, not from L2.
//   * the merge happens where the row already is: a kept row stays in its slot as the pending anchor; flagged
//     successors are added into it one at a time with a rounding to T per add (the order torch-CPU index_add_
//     uses, main.py:304-311), and when the run ends the sum is divided by T(L+1) (main.py:314-317) and the slot is
//     written to its compacted position with a TMA bulk store.  The aux rows (cos, sin) ride in the same slot.
//   * the compacted position of a kept row i is the number of kept rows before i.  Teams publish one flag byte
//     per row (tagged with the call's epoch, so nothing is cleared between calls) and every team counts the
//     flags between two consecutive rows of its own chain — a redundant, fully parallel scan: S bytes per team
//     from L2 instead of a serial look-back chain.  The count for row t-1 is taken while row t is being processed,
//     when those flags are one row-time old, so in steady state nobody waits; when CTAs drift apart the wait
//     is bounded by the slots a team has for rows in flight.
//
// Progress: a team publishes the flag of row t BEFORE it waits for anything, and only ever waits for flags of
// rows with a smaller sequence index than a row it has already published; the wait-for relation strictly
// decreases in sequence index, all CTAs are co-resident (grid <= SM count, one CTA per SM), TMA always
// completes: no deadlock.
//
// The branch decision (main.py:114-116) needs the global count, known only at the end: the kernel speculates
// on the threshold branch, the last CTA to finish checks count / n_vis < bound and otherwise reports
// FF_ST_ERROR = 3; the host then redoes the call with the generic multi-kernel path (top-k branch, at most once
// per prefill).  The input is never modified, so the redo sees the original rows.
//
// Rows outside the chains (text, ids >= patch_num) sit in bucket n_ids of the chain lists; one extra warp per
// CTA copies its share of them through registers.
#pragma once
#include "ff_common.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int ST_TEAM_WARPS = 2;
constexpr int ST_NT = ST_TEAM_WARPS * 32;          // threads per team
constexpr int ST_MAX_VPT = 8;                      // 16-byte vectors per thread: rows up to 8 KB
constexpr int ST_MAX_TEAMS = 7;                    // named barriers 1..7
constexpr int ST_MAX_SLOTS = 12;
constexpr int ST_MIN_SLOTS = 5;
constexpr int ST_IDX_WIN = 64;                     // chain-list window kept in shared memory (two halves of 32)
constexpr int ST_MAX_TMA_AUX = 6;                  // (aux, plane) pairs carried through the slot
constexpr int ST_MAX_SMALL_AUX = 2;                // 8-byte aux rows (patch_type, position ids) carried in registers

struct StreamAux {
    const char* src;
    char* dst;
    int bytes;                                     // per row
    int slot_off;                                  // offset inside the slot (TMA aux only)
};

struct StreamArgs {
    const char* hidden;
    char* out;
    int S;
    int nvec;                                      // 16-byte vectors per row
    int row_bytes;
    int slot_bytes;                                // row + aux area, multiple of 128
    int n_slots;                                   // per team
    int n_ids;                                     // chain buckets; bucket n_ids = rows outside the chains
    int cpc;                                       // chains per CTA
    int lag;                                       // iterations between a row's arrival and the scan that positions it
    const int* order;                              // chain lists: order[base[id] + t] = sequence index
    const int* base;
    const int* len;
    int* order_next;                               // same layout, indices of the compacted sequence
    int* len_next;
    uint8_t* state;                                // per sequence row: (tag << 1) | merged
    float* sim_seq;                                // per sequence row: similarity with the chain predecessor (introspection)
    int* dst;                                      // per sequence row: compacted position or -1
    int64_t* counters;
    int64_t* counters_next;
    int64_t* status;
    float thr;
    double bound;
    unsigned tag;                                  // 1..127
    int n_tma_aux, n_small_aux;
    StreamAux tma_aux[ST_MAX_TMA_AUX];
    StreamAux small_aux[ST_MAX_SMALL_AUX];
};

template <int DT>
__device__ __forceinline__ uint4 add_round(const uint4& a, const uint4& b) {      // T(a + b), elementwise
    float x[Num<DT>::EPV], y[Num<DT>::EPV];
    Num<DT>::unpack(a, x);
    Num<DT>::unpack(b, y);
#pragma unroll
    for (int e = 0; e < Num<DT>::EPV; ++e) x[e] = x[e] + y[e];
    return Num<DT>::pack(x);
}

// T(x / div) elementwise, div = T(L + 1) (main.py:314-317: one true division, rounded to T).
// bf16 fast path: a power-of-two divisor is an exact scaling; otherwise q0 = x * RN(1/div) is within 2 float32 ulp
// of the correctly rounded quotient, so both round to the same bf16 unless q0 sits within a few ulp of a bf16
// rounding boundary (low 16 bits ~ 0x8000) — those elements (about 1e-4 of them) take the IEEE division.
template <int DT>
struct Divider {
    float div, rcp;
    bool pow2;
    __device__ __forceinline__ explicit Divider(int n) {
        div = Num<DT>::rnd((float)n);
        rcp = 1.0f / div;
        pow2 = (n & (n - 1)) == 0;
    }
    __device__ __forceinline__ float one(float x) const {
        if (DT != FF_BF16) return x / div;
        const float q0 = x * rcp;
        if (pow2) return q0;
        const uint32_t u = __float_as_uint(q0);
        const bool risky = ((u & 0xffffu) - 0x7ff8u) <= 0x10u || ((u & 0x7f800000u) == 0u && (u << 1) != 0u);
        return risky ? x / div : q0;
    }
    __device__ __forceinline__ uint4 vec(const uint4& a) const {
        float x[Num<DT>::EPV];
        Num<DT>::unpack(a, x);
#pragma unroll
        for (int e = 0; e < Num<DT>::EPV; ++e) x[e] = one(x[e]);
        return Num<DT>::pack(x);
    }
};

// Row sums for one 16-byte vector pair: dot += T(a*b), nb += b*b (the norm of `a` is carried over from the previous
// row).  Two interleaved float32 accumulators per sum (even / odd elements), folded by the caller.
// bf16: the product tensor element T(a*b) is one mul.rn.bf16x2 (exact product, one rounding — what the reference's
// bf16 multiply does), sums run on the packed float32x2 pipe of sm_100.
template <int DT>
__device__ __forceinline__ void acc_dot_norm(const uint4& va, const uint4& vb, float2& dot, float2& nb) {
    if (DT == FF_BF16) {
        const uint32_t aw[4] = {va.x, va.y, va.z, va.w}, bw[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __nv_bfloat162 pa = *reinterpret_cast<const __nv_bfloat162*>(&aw[q]);
            __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&bw[q]);
            __nv_bfloat162 pp = __hmul2(pa, pb);
            const uint32_t pw = *reinterpret_cast<uint32_t*>(&pp);
            dot = __fadd2_rn(dot, make_float2(__uint_as_float(pw << 16), __uint_as_float(pw & 0xffff0000u)));
            const float2 bf = make_float2(__uint_as_float(bw[q] << 16), __uint_as_float(bw[q] & 0xffff0000u));
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
            nb.x = fmaf(b[e], b[e], nb.x);
            nb.y = fmaf(b[e + 1], b[e + 1], nb.y);
        }
    }
}

template <int DT>
__device__ __forceinline__ void acc_norm(const uint4& vb, float2& nb) {
    float b[Num<DT>::EPV];
    Num<DT>::unpack(vb, b);
#pragma unroll
    for (int e = 0; e < Num<DT>::EPV; e += 2) { nb.x = fmaf(b[e], b[e], nb.x); nb.y = fmaf(b[e + 1], b[e + 1], nb.y); }
}

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void team_bar(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(ST_NT) : "memory"); }

__device__ __forceinline__ uint4 ld_flags16(const uint8_t* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_flag(uint8_t* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u8 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// kept rows among the flag bytes [lo, hi) that fall into the 16-byte vector `v` loaded from byte offset `at`
// (16-aligned); *ok is cleared if a byte in range does not carry the call's tag yet.  tagm = tag4 << 1.
__device__ __forceinline__ int count_kept_vec(const uint4& v, int at, int lo, int hi, unsigned tagm, bool* ok) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int kept = 0;
    if (at >= lo && at + 16 <= hi) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if ((w[q] & 0xfefefefeu) != tagm) *ok = false;
            kept += __popc(~w[q] & 0x01010101u);
        }
        return kept;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int b0 = at + 4 * q;
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b0 + b >= lo && b0 + b < hi) m |= 0xffu << (8 * b);
        if (((w[q] & 0xfefefefeu) ^ tagm) & m) *ok = false;
        kept += __popc(~w[q] & 0x01010101u & m);
    }
    return kept;
}

__device__ __forceinline__ int count_kept16(const uint8_t* state, int at, int lo, int hi, unsigned tagm, bool* ok) {
    return count_kept_vec(ld_flags16(state + at), at, lo, hi, tagm, ok);
}

// ---- the kernel --------------------------------------------------------------------------------------------
// dynamic shared memory layout (per CTA):
//   [teams][n_slots] slots of slot_bytes           (128-byte aligned)
//   [teams][n_slots] mbarriers (8 bytes)
//   [teams] exchange area: 2 phases x ST_TEAM_WARPS x 4 words
//   [teams] chain-list window: ST_IDX_WIN ints
struct TeamXchg {
    float dot[2][ST_TEAM_WARPS];
    float nrm[2][ST_TEAM_WARPS];
    int kept[2][ST_TEAM_WARPS];
    int ok[2][ST_TEAM_WARPS];
};

template <int DT, int VPT>
__global__ void __launch_bounds__(ST_NT * ST_MAX_TEAMS + 32, 1)
k_stream_merge(const StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_last_cta;
    const int n_teams = a.cpc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int team = warp / ST_TEAM_WARPS;
    const unsigned tag4 = (a.tag * 0x01010101u) << 1;       // the tag as it sits in every flag byte

    unsigned char* slots_base = smem;
    uint64_t* bars_base = reinterpret_cast<uint64_t*>(smem + (size_t)n_teams * a.n_slots * a.slot_bytes);
    TeamXchg* xchg_base = reinterpret_cast<TeamXchg*>(bars_base + n_teams * a.n_slots);
    int* idx_base = reinterpret_cast<int*>(xchg_base + n_teams);
    uint64_t* small_base = reinterpret_cast<uint64_t*>(idx_base + n_teams * ST_IDX_WIN);

    if (threadIdx.x == 0) {
        for (int b = 0; b < n_teams * a.n_slots; ++b) mbar_init(smem_u32(bars_base + b), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int my_hits = 0;                                        // tokens merged away (counted by team thread 0)

    if (team < n_teams) {
        // =========================== chain team ===========================
        const int tid = threadIdx.x - team * ST_NT;          // 0..63
        const int tw = tid >> 5;                             // warp inside the team
        const int id = blockIdx.x * a.cpc + team;
        const int bar_id = 1 + team;
        const int K = a.lag;                                 // the position of row r is taken at iteration r + K
        unsigned char* slots = slots_base + (size_t)team * a.n_slots * a.slot_bytes;
        uint64_t* bars = bars_base + team * a.n_slots;
        TeamXchg* xc = xchg_base + team;
        int* s_idx = idx_base + team * ST_IDX_WIN;
        uint64_t* q_small = small_base + team * 8 * ST_MAX_SMALL_AUX;   // thread 0 only: 8-byte aux values of recent anchors
        const int len = id < a.n_ids ? __ldg(a.len + id) : 0;
        const int cbase = id < a.n_ids ? __ldg(a.base + id) : 0;

        // chain-list window: 64 consecutive entries of the chain, slid by 32 (see the refill below)
        if (tid < ST_IDX_WIN) s_idx[tid] = tid < len ? __ldg(a.order + cbase + tid) : 0;
        team_bar(bar_id);

        // --- slot bookkeeping.  Every thread of the team tracks it identically (it only depends on uniform data);
        //     thread 0 alone executes the TMA / bulk-group instructions.
        uint32_t free_mask = (1u << a.n_slots) - 1u;
        uint32_t parity = 0;                                 // per-slot mbarrier phase
        int issued = 0;                                      // rows whose load was issued
        int pend0 = -1, pend1 = -1;                          // slots with a bulk store in flight (oldest first)
        unsigned long long ring = 0;                         // slot of row r at bits 4*(r & 15): <= 12 rows in flight

        uint32_t tx_bytes = (uint32_t)a.row_bytes;
        for (int q = 0; q < a.n_tma_aux; ++q) tx_bytes += (uint32_t)a.tma_aux[q].bytes;

        auto issue_loads = [&](int upto) {                   // prefetch rows while slots are free
            while (issued < len && issued < upto && free_mask) {
                const int s = __ffs(free_mask) - 1;
                free_mask &= ~(1u << s);
                if (tid == 0) {
                    const int i = s_idx[issued & (ST_IDX_WIN - 1)];
                    const uint32_t bar = smem_u32(bars + s);
                    const uint32_t dst = smem_u32(slots + (size_t)s * a.slot_bytes);
                    mbar_expect_tx(bar, tx_bytes);
                    tma_load(dst, a.hidden + (size_t)i * a.row_bytes, (uint32_t)a.row_bytes, bar);
                    for (int q = 0; q < a.n_tma_aux; ++q)
                        tma_load(dst + a.tma_aux[q].slot_off, a.tma_aux[q].src + (size_t)i * a.tma_aux[q].bytes,
                                 (uint32_t)a.tma_aux[q].bytes, bar);
                }
                const int sh = 4 * (issued & 15);
                ring = (ring & ~(0xfull << sh)) | ((unsigned long long)s << sh);
                ++issued;
            }
        };
        auto retire_oldest = [&]() {                         // frees the slot of the oldest bulk store once it was read
            if (pend0 < 0) return;
            if (pend1 < 0) { if (tid == 0) tma_wait_read_0(); }
            else { if (tid == 0) tma_wait_read_1(); }
            free_mask |= 1u << pend0;
            pend0 = pend1;
            pend1 = -1;
        };

        // chain state (uniform across the team)
        int acc_slot = -1, last_slot = -1;                   // pending anchor / previous row of the chain
        int L = 0;                                           // members merged into the pending anchor
        int anchor_t = -1, anchor_i = -1;                    // its chain position and sequence index
        int anchor_pos = -1;                                 // its compacted position (-1: not known yet)
        float n_last = 0.f;                                  // |last|^2
        int cnt = 0;                                         // kept rows in [0, i_r) after the scan step of row r
        uint32_t kept_hist = 0;                              // bit (t & 31): row t of the chain was kept
        uint32_t q_fin = 0, q_slot4 = 0;                     // closed anchors waiting for their position, by (t & 7)
        int kept_idx = 0;                                    // kept rows of this chain written so far (next-call list position)
        uint64_t small_acc[ST_MAX_SMALL_AUX] = {0, 0};       // thread 0: 8-byte aux values of the pending anchor

        issue_loads(20);
        int phase = 0;

        // one kept row, final in its slot, goes to its compacted position; uniform call
        auto write_row = [&](int slot, int t_a, int i_a, int pos) {
            if (pend1 >= 0) retire_oldest();                 // at most two stores in flight
            if (tid == 0) {
                const uint32_t src = smem_u32(slots + (size_t)slot * a.slot_bytes);
                tma_store(a.out + (size_t)pos * a.row_bytes, src, (uint32_t)a.row_bytes);
                for (int q = 0; q < a.n_tma_aux; ++q)
                    tma_store(a.tma_aux[q].dst + (size_t)pos * a.tma_aux[q].bytes, src + a.tma_aux[q].slot_off,
                              (uint32_t)a.tma_aux[q].bytes);
                tma_commit();
                for (int q = 0; q < a.n_small_aux; ++q)
                    *reinterpret_cast<uint64_t*>(a.small_aux[q].dst + (size_t)pos * 8) = q_small[(t_a & 7) * ST_MAX_SMALL_AUX + q];
                a.dst[i_a] = pos;
                a.order_next[cbase + kept_idx] = pos;
            }
            if (pend0 < 0) pend0 = slot; else pend1 = slot;
            ++kept_idx;
        };

        // the run of the pending anchor is over: average it (main.py:314-317) and write it, or park it until its
        // position is known; uniform call
        auto close_anchor = [&]() {
            if (L > 0) {
                unsigned char* arow = slots + (size_t)acc_slot * a.slot_bytes;
                const Divider<DT> dv(L + 1);
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int v = tid + ST_NT * k;
                    if (v < a.nvec) {
                        uint4* p = reinterpret_cast<uint4*>(arow) + v;
                        *p = dv.vec(*p);
                    }
                }
                fence_async_smem();
                team_bar(bar_id);
            }
            if (tid == 0)
                for (int q = 0; q < a.n_small_aux; ++q) q_small[(anchor_t & 7) * ST_MAX_SMALL_AUX + q] = small_acc[q];
            if (anchor_pos >= 0) {
                write_row(acc_slot, anchor_t, anchor_i, anchor_pos);
            } else {
                const int sh = 4 * (anchor_t & 7);
                q_fin |= 1u << (anchor_t & 7);
                q_slot4 = (q_slot4 & ~(0xfu << sh)) | ((uint32_t)acc_slot << sh);
            }
        };

        // counts the kept rows in (from, to) exclusive, waiting until every flag there is published
        auto scan_between = [&](int from, int to) -> int {
            const int lo = from + 1, hi = to;
            if (hi <= lo) return 0;
            const int lo_al = lo & ~15;
            int total = 0;
            for (int chunk = lo_al; chunk < hi; chunk += ST_NT * 16) {
                const int at = chunk + tid * 16;
                for (;;) {
                    bool ok = true;
                    int kept = 0;
                    if (at < hi) kept = count_kept16(a.state, at, lo, hi, tag4, &ok);
                    kept = warp_sum_int(kept);
                    const int wok = __all_sync(FULL, ok);
                    if (lane == 0) { xc->kept[phase][tw] = kept; xc->ok[phase][tw] = wok; }
                    team_bar(bar_id);
                    int k2 = 0, o2 = 1;
#pragma unroll
                    for (int w = 0; w < ST_TEAM_WARPS; ++w) { k2 += xc->kept[phase][w]; o2 &= xc->ok[phase][w]; }
                    phase ^= 1;
                    if (o2) { total += k2; break; }
                    __nanosleep(100);
                }
            }
            return total;
        };

        for (int t = 0; t < len + K; ++t) {
            // slide the chain-list window: entries older than t - 8 make room for [t + 24, t + 56)
            if ((t & 31) == 8 && t >= 40) {
                const int e = t + 24 + tid;
                if (tid < 32) s_idx[e & (ST_IDX_WIN - 1)] = e < len ? __ldg(a.order + cbase + e) : 0;
                // visible to thread 0 after the next team barrier; loads are issued at most 20 rows ahead
            }
            const bool have_row = t < len;
            const int r = t - K;                             // the row whose position this iteration settles
            const int i = s_idx[t & (ST_IDX_WIN - 1)];
            const int s = (int)((ring >> (4 * (t & 15))) & 0xfull);
            unsigned char* crow = slots + (size_t)s * a.slot_bytes;

            // small aux rows of this token: issued now, consumed after the row arrived
            uint64_t small_new[ST_MAX_SMALL_AUX] = {0, 0};
            if (tid == 0 && have_row)
                for (int q = 0; q < a.n_small_aux; ++q)
                    small_new[q] = __ldg(reinterpret_cast<const uint64_t*>(a.small_aux[q].src + (size_t)i * 8));

            // flags between rows r-1 and r of the chain: loaded now, looked at after the similarity, so that the L2
            // round trip hides behind the arrival of the row and the arithmetic
            const int i_r = r >= 0 ? s_idx[r & (ST_IDX_WIN - 1)] : 0;
            const int i_r1 = r >= 1 ? s_idx[(r - 1) & (ST_IDX_WIN - 1)] : -1;
            const int f_lo = i_r1 + 1, f_hi = i_r;
            const int f_at = (f_lo & ~15) + tid * 16;
            const bool f_mine = r >= 0 && f_at < f_hi;
            const bool f_one_chunk = (f_hi - (f_lo & ~15)) <= ST_NT * 16;
            uint4 fl = make_uint4(0, 0, 0, 0);
            if (f_mine) fl = ld_flags16(a.state + f_at);

            float2 dot2 = make_float2(0.f, 0.f), nb2 = make_float2(0.f, 0.f);
            if (have_row) {
                mbar_wait(smem_u32(bars + s), (parity >> s) & 1u);
                parity ^= 1u << s;
                // ---- similarity with the previous row of the chain (main.py:345-349 rounding chain)
                if (t > 0) {
                    const unsigned char* lrow = slots + (size_t)last_slot * a.slot_bytes;
#pragma unroll
                    for (int k = 0; k < VPT; ++k) {
                        const int v = tid + ST_NT * k;
                        if (v < a.nvec) {
                            const uint4 x = reinterpret_cast<const uint4*>(lrow)[v];
                            const uint4 y = reinterpret_cast<const uint4*>(crow)[v];
                            acc_dot_norm<DT>(x, y, dot2, nb2);
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < VPT; ++k) {
                        const int v = tid + ST_NT * k;
                        if (v < a.nvec) acc_norm<DT>(reinterpret_cast<const uint4*>(crow)[v], nb2);
                    }
                }
            }
            bool f_ok = true;
            int f_kept = 0;
            if (f_mine) f_kept = count_kept_vec(fl, f_at, f_lo, f_hi, tag4, &f_ok);
            const float dot = warp_sum(dot2.x + dot2.y);
            const float nb = warp_sum(nb2.x + nb2.y);
            f_kept = warp_sum_int(f_kept);
            const int f_wok = __all_sync(FULL, f_ok);
            if (lane == 0) { xc->dot[phase][tw] = dot; xc->nrm[phase][tw] = nb; xc->kept[phase][tw] = f_kept; xc->ok[phase][tw] = f_wok; }
            team_bar(bar_id);
            float dsum = 0.f, nsum = 0.f;
            int ksum = 0, oksum = 1;
#pragma unroll
            for (int w = 0; w < ST_TEAM_WARPS; ++w) {
                dsum += xc->dot[phase][w]; nsum += xc->nrm[phase][w]; ksum += xc->kept[phase][w]; oksum &= xc->ok[phase][w];
            }
            phase ^= 1;

            if (have_row) {
                int hit = 0;
                float sim = -2.0f;
                if (t > 0) {
                    sim = finish_cosine<DT>(dsum, n_last, nsum);
                    hit = sim >= a.thr;
                }
                if (tid == 0) {
                    a.sim_seq[i] = sim;
                    st_flag(a.state + i, (a.tag << 1) | (unsigned)hit);
                    if (hit) a.dst[i] = -1;
                }
                my_hits += hit;

                // ---- merge into the pending anchor, or close its run and open a new one
                if (hit) {
                    unsigned char* arow = slots + (size_t)acc_slot * a.slot_bytes;
#pragma unroll
                    for (int k = 0; k < VPT; ++k) {
                        const int v = tid + ST_NT * k;
                        if (v < a.nvec) {
                            uint4* p = reinterpret_cast<uint4*>(arow) + v;
                            *p = add_round<DT>(*p, reinterpret_cast<const uint4*>(crow)[v]);
                        }
                    }
                    fence_async_smem();                      // ordered before the bulk store by a later team barrier
                    if (L > 0) free_mask |= 1u << last_slot; // nobody reads the old `last` any more
                    ++L;
                } else {
                    if (t > 0) {
                        close_anchor();
                        if (L > 0) free_mask |= 1u << last_slot;
                    }
                    acc_slot = s;
                    L = 0;
                    anchor_t = t;
                    anchor_i = i;
                    anchor_pos = -1;
#pragma unroll
                    for (int q = 0; q < ST_MAX_SMALL_AUX; ++q) small_acc[q] = small_new[q];
                    kept_hist |= 1u << (t & 31);
                }
                if (hit) kept_hist &= ~(1u << (t & 31));
                last_slot = s;
                n_last = nsum;
            } else if (t == len && len > 0) {
                close_anchor();                              // end of the chain: the last anchor
                if (L > 0) free_mask |= 1u << last_slot;
                anchor_t = -1;
            }

            // ---- position of row r: kept rows in [0, i_r).  The flags were usually all there; otherwise (or when
            //      the gap spans more than one chunk) poll until they are.
            if (r >= 0) {
                if (!(oksum && f_one_chunk)) ksum = scan_between(i_r1, i_r);
                const int prev_kept = r >= 1 ? (int)((kept_hist >> ((r - 1) & 31)) & 1u) : 0;
                cnt += prev_kept + ksum;
                if ((kept_hist >> (r & 31)) & 1u) {
                    if (q_fin & (1u << (r & 7))) {
                        q_fin &= ~(1u << (r & 7));
                        write_row((int)((q_slot4 >> (4 * (r & 7))) & 0xfu), r, i_r, cnt);
                    } else {
                        anchor_pos = cnt;                    // row r is the pending anchor, still open
                    }
                }
            }

            if (!free_mask && issued < len) retire_oldest();
            issue_loads(t + 1 + 20);
        }

        if (tid == 0) {
            tma_wait_all();
            if (id < a.n_ids) a.len_next[id] = kept_idx;
            if (my_hits) atomicAdd((unsigned long long*)&a.counters[C_COUNT], (unsigned long long)my_hits);
        }
    } else if (warp == n_teams * ST_TEAM_WARPS) {
        // =========================== rows outside the chains ===========================
        const int n_text = __ldg(a.len + a.n_ids);
        const int tbase = __ldg(a.base + a.n_ids);
        const int G = gridDim.x;
        // publish the flags first: nobody depends on anything here
        for (int k = blockIdx.x + G * lane; k < n_text; k += G * 32) {
            const int i = __ldg(a.order + tbase + k);
            st_flag(a.state + i, a.tag << 1);
        }
        __syncwarp();
        int cursor = 0, cnt = 0;                            // kept rows in [0, cursor)
        for (int k = blockIdx.x; k < n_text; k += G) {
            const int i = __ldg(a.order + tbase + k);
            // advance the scan to i, 512 bytes per step, waiting for unpublished flags
            int at0 = cursor & ~15;
            while (at0 < i) {
                const int at = at0 + lane * 16;
                bool ok = true;
                int kept = 0;
                if (at < i) kept = count_kept16(a.state, at, cursor, i, tag4, &ok);
                if (!__all_sync(FULL, ok)) { __nanosleep(200); continue; }
                cnt += warp_sum_int(kept);
                at0 += 32 * 16;
                cursor = at0 < i ? at0 : i;
            }
            cursor = i;
            const int pos = cnt;
            // copy the row and its aux rows through registers
            const char* src = a.hidden + (size_t)i * a.row_bytes;
            char* o = a.out + (size_t)pos * a.row_bytes;
            for (int v = lane; v < a.nvec; v += 32) st_stream16(o + (size_t)v * 16, ld_stream16(src + (size_t)v * 16));
            for (int q = 0; q < a.n_tma_aux; ++q) {
                const StreamAux& x = a.tma_aux[q];
                for (int v = lane; v < x.bytes / 16; v += 32)
                    reinterpret_cast<uint4*>(x.dst + (size_t)pos * x.bytes)[v] =
                        __ldg(reinterpret_cast<const uint4*>(x.src + (size_t)i * x.bytes) + v);
            }
            if (lane == 0) {
                for (int q = 0; q < a.n_small_aux; ++q)
                    *reinterpret_cast<uint64_t*>(a.small_aux[q].dst + (size_t)pos * 8) =
                        __ldg(reinterpret_cast<const uint64_t*>(a.small_aux[q].src + (size_t)i * 8));
                a.dst[i] = pos;
                a.sim_seq[i] = -2.0f;
                a.order_next[tbase + k] = pos;
            }
        }
        if (blockIdx.x == 0 && lane == 0) a.len_next[a.n_ids] = n_text;
    }

    // ---- last CTA out: the branch decision and the status block (main.py:112-127)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd((unsigned long long*)&a.counters[C_TICKET], 1ull);
        s_last_cta = (done == (unsigned long long)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last_cta && threadIdx.x == 0) {
        __threadfence();
        const long long count = *(volatile long long*)&a.counters[C_COUNT];
        const long long n_vis = a.counters[C_NVIS], N = a.counters[C_N];
        int err = 0;
        if (n_vis == 0) err = 1;
        else if (!((double)count / (double)n_vis < a.bound)) err = 3;       // top-k branch: not ours
        const long long s_keep = (long long)a.S - count;
        a.counters[C_SKEEP] = s_keep;
        a.counters[C_NMERGED] = count;
        a.counters[C_TICKET] = 0;
        a.counters_next[C_N] = N - count;
        a.counters_next[C_NVIS] = n_vis - count;
        a.counters_next[C_COUNT] = 0;
        a.counters_next[C_TICKET] = 0;
        a.counters_next[C_TICKET2] = 0;
        a.status[FF_ST_SEQ_KEEP] = s_keep;
        a.status[FF_ST_COUNT] = count;
        a.status[FF_ST_NVIS] = n_vis;
        a.status[FF_ST_NCHAIN] = N;
        a.status[FF_ST_BRANCH] = 0;
        a.status[FF_ST_TOPK] = 0;
        a.status[FF_ST_NMERGED] = count;
        a.status[FF_ST_FUSED] = 1;
        a.status[FF_ST_ERROR] = err;
    }
}

struct StreamPlan {
    int cpc, grid, n_slots, slot_bytes, threads, lag;
    size_t smem;
};

// Returns false when the shape is outside the single-pass kernel (caller falls back to the generic path).
inline bool plan_stream(int sm_count, int max_smem, int64_t row_bytes, int n_ids, int aux_bytes, StreamPlan* p) {
    if (n_ids < 1) return false;
    const int cpc = (n_ids + sm_count - 1) / sm_count;
    if (cpc > ST_MAX_TEAMS) return false;
    const int slot = (int)((row_bytes + aux_bytes + 127) / 128 * 128);
    const size_t fixed = (size_t)cpc * (sizeof(TeamXchg) + ST_IDX_WIN * 4 + 8 * ST_MAX_SMALL_AUX * 8) + 256;
    const size_t static_smem = 64;
    if ((size_t)max_smem < fixed + static_smem) return false;
    size_t avail = (size_t)max_smem - fixed - static_smem;
    int n_slots = (int)(avail / ((size_t)cpc * (slot + 8)));
    if (n_slots > ST_MAX_SLOTS) n_slots = ST_MAX_SLOTS;
    if (n_slots < ST_MIN_SLOTS) return false;
    p->cpc = cpc;
    p->grid = (n_ids + cpc - 1) / cpc;
    p->n_slots = n_slots;
    p->lag = n_slots >= 7 ? 3 : (n_slots == 6 ? 2 : 1);    // worst case held: anchor + last + (lag - 1) parked + 2 stores
    p->slot_bytes = slot;
    p->threads = cpc * ST_NT + 32;
    p->smem = (size_t)cpc * n_slots * (slot + 8) + fixed;
    return true;
}

template <int DT, int VPT>
inline int launch_stream_t(const StreamArgs& a, const StreamPlan& p, cudaStream_t st) {
    static size_t attr_set = 0;
    if (p.smem > attr_set) {
        if (cudaFuncSetAttribute(k_stream_merge<DT, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess)
            return FF_E_CUDA;
        attr_set = p.smem;
    }
    k_stream_merge<DT, VPT><<<p.grid, p.threads, p.smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? FF_OK : FF_E_CUDA;
}

template <int DT>
inline int launch_stream_dt(const StreamArgs& a, const StreamPlan& p, cudaStream_t st) {
    const int vpt = (a.nvec + ST_NT - 1) / ST_NT;
    if (vpt <= 2) return launch_stream_t<DT, 2>(a, p, st);
    if (vpt <= 4) return launch_stream_t<DT, 4>(a, p, st);
    if (vpt <= 7) return launch_stream_t<DT, 7>(a, p, st);
    return launch_stream_t<DT, 8>(a, p, st);
}

inline int launch_stream(int dtype, const StreamArgs& a, const StreamPlan& p, cudaStream_t st) {
    switch (dtype) {
        case FF_BF16: return launch_stream_dt<FF_BF16>(a, p, st);
        case FF_F16: return launch_stream_dt<FF_F16>(a, p, st);
        case FF_F32: return launch_stream_dt<FF_F32>(a, p, st);
    }
    return FF_E_UNSUPPORTED;
}

}  // namespace ff
