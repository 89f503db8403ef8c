// Single-pass streaming merge kernel: similarity + threshold select + run merge + compaction of hidden_states,
// cos / sin / patch_type / position ids in ONE sweep in which every row of hidden_states is read from HBM exactly
// once and never again (main.py:104-138, threshold branch — every merge call of a prefill but possibly the last).
//
// Shape of the problem.  Similarities couple a token only with the previous surviving token of the SAME patch id
// (a "chain", main.py:216-238), while the output is compacted in SEQUENCE order (main.py:132-138).  So:
//
//   * chains are owned: a persistent grid of <= one CTA per SM, each CTA owns `cpc` consecutive patch ids for the
//     whole kernel.  Rows of a chain arrive in shared memory by TMA (cp.async.bulk + mbarrier) into a small ring
//     of slots, several rows ahead; the previous row of the chain is still there, so sim(prev, cur) costs no
//     second read — not from HBM, not from L2.
//   * inside a chain the work is a pipeline of warps.  SIM warps take the rows round-robin: wait for the row,
//     compute sim(row t-1, row t) with the reference's rounding chain, publish the merge flag.  Nothing in that
//     depends on the merge itself, so similarities run ahead.  One MERGE warp per chain follows in row order
//     and holds the pending anchor in REGISTERS: a flagged row is added into it with one rounding to T per add
//     (the order torch-CPU index_add_ uses, main.py:304-311); when the run ends the sum is divided by T(L+1)
//     (main.py:314-317) and stored straight to its compacted position with 16-byte streaming stores, together
//     with its cos / sin / patch_type entries.  A slot is recycled — and the next row of the chain requested — by
//     whichever warp drops its last reference.
//   * the compacted position of a kept row i is the number of kept rows before i.  Every row gets one flag byte
//     in global memory (tagged with the call's epoch, so nothing is cleared between calls); the sim warp of row
//     t counts the flags between rows t-2 and t-1 of its own chain (the loads are issued before it waits for
//     the row, so the L2 round trip is hidden) and the merge warp adds these gap counts up.  Summed over the
//     grid every team reads all S flag bytes: a redundant, fully parallel scan instead of a look-back chain.
//
// Progress: a flag is published before its warp waits for anything else; a gap count only waits for flags of
// rows with a smaller sequence index than a row whose flag that chain has already published, so the wait-for
// relation strictly decreases in sequence index; all CTAs are co-resident (grid <= SM count, one CTA per SM);
// a slot's references (row t as "current", as "previous" of row t+1, as merge input) only need rows t and t+1
// resident, and rows are requested in order: no deadlock.
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

constexpr int ST_MAX_CHAINS = 8;                   // chains per CTA
constexpr int ST_MAX_SLOTS = 12;                   // row slots per chain
constexpr int ST_MIN_SLOTS = 3;
constexpr int ST_RING = 32;                        // depth of the per-chain hand-over rings: how far the positioner may trail
constexpr int ST_SCRATCH = 8;                      // averaged anchors per chain waiting in global memory for their position
constexpr int ST_MAX_LEN = 256;                    // rows per chain the index window holds
constexpr int ST_MAX_VPL = 16;                     // 16-byte vectors per lane: rows up to 8 KB
constexpr int ST_MAX_TMA_AUX = 6;                  // (aux, plane) pairs carried through the slot, <= 512 bytes each
constexpr int ST_MAX_SMALL_AUX = 2;                // 8-byte aux rows (patch_type, position ids)
constexpr int ST_TRACE_T = 96, ST_TRACE_K = 8;    // development aid: rows x stamps per chain
constexpr int ST_POLL_NS = 250;                    // back-off of the shared-memory polls: spinning warps steal issue slots
constexpr int ST_MAX_WARPS = 20;                   // cpc * (n_sim + 2): 640 threads leave 96 registers each

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
    int n_slots;                                   // per chain
    int n_ids;                                     // chain buckets; bucket n_ids = rows outside the chains
    int cpc;                                       // chains per CTA
    int n_sim;                                     // sim warps per chain
    int lag;                                       // the sim warp of row t counts the gap that positions row t - lag
    const int* order;                              // chain lists: order[base[id] + t] = sequence index
    const int* base;
    const int* len;
    int* order_next;                               // same layout, indices of the compacted sequence
    int* len_next;
    char* scratch;                                 // [n_ids][ST_SCRATCH][row_bytes] averaged anchors on their way out
    uint8_t* state;                                // per sequence row: (tag << 1) | merged
    float* sim_seq;                                // per sequence row: similarity with the chain predecessor (introspection)
    int* dst;                                      // per sequence row: compacted position or -1
    int64_t* counters;
    int64_t* counters_next;
    int64_t* status;
    float thr;
    double bound;
    unsigned tag;                                  // 1..127
    long long* trace;                              // development aid: [chain][ST_TRACE_T][ST_TRACE_K] globaltimer stamps, or null
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

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// stamp k of row t of chain `id` (one lane)
#define ST_STAMP(t, k) do { if (a.trace && (t) < ST_TRACE_T) a.trace[((size_t)id * ST_TRACE_T + (t)) * ST_TRACE_K + (k)] = gtime(); } while (0)

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
        const int l = max(lo - b0, 0), h = min(hi - b0, 4);          // bytes [l, h) of this word are in range
        uint32_t m = 0;
        if (h > l) m = (0xffffffffu >> (8 * (4 - h))) & (0xffffffffu << (8 * l));
        if (((w[q] & 0xfefefefeu) ^ tagm) & m) *ok = false;
        kept += __popc(~w[q] & 0x01010101u & m);
    }
    return kept;
}

__device__ __forceinline__ int count_kept16(const uint8_t* state, int at, int lo, int hi, unsigned tagm, bool* ok) {
    return count_kept_vec(ld_flags16(state + at), at, lo, hi, tagm, ok);
}

// ---- the kernel --------------------------------------------------------------------------------------------
// A kept row ("anchor") between its opening and its flush.  Shared by three parties under ChainShared::lock:
// the sim warps close it (L is final, its average is in the scratch ring), the positioner gives it its compacted
// position, whoever completes the pair hands it to the flusher.
struct AnchorRec {
    int L, pos, t;                                  // t = chain position of the anchor
    short closed, positioned;
    int live;                                       // from the opening until it is handed to the flusher
};

struct FlushJob {
    int L, pos, k, t_a;
};

// per-chain shared state (one per chain of the CTA, after the slots in dynamic shared memory)
struct ChainShared {
    uint64_t bars[ST_MAX_SLOTS];                    // one mbarrier per slot
    AnchorRec arec[ST_RING];                        // the k-th kept row of the chain at [k % ST_RING]
    int scr_busy[ST_SCRATCH];                       // scratch entry holds an average the flusher has not copied yet
    FlushJob fq[ST_RING];                           // flusher queue
    int idx[ST_MAX_LEN];                            // sequence index of every row of the chain
    int slot_of[ST_RING];                           // ((row + 1) << 8) | (parity << 4) | slot
    int kc[ST_RING];                                // (kept rows of this chain before row t) << 1 | row t was kept
    int refcnt[ST_MAX_SLOTS];
    int uses[ST_MAX_SLOTS];
    int issued;                                     // rows requested so far
    int chain_id;
    int token_a;                                    // row whose section A may run next (sim warps, row order)
    int token_b;                                    // rows the positioner is done with
    int lock;                                       // guards arec[] and the flusher queue head
    int fq_head;                                    // jobs pushed
    int flushed;                                    // jobs the flusher has finished
    int a_done;                                     // section A of the last row has run: n_kept is final
    // owned by section A
    int anchor_t, anchor_slot, anchor_L;            // the open anchor
    int n_kept;                                     // anchors opened so far
    int hits;                                       // rows merged away so far
};

__device__ __forceinline__ int ring_wait(const int* slot, int want_tag, int tag_shift) {
    int v;
    while (((unsigned)(v = *(volatile const int*)slot) >> tag_shift) != (unsigned)want_tag) __nanosleep(ST_POLL_NS);
    return v;
}

__device__ __forceinline__ void chain_lock(ChainShared* cs) {
    while (atomicCAS(&cs->lock, 0, 1) != 0) __nanosleep(20);
    __threadfence_block();
}
__device__ __forceinline__ void chain_unlock(ChainShared* cs) {
    __threadfence_block();
    atomicExch(&cs->lock, 0);
}

// sequence index of row t of a chain: from the shared-memory window, from the list in global memory beyond it
__device__ __forceinline__ int row_index(const ChainShared* cs, const int* chain_order, int t) {
    return t < ST_MAX_LEN ? cs->idx[t] : __ldg(chain_order + t);
}

// requests the next row of the chain into `slot` (called by one lane, only by whoever freed the slot)
__device__ __noinline__ void issue_row(const StreamArgs& a, ChainShared* cs, const int* chain_order, unsigned char* slots,
                                       int slot, int len) {
    const int row = atomicAdd(&cs->issued, 1);
    if (row >= len) return;
    if (a.trace) { const int id = cs->chain_id; ST_STAMP(row, 0); }
    const int parity = cs->uses[slot] & 1;
    cs->uses[slot] += 1;
    cs->refcnt[slot] = row + 1 < len ? 3 : 2;               // current of sim(row), previous of sim(row + 1), merge input / anchor
    __threadfence_block();
    *(volatile int*)&cs->slot_of[row & (ST_RING - 1)] = ((row + 1) << 8) | (parity << 4) | slot;
    const int i = row_index(cs, chain_order, row);
    const uint32_t bar = smem_u32(&cs->bars[slot]);
    mbar_expect_tx(bar, (uint32_t)a.row_bytes);
    tma_load(smem_u32(slots + (size_t)slot * a.slot_bytes), a.hidden + (size_t)i * a.row_bytes, (uint32_t)a.row_bytes, bar);
}

// drops `n` references of a slot; the last one out requests the next row of the chain into it (one lane)
__device__ __forceinline__ void release_slot(const StreamArgs& a, ChainShared* cs, const int* chain_order, unsigned char* slots,
                                             int slot, int n, int len) {
    __threadfence_block();                                  // our reads of the slot are done before the count drops
    if (atomicSub(&cs->refcnt[slot], n) == n) {
        __threadfence_block();                              // ... and everybody else's before the TMA engine overwrites it
        fence_async_smem();
        issue_row(a, cs, chain_order, slots, slot, len);
    }
}

// hands a closed and positioned anchor to the flusher (one lane, chain lock held)
__device__ __forceinline__ void push_flush(ChainShared* cs, AnchorRec& r, int k) {
    FlushJob& j = cs->fq[cs->fq_head & (ST_RING - 1)];
    j.L = r.L; j.pos = r.pos; j.k = k; j.t_a = r.t;
    r.live = 0;
    __threadfence_block();
    *(volatile int*)&cs->fq_head = cs->fq_head + 1;
}

__device__ __forceinline__ uint4 ld_cg16(const void* p) {                  // L2-coherent 16-byte load (scratch ring)
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

// ---- rows outside the chains (text, ids >= patch_num): bucket n_ids of the chain lists.  CTA b takes the rows
// k = b, b + G, ...  Their flags are published before anything else (nothing depends on them being late), the rows
// themselves are copied once the flags of everything before them are there.
__device__ __noinline__ void text_publish(const StreamArgs& a) {
    const int lane = threadIdx.x & 31;
    const int n_text = __ldg(a.len + a.n_ids);
    const int tbase = __ldg(a.base + a.n_ids);
    const int G = gridDim.x;
    for (int k = blockIdx.x + G * lane; k < n_text; k += G * 32) {
        const int i = __ldg(a.order + tbase + k);
        st_flag(a.state + i, a.tag << 1);
    }
    __syncwarp();
}

__device__ __noinline__ void text_copy(const StreamArgs& a, unsigned tag4) {
    const int lane = threadIdx.x & 31;
    const int n_text = __ldg(a.len + a.n_ids);
    const int tbase = __ldg(a.base + a.n_ids);
    const int G = gridDim.x;
    int cursor = 0, cnt = 0;                                // kept rows in [0, cursor)
    for (int k = blockIdx.x; k < n_text; k += G) {
        const int i = __ldg(a.order + tbase + k);
        // advance the scan to i, 512 bytes per step, waiting for unpublished flags
        int at0 = cursor & ~15;
        while (at0 < i) {
            const int at = at0 + lane * 16;
            bool ok = true;
            int kept = 0;
            if (at < i) kept = count_kept16(a.state, at, cursor, i, tag4, &ok);
            if (!__all_sync(FULL, ok)) { __nanosleep(500); continue; }
            cnt += warp_sum_int(kept);
            at0 += 32 * 16;
            cursor = at0 < i ? at0 : i;
        }
        cursor = i;
        const int pos = cnt;
        // copy the row and its aux rows through registers
        const char* src = a.hidden + (size_t)i * a.row_bytes;
        char* o = a.out + (size_t)pos * a.row_bytes;
#pragma unroll 4
        for (int v = lane; v < a.nvec; v += 32) st_stream16(o + (size_t)v * 16, ld_stream16(src + (size_t)v * 16));
#pragma unroll 1
        for (int q = 0; q < a.n_tma_aux; ++q) {
            const StreamAux& x = a.tma_aux[q];
            if (lane * 16 < x.bytes)
                st_stream16(x.dst + (size_t)pos * x.bytes + lane * 16, ld_stream16(x.src + (size_t)i * x.bytes + lane * 16));
        }
        if (lane == 0) {
            if (a.n_small_aux > 0) *reinterpret_cast<unsigned long long*>(a.small_aux[0].dst + (size_t)pos * 8) =
                __ldg(reinterpret_cast<const unsigned long long*>(a.small_aux[0].src + (size_t)i * 8));
            if (a.n_small_aux > 1) *reinterpret_cast<unsigned long long*>(a.small_aux[1].dst + (size_t)pos * 8) =
                __ldg(reinterpret_cast<const unsigned long long*>(a.small_aux[1].src + (size_t)i * 8));
            a.dst[i] = pos;
            a.sim_seq[i] = -2.0f;
            a.order_next[tbase + k] = pos;
        }
    }
    if (blockIdx.x == 0 && lane == 0) a.len_next[a.n_ids] = n_text;
}

template <int DT>
__global__ void __launch_bounds__(ST_MAX_WARPS * 32, 1)
k_stream_merge(const __grid_constant__ StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_last_cta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = a.n_sim + 2;                             // warps per chain: sims, positioner, flusher
    const int chain = warp / wpc, role = warp - chain * wpc;
    const unsigned tag4 = (a.tag * 0x01010101u) << 1;       // the tag as it sits in every flag byte

    ChainShared* cs_base = reinterpret_cast<ChainShared*>(smem + (size_t)a.cpc * a.n_slots * a.slot_bytes);

    // ---- set-up: the first warp of every chain fills the chain's shared state
    if (chain < a.cpc && role == 0) {
        ChainShared* cs = cs_base + chain;
        const int id = blockIdx.x * a.cpc + chain;
        const int len = id < a.n_ids ? __ldg(a.len + id) : 0;
        const int cbase = id < a.n_ids ? __ldg(a.base + id) : 0;
        for (int e = lane; e < ST_MAX_LEN; e += 32) cs->idx[e] = e < len ? __ldg(a.order + cbase + e) : 0;
        for (int e = lane; e < ST_RING; e += 32) { cs->slot_of[e] = 0; cs->kc[e] = 0; cs->arec[e].live = 0; }
        if (lane < ST_SCRATCH) cs->scr_busy[lane] = 0;
        if (lane < ST_MAX_SLOTS) { cs->refcnt[lane] = 0; cs->uses[lane] = 0; }
        if (lane == 0) {
            cs->issued = 0;
            cs->chain_id = id;
            cs->token_a = 0; cs->token_b = 0; cs->lock = 0; cs->fq_head = 0; cs->flushed = 0; cs->a_done = 0;
            cs->anchor_t = -1; cs->anchor_slot = -1; cs->anchor_L = 0;
            cs->n_kept = 0; cs->hits = 0;
            for (int b = 0; b < a.n_slots; ++b) mbar_init(smem_u32(&cs->bars[b]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    if (chain < a.cpc) {
        ChainShared* cs = cs_base + chain;
        unsigned char* slots = smem + (size_t)chain * a.n_slots * a.slot_bytes;
        const int id = blockIdx.x * a.cpc + chain;
        const int len = id < a.n_ids ? __ldg(a.len + id) : 0;
        const int cbase = id < a.n_ids ? __ldg(a.base + id) : 0;
        const int* chain_order = a.order + cbase;
        char* scratch = a.scratch + (size_t)id * ST_SCRATCH * a.row_bytes;       // averaged anchors waiting for their position

        if (role < a.n_sim) {
            // =========================== sim warp ===========================
            // Rows t = role, role + n_sim, ...; task t == len only closes the last anchor.  A task: similarity and merge
            // flag of its row (free-running), then section A — entered in row order through the chain's token —
            // which adds the row into the open anchor or closes that one and opens a new one.  A closed anchor that
            // absorbed rows is averaged into the chain's scratch ring (global memory, L2 resident); its slot is free
            // again at once.  Sim warps never wait for another chain.
            if (role == 0 && lane == 0)
                for (int b = 0; b < a.n_slots && b < len; ++b) issue_row(a, cs, chain_order, slots, b, len);

            for (int t = role; t <= len && len > 0; t += a.n_sim) {
                const bool have_row = t < len;
                int s_cur = -1, hit = 0;
                if (have_row) {
                    const int i = row_index(cs, chain_order, t);
                    if (lane == 0) ST_STAMP(t, 1);                            // task starts
                    // ---- the row and its predecessor: slots and arrival
                    int e_cur = 0, e_last = 0;
                    if (lane == 0) {
                        e_cur = ring_wait(&cs->slot_of[t & (ST_RING - 1)], t + 1, 8);
                        if (t > 0) e_last = ring_wait(&cs->slot_of[(t - 1) & (ST_RING - 1)], t, 8);
                    }
                    e_cur = __shfl_sync(FULL, e_cur, 0);
                    e_last = __shfl_sync(FULL, e_last, 0);
                    s_cur = e_cur & 15;
                    const int s_last = e_last & 15;
                    mbar_wait(smem_u32(&cs->bars[s_cur]), (e_cur >> 4) & 1);
                    float sim = -2.0f;
                    if (lane == 0) ST_STAMP(t, 2);                            // row arrived
                    if (t > 0) {
                        mbar_wait(smem_u32(&cs->bars[s_last]), (e_last >> 4) & 1);
                        // ---- similarity with the previous row of the chain (main.py:345-349 rounding chain)
                        const uint4* lrow = reinterpret_cast<const uint4*>(slots + (size_t)s_last * a.slot_bytes);
                        const uint4* crow = reinterpret_cast<const uint4*>(slots + (size_t)s_cur * a.slot_bytes);
                        float2 dot2 = make_float2(0.f, 0.f), na2 = dot2, nb2 = dot2;
#pragma unroll 2                                             // rolled: the I-cache is 32 KB, every role must stay small
                        for (int v = lane; v < a.nvec; v += 32) acc_pair2<DT>(lrow[v], crow[v], dot2, na2, nb2);
                        const float dot = warp_sum(dot2.x + dot2.y);
                        const float na = warp_sum(na2.x + na2.y);
                        const float nb = warp_sum(nb2.x + nb2.y);
                        sim = finish_cosine<DT>(dot, na, nb);
                        hit = sim >= a.thr;
                    }
                    if (lane == 0) {
                        st_flag(a.state + i, (a.tag << 1) | (unsigned)hit);
                        a.sim_seq[i] = sim;
                        if (hit) a.dst[i] = -1;
                        ST_STAMP(t, 3);                                       // flag published
                        // the previous row is not read again by this task
                        if (t > 0) release_slot(a, cs, chain_order, slots, s_last, 1, len);
                    }
                }

                // ---- section A (row order)
                if (lane == 0) {
                    while (*(volatile int*)&cs->token_a != t) __nanosleep(ST_POLL_NS);
                    // the rings are ST_RING deep: stay within reach of the positioner
                    while (*(volatile int*)&cs->token_b < t - (ST_RING - 4)) __nanosleep(2 * ST_POLL_NS);
                    if (have_row) ST_STAMP(t, 5);
                }
                __syncwarp();
                __threadfence_block();
                int c_t = -1, c_slot = -1, c_L = 0, c_k = 0;                  // the anchor this section closes
                if (have_row && hit) {
                    const int an_slot = cs->anchor_slot;
                    uint4* arow = reinterpret_cast<uint4*>(slots + (size_t)an_slot * a.slot_bytes);
                    const uint4* crow = reinterpret_cast<const uint4*>(slots + (size_t)s_cur * a.slot_bytes);
#pragma unroll 2
                    for (int v = lane; v < a.nvec; v += 32) arow[v] = add_round<DT>(arow[v], crow[v]);
                    if (lane == 0) {
                        cs->anchor_L += 1;
                        cs->hits += 1;
                        cs->kc[t & (ST_RING - 1)] = cs->n_kept << 1;
                    }
                } else {
                    c_t = cs->anchor_t; c_slot = cs->anchor_slot; c_L = cs->anchor_L;
                    c_k = cs->n_kept - 1;                                     // the open anchor is the latest kept row
                    __syncwarp();
                    if (lane == 0) {
                        if (have_row) {                                       // this row is the new anchor
                            const int n_kept = cs->n_kept;
                            AnchorRec& r = cs->arec[n_kept & (ST_RING - 1)];
                            // the record and the flusher queue are rings: wait for their previous occupants
                            while (*(volatile int*)&r.live || *(volatile int*)&cs->flushed < n_kept - (ST_RING - 8)) __nanosleep(2 * ST_POLL_NS);
                            chain_lock(cs);
                            r.L = 0; r.pos = -1; r.t = t; r.closed = 0; r.positioned = 0; r.live = 1;
                            chain_unlock(cs);
                            cs->anchor_t = t; cs->anchor_slot = s_cur; cs->anchor_L = 0;
                            cs->kc[t & (ST_RING - 1)] = (n_kept << 1) | 1;
                            cs->n_kept = n_kept + 1;
                        } else {
                            cs->anchor_t = -1;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    *(volatile int*)&cs->token_a = t + 1;
                    if (have_row) ST_STAMP(t, 6);                             // section A left
                }

                // ---- free-running again: finish the anchor that was closed, then the references of this task
                if (c_t >= 0) {
                    if (c_L > 0) {
                        // average of the run (main.py:314-317) -> scratch ring, where the flusher picks it up
                        if (lane == 0) {
                            while (*(volatile int*)&cs->scr_busy[c_k % ST_SCRATCH]) __nanosleep(2 * ST_POLL_NS);
                            cs->scr_busy[c_k % ST_SCRATCH] = 1;
                        }
                        __syncwarp();
                        const uint4* arow = reinterpret_cast<const uint4*>(slots + (size_t)c_slot * a.slot_bytes);
                        char* srow = scratch + (size_t)(c_k % ST_SCRATCH) * a.row_bytes;
                        const Divider<DT> dv(c_L + 1);
#pragma unroll 1
                        for (int v = lane; v < a.nvec; v += 32) *reinterpret_cast<uint4*>(srow + (size_t)v * 16) = dv.vec(arow[v]);
                        __threadfence();                                      // visible device-wide before it is announced
                    }
                    __syncwarp();
                    if (lane == 0) {
                        chain_lock(cs);
                        AnchorRec& r = cs->arec[c_k & (ST_RING - 1)];
                        r.L = c_L;
                        r.closed = 1;
                        if (r.positioned) push_flush(cs, r, c_k);
                        chain_unlock(cs);
                        release_slot(a, cs, chain_order, slots, c_slot, 1, len);     // the anchor's own reference
                    }
                }
                if (lane == 0) {
                    if (have_row) {
                        // this row as "current", and as merge input when it was merged away (an anchor keeps that
                        // reference until it is closed)
                        release_slot(a, cs, chain_order, slots, s_cur, hit ? 2 : 1, len);
                    } else {
                        __threadfence_block();
                        *(volatile int*)&cs->a_done = 1;
                    }
                }
                __syncwarp();
            }
        } else if (role == a.n_sim) {
            // =========================== positioner ===========================
            // Walks the chain's rows in order and gives every kept one its compacted position: the kept rows outside
            // the chain before it (the flag bytes between consecutive rows of the chain, counted gap by gap) + the
            // kept rows of the chain before it.  The only warp of the chain that waits for other chains.
            int cum = 0;
            // the flag bytes of the next gap are requested one row ahead (when the gap fits two 16-byte loads per lane),
            // so that the L2 round trip overlaps the bookkeeping of the current row
            auto gap_bounds = [&](int r, int& lo, int& hi) {
                hi = row_index(cs, chain_order, r);
                lo = (r >= 1 ? row_index(cs, chain_order, r - 1) : -1) + 1;
            };
            uint4 pf0 = make_uint4(0, 0, 0, 0), pf1 = pf0;
            bool pf_valid = false;
            if (len > 0) {
                int lo, hi;
                gap_bounds(0, lo, hi);
                const int al = lo & ~15;
                pf_valid = (hi - al) <= 1024;
                if (pf_valid && al + lane * 16 < hi) pf0 = ld_flags16(a.state + al + lane * 16);
                if (pf_valid && al + 512 + lane * 16 < hi) pf1 = ld_flags16(a.state + al + 512 + lane * 16);
            }
            for (int r = 0; r < len; ++r) {
                int f_lo, f_hi;
                gap_bounds(r, f_lo, f_hi);
                const int f_al = f_lo & ~15;
                uint4 fl0 = pf0, fl1 = pf1;
                bool first = pf_valid;                      // the prefetched vectors cover the whole gap
                if (r + 1 < len) {                          // request the next gap now
                    int lo, hi;
                    gap_bounds(r + 1, lo, hi);
                    const int al = lo & ~15;
                    pf_valid = (hi - al) <= 1024;
                    if (pf_valid && al + lane * 16 < hi) pf0 = ld_flags16(a.state + al + lane * 16);
                    if (pf_valid && al + 512 + lane * 16 < hi) pf1 = ld_flags16(a.state + al + 512 + lane * 16);
                }
                int kept = 0;
                for (int chunk = f_al; chunk < f_hi; chunk += 1024) {
                    const int at0 = chunk + lane * 16, at1 = at0 + 512;
                    for (;;) {
                        bool ok = true;
                        int k2 = 0;
                        if (!first) {
                            if (at0 < f_hi) fl0 = ld_flags16(a.state + at0);
                            if (at1 < f_hi) fl1 = ld_flags16(a.state + at1);
                        }
                        if (at0 < f_hi) k2 += count_kept_vec(fl0, at0, f_lo, f_hi, tag4, &ok);
                        if (at1 < f_hi) k2 += count_kept_vec(fl1, at1, f_lo, f_hi, tag4, &ok);
                        if (__all_sync(FULL, ok)) { kept += warp_sum_int(k2); first = false; break; }
                        if (!first) __nanosleep(3 * ST_POLL_NS);
                        first = false;
                    }
                }
                cum += kept;
                if (lane == 0) {
                    while (*(volatile int*)&cs->token_a <= r) __nanosleep(ST_POLL_NS);   // section A of row r has run
                    __threadfence_block();
                    const int kc = cs->kc[r & (ST_RING - 1)];
                    if (kc & 1) {
                        chain_lock(cs);
                        AnchorRec& rec = cs->arec[(kc >> 1) & (ST_RING - 1)];
                        rec.pos = cum + (kc >> 1);
                        rec.positioned = 1;
                        if (rec.closed) push_flush(cs, rec, kc >> 1);
                        chain_unlock(cs);
                    }
                    __threadfence_block();
                    *(volatile int*)&cs->token_b = r + 1;
                    ST_STAMP(r, 4);                                            // row r positioned
                }
                __syncwarp();
            }
        } else {
            // =========================== flusher ===========================
            // Copies closed and positioned anchors to their compacted positions: the row from hidden_states (still in
            // L2 — it went through the TMA a moment ago) or, when it absorbed other rows, from the scratch ring; the
            // cos / sin / patch_type entries straight from their source tensors.
            if (chain == 0) text_publish(a);                  // the CTA's share of the rows outside the chains
            int done = 0;
            for (;;) {
                int have = 0, fin = 0;
                if (lane == 0) {
                    for (;;) {
                        if (*(volatile int*)&cs->fq_head > done) { have = 1; break; }
                        if (len == 0 || (*(volatile int*)&cs->a_done && done == *(volatile int*)&cs->n_kept)) { fin = 1; break; }
                        __nanosleep(2 * ST_POLL_NS);
                    }
                }
                have = __shfl_sync(FULL, have, 0);
                fin = __shfl_sync(FULL, fin, 0);
                if (fin && !have) break;
                __threadfence_block();
                const FlushJob& j = cs->fq[done & (ST_RING - 1)];
                const int L = j.L, pos = j.pos, k = j.k, t_a = j.t_a;
                const int i_a = row_index(cs, chain_order, t_a);
                __syncwarp();

                // the whole row is requested at once (up to 16 x 512 bytes in flight per warp): one L2 round trip per row
                char* orow = a.out + (size_t)pos * a.row_bytes;
                uint4 buf[ST_MAX_VPL];
                if (L > 0) {
                    const char* srow = scratch + (size_t)(k % ST_SCRATCH) * a.row_bytes;
#pragma unroll
                    for (int q = 0; q < ST_MAX_VPL; ++q)
                        if (lane + 32 * q < a.nvec) buf[q] = ld_cg16(srow + (size_t)(lane + 32 * q) * 16);
                } else {
                    const char* srow = a.hidden + (size_t)i_a * a.row_bytes;
#pragma unroll
                    for (int q = 0; q < ST_MAX_VPL; ++q)
                        if (lane + 32 * q < a.nvec) buf[q] = ld_stream16(srow + (size_t)(lane + 32 * q) * 16);
                }
#pragma unroll
                for (int q = 0; q < ST_MAX_VPL; ++q)
                    if (lane + 32 * q < a.nvec) st_stream16(orow + (size_t)(lane + 32 * q) * 16, buf[q]);
#pragma unroll 1
                for (int q = 0; q < a.n_tma_aux; ++q) {
                    const StreamAux& x = a.tma_aux[q];
                    if (lane * 16 < x.bytes)
                        st_stream16(x.dst + (size_t)pos * x.bytes + lane * 16, ld_stream16(x.src + (size_t)i_a * x.bytes + lane * 16));
                }
                if (lane == 0) {
                    if (a.n_small_aux > 0) *reinterpret_cast<unsigned long long*>(a.small_aux[0].dst + (size_t)pos * 8) =
                        __ldg(reinterpret_cast<const unsigned long long*>(a.small_aux[0].src + (size_t)i_a * 8));
                    if (a.n_small_aux > 1) *reinterpret_cast<unsigned long long*>(a.small_aux[1].dst + (size_t)pos * 8) =
                        __ldg(reinterpret_cast<const unsigned long long*>(a.small_aux[1].src + (size_t)i_a * 8));
                    a.dst[i_a] = pos;
                    a.order_next[cbase + k] = pos;
                    ST_STAMP(t_a, 7);                                          // anchor written
                }
                __syncwarp();
                ++done;
                if (lane == 0) {
                    if (L > 0) *(volatile int*)&cs->scr_busy[k % ST_SCRATCH] = 0;   // the scratch entry of this job is free again
                    *(volatile int*)&cs->flushed = done;
                }
            }
            if (lane == 0) {
                // the chain is finished: every kept row is written
                if (id < a.n_ids) a.len_next[id] = cs->n_kept;
                if (cs->hits) atomicAdd((unsigned long long*)&a.counters[C_COUNT], (unsigned long long)cs->hits);
            }
            if (chain == 0) text_copy(a, tag4);
        }
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
    int cpc, grid, n_slots, slot_bytes, threads, n_sim, lag;
    size_t smem;
};

// Returns false when the shape is outside the single-pass kernel (caller falls back to the generic path).
inline bool plan_stream(int sm_count, int max_smem, int64_t row_bytes, int n_ids, StreamPlan* p) {
    if (n_ids < 1) return false;
    const int cpc = (n_ids + sm_count - 1) / sm_count;
    if (cpc > ST_MAX_CHAINS) return false;
    int n_sim = ST_MAX_WARPS / cpc - 2;                    // warps per chain: sims + positioner + flusher
    if (n_sim > 4) n_sim = 4;
    if (n_sim < 1) return false;
    const int slot = (int)((row_bytes + 127) / 128 * 128);
    const size_t fixed = (size_t)cpc * sizeof(ChainShared) + 256;
    if ((size_t)max_smem < fixed + 64) return false;
    const size_t avail = (size_t)max_smem - fixed - 64;
    int n_slots = (int)(avail / ((size_t)cpc * slot));
    if (n_slots > ST_MAX_SLOTS) n_slots = ST_MAX_SLOTS;
    if (n_slots < ST_MIN_SLOTS) return false;
    p->cpc = cpc;
    p->grid = (n_ids + cpc - 1) / cpc;
    p->n_slots = n_slots;
    p->n_sim = n_sim;
    p->lag = n_slots >= 7 ? 3 : (n_slots >= 5 ? 2 : 1);    // rows t .. t + lag - 1 stay resident while merge(t) waits
    p->slot_bytes = slot;
    p->threads = cpc * (n_sim + 2) * 32;
    p->smem = (size_t)cpc * n_slots * slot + fixed;
    return true;
}

template <int DT>
inline int launch_stream_dt(const StreamArgs& a, const StreamPlan& p, cudaStream_t st) {
    static size_t attr_set = 0;
    if (p.smem > attr_set) {
        if (cudaFuncSetAttribute(k_stream_merge<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess)
            return FF_E_CUDA;
        attr_set = p.smem;
    }
    k_stream_merge<DT><<<p.grid, p.threads, p.smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? FF_OK : FF_E_CUDA;
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
