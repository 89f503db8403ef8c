// Read-once merge kernel: similarity + threshold select + run merge + compaction of hidden_states and of the aux
// tensors (cos / sin / patch_type / position ids) in ONE launch that fetches every row of hidden_states from HBM
// exactly once (main.py:104-138, threshold branch — every merge call of a prefill but possibly the last).
//
// Shape of the problem.  A token is compared with the previous surviving token of the SAME patch id (its chain
// predecessor, main.py:216-238) — 576 rows (4 MB) back on the first call of a uniform video — while the output is
// compacted in SEQUENCE order (main.py:132-138), and a kept row can only be written once the flags of its chain
// successors are known (they are averaged into it, main.py:285-317).  So:
//
//   * tiles of FU_WARPS consecutive rows are handed out in sequence order by a ticket; warp w of the CTA owns row
//     tile * W + w.  Its own row (from HBM) and its chain predecessor (pred[r], an L2 hit: that row was some
//     tile's "own" row a few microseconds ago) are staged by TMA (cp.async.bulk + mbarrier) into the warp's two
//     shared-memory slots — no registers are tied up while the rows travel.  Three row sums out of shared memory, warp
//     shuffles, the reference's rounding chain -> sim, flag.
//   * compacted position = kept rows before it: tile aggregate + decoupled look-back over the tile descriptors
//     (one warp per tile), then one 8-byte state word per row is published: kept + destination row, or merged away +
//     (destination row of the run's anchor, members so far).
//   * runs are accumulated IN the output: a merged-away row r adds itself to the running sum of its run, which lives
//     in the anchor's destination row — T(acc + r) with one rounding to T per add, in chain order, exactly the
//     sequence torch-CPU index_add_ performs (main.py:304-311); nothing is lost by parking the sum in a row of T
//     because every partial sum is a value of T already.  Its state word carries (anchor destination, members so
//     far), so the next row of the chain needs no walk.  The first member finds both operands in its two slots; a
//     later one fetches the running sum (an L2 hit, by TMA, while the look-back runs).
//   * the CLOSING row finishes: a row that is NOT flagged ends the run of its predecessor.  Predecessor a plain kept
//     row (the common case): its staged copy leaves shared memory with one TMA bulk store — no second read, no
//     registers.  Predecessor merged away: the running sum is divided once by T(L+1) (main.py:314-317) in place.
//     Chain tails close their own run.  Every dependency points to a row with a SMALLER sequence index, tickets are
//     taken in order by CTAs that are running, so the lowest unfinished tile never waits: no deadlock, whatever is
//     resident.
//   * the links of the next call (pred / succ of every kept row, by destination index) fall out of the same step:
//     the closer knows both ends.
//   * the aux rows (cos / sin / patch_type / position ids) of a tile are copied by worker warps of the same CTA, fed
//     through a small ring in shared memory: their latency never sits on the tile warps' path.
//
// The branch decision (main.py:114-116) needs the global count, known only at the end: the kernel speculates on the
// threshold branch, the CTA of the last tile checks count / n_vis < bound and otherwise reports FF_ST_ERROR = 3; the
// host then redoes the call with the multi-kernel path (top-k branch, at most once per prefill).  The input is
// never modified, so the redo sees the original rows.
#pragma once
#include "ff_common.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FU_WARPS = 8;                        // most rows per tile = warps per CTA (fewer when the rows are long)
constexpr int FU_WORKERS = 4;                      // warps per CTA besides the tile warps: one scan warp + aux workers
constexpr int FU_QSIZE = 32;                       // ring entries between the scan warp and the aux workers
constexpr int FU_SCANQ = 8;                        // ring entries between the tile warps and the scan warp
constexpr int FU_SPIN_LIMIT = 1 << 18;             // polls (~64 ns apart) before a wait gives up and reports FF_ST_INTERNAL
constexpr unsigned long long FU_AGG = 1ull << 32, FU_INCL = 2ull << 32;

struct FusedArgs {
    const char* hidden;
    char* out;
    int S, nvec, row_bytes, slot_bytes, ntiles, tile_rows;
    const int2* link;                              // [S] (pred, succ): row index, -1 = chain head / tail, -2 = not a chain row
    int2* link_next;                               // [S_keep] the same for the compacted sequence
    unsigned long long* fstate;                    // [S] zero on entry
    unsigned long long* desc;                      // [1 + ntiles] zero on entry: ticket, tile descriptors
    unsigned long long* fstate_clr;                // other bank: cleared for the next call
    unsigned long long* desc_clr;
    float* sim_seq;                                // [S] similarity with the chain predecessor (introspection)
    int* dst;                                      // [S] destination row or -1
    int64_t* counters;
    int64_t* counters_next;
    int64_t* status;
    float thr;
    double bound;
};

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
__device__ __forceinline__ void tma_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// state word of row x, once it has been published (code != 0).  Gives up after FU_SPIN_LIMIT polls: *err is set and the
// caller skips what depended on it, so the kernel always terminates.
__device__ __forceinline__ unsigned long long wait_state(const unsigned long long* fstate, int x, int* err) {
    unsigned long long v = ld_relaxed64(fstate + x);
    int spins = 0;
    while ((v & 3ull) == 0ull) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(64);
        v = ld_relaxed64(fstate + x);
    }
    return v;
}

// the aux rows of sequence row r are wanted a few microseconds from now (first plane of each tensor, first 128 bytes
// per lane q; rows are at most a few hundred bytes): pull them into the L2 behind the hidden_states row
__device__ __forceinline__ void prefetch_aux(const AuxPack& aux, int r, int lane) {
#pragma unroll
    for (int q = 0; q < FF_MAX_AUX; ++q)
        if (q < aux.n && (lane >> 2) == q) {
            const int64_t off = (int64_t)(lane & 3) * 128;
            if (off < aux.a[q].row_bytes) prefetch_l2((const char*)aux.a[q].src + (int64_t)r * aux.a[q].row_bytes + off);
        }
}

__device__ __forceinline__ void copy_row(const char* src, char* dst, int nvec, int lane) {
    for (int vb = 0; vb < nvec; vb += 256) {                // eight 16-byte vectors per lane in flight
        const int v0 = vb + lane;
        uint4 x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (v0 + 32 * q < nvec) x[q] = ld_stream16(src + (int64_t)(v0 + 32 * q) * 16);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (v0 + 32 * q < nvec) st_stream16(dst + (int64_t)(v0 + 32 * q) * 16, x[q]);
    }
}

// state word of a row: bits [1:0] 0 = not yet known, 1 = merged away, 2 = kept; bits [33:2] destination row (of the row
// itself if kept, of its run's anchor if merged away); bits [63:34] members of the run so far (merged away only)
__device__ __forceinline__ unsigned long long state_kept(int d) { return 2ull | ((unsigned long long)(uint32_t)d << 2); }
__device__ __forceinline__ unsigned long long state_merged(int d_anchor, int L) {
    return 1ull | ((unsigned long long)(uint32_t)d_anchor << 2) | ((unsigned long long)(uint32_t)L << 34);
}
__device__ __forceinline__ int state_type(unsigned long long st) { return (int)(st & 3ull); }
__device__ __forceinline__ int state_dst(unsigned long long st) { return (int)(uint32_t)(st >> 2); }
__device__ __forceinline__ int state_len(unsigned long long st) { return (int)(st >> 34); }

// Decoupled look-back over the tile descriptors (one warp per tile): the tile's kept-row count is posted first, before
// the warp waits for anything, then the exclusive prefix over all earlier tiles is resolved.
__device__ __forceinline__ void tile_post(unsigned long long* D, int tile, int total, int lane) {
    if (lane == 0) st_relaxed64(D + tile, (tile == 0 ? FU_INCL : FU_AGG) | (unsigned)total);
}
__device__ __forceinline__ int tile_lookback(unsigned long long* D, int tile, int total, int lane, int* err) {
    if (tile == 0) return 0;
    int excl = 0, base = tile - 1, spins = 0;
    while (true) {
        const int idx = base - lane;
        unsigned long long d = idx >= 0 ? ld_relaxed64(D + idx) : FU_INCL;
        if (__ballot_sync(FULL, (d >> 32) == 0ull)) {       // a predecessor has not posted yet
            if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
            __nanosleep(32);
            continue;
        }
        const unsigned incl = __ballot_sync(FULL, (d >> 32) == 2ull);
        int v = (int)(uint32_t)d;
        if (incl) {
            const int first = __ffs(incl) - 1;              // nearest predecessor that knows its inclusive prefix
            if (lane > first) v = 0;
            excl += warp_sum_int(v);
            break;
        }
        excl += warp_sum_int(v);
        base -= 32;
    }
    if (lane == 0) st_relaxed64(D + tile, FU_INCL | (unsigned)(excl + total));
    return excl;
}

// ---- shared memory behind the row slots: the hand-over rings between the three kinds of warps of a CTA
//   tile warps --(tile, kept / merged masks)--> scan warp --(tile, prefix, kept mask)--> aux workers
constexpr unsigned long long FU_ITEM_TILE = 1ull << 62, FU_ITEM_AUX = 2ull << 62, FU_ITEM_EXIT = 3ull << 62;

struct FusedQueue {                                         // one producer lane, several consumer warps
    unsigned long long item[FU_QSIZE];
    unsigned seq[FU_QSIZE];                                 // item[i] of lap n is valid once seq[i] == n + 1
    unsigned tail, head, freed, pad;
};

struct FusedShared {
    unsigned long long bars[FU_WARPS];                      // one mbarrier per tile warp
    int flags[2][FU_WARPS];                                 // per iteration parity: 0 kept, 1 merged away, 2 no row
    int next_tile[2], next_iter[2];                         // the ticket of the next tile, valid once next_iter == iteration + 1
    unsigned long long scan_item[FU_SCANQ];                 // tile warps (thread 0) -> scan warp
    unsigned scan_tail, scan_head;
    FusedQueue auxq;                                        // scan warp -> aux workers
};

__device__ __forceinline__ void queue_push(FusedQueue* q, unsigned long long item, int* err) {     // one lane
    const unsigned t = atomicAdd(&q->tail, 1u);
    int spins = 0;
    while ((int)(t - *(volatile unsigned*)&q->freed) >= FU_QSIZE) {                  // ring full: the workers are behind
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(100);
    }
    *(volatile unsigned long long*)&q->item[t % FU_QSIZE] = item;
    __threadfence_block();
    *(volatile unsigned*)&q->seq[t % FU_QSIZE] = t / FU_QSIZE + 1;
}

__device__ __forceinline__ unsigned long long queue_pop(FusedQueue* q, int lane) {             // whole warp
    unsigned h = 0;
    if (lane == 0) h = atomicAdd(&q->head, 1u);
    h = __shfl_sync(FULL, h, 0);
    const unsigned want = h / FU_QSIZE + 1;
    while (*(volatile unsigned*)&q->seq[h % FU_QSIZE] != want) __nanosleep(200);    // the producer always ends with EXIT items
    __threadfence_block();
    const unsigned long long item = *(volatile unsigned long long*)&q->item[h % FU_QSIZE];
    __syncwarp();
    if (lane == 0) atomicAdd(&q->freed, 1u);
    return item;
}

__device__ __forceinline__ void tile_barrier(int n_threads) {                       // the tile warps only (named barrier 1)
    asm volatile("bar.sync 1, %0;" :: "r"(n_threads) : "memory");
}

// CTA = W tile warps (W rows per tile, two shared-memory slots each) + one scan warp + FU_WORKERS - 1 aux workers.
template <int DT>
__global__ void __launch_bounds__((FU_WARPS + FU_WORKERS) * 32, 2)
k_fused_merge(const __grid_constant__ FusedArgs a, const __grid_constant__ AuxPack aux) {
    extern __shared__ __align__(128) unsigned char fu_smem[];
    pdl_enter();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, W = a.tile_rows;
    FusedShared* sh = reinterpret_cast<FusedShared*>(fu_smem + (size_t)(2 * W) * a.slot_bytes);
    unsigned long long* D = a.desc + 1;
    if (wid < W && lane == 0) mbar_init(smem_u32(&sh->bars[wid]), 1);
    if (threadIdx.x == 0) {
        sh->next_tile[0] = (int)atomicAdd(a.desc, 1ull);
        sh->next_iter[0] = sh->next_iter[1] = 0;
        sh->scan_tail = sh->scan_head = 0;
        sh->auxq.tail = sh->auxq.head = sh->auxq.freed = 0;
    }
    for (int i = threadIdx.x; i < FU_QSIZE; i += blockDim.x) sh->auxq.seq[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    int err = 0;
    const int nvec = a.nvec;
    const int64_t row_bytes = a.row_bytes;

    if (wid > W) {
        // ---- aux workers: the aux rows of a tile, off everybody's critical path
        while (true) {
            const unsigned long long item = queue_pop(&sh->auxq, lane);
            if ((item & (3ull << 62)) == FU_ITEM_EXIT) break;
            const int tile = (int)((item >> 38) & 0xffffffull), excl = (int)((item >> 8) & 0x3fffffffull);
            const unsigned mask = (unsigned)(item & 0xffull);
#pragma unroll 1
            for (int w = 0; w < W; ++w)
                if (mask >> w & 1u) gather_aux_rows(aux, tile * W + w, excl + __popc(mask & ((1u << w) - 1u)), lane);
        }
        return;
    }

    if (wid == W) {
        // ---- scan warp: destination rows.  For every tile of this CTA, in order: resolve the exclusive prefix (the tile's
        // count was posted by the tile warps), publish the state words of the kept rows, write the links of the next call.
        const int n_aux_workers = (int)(blockDim.x >> 5) - W - 1;
        unsigned head = 0;
        while (true) {
            int spins = 0;
            while (*(volatile unsigned*)&sh->scan_tail == head) {
                if (++spins > (FU_SPIN_LIMIT << 4)) { err = 1; break; }
                __nanosleep(100);
            }
            __threadfence_block();
            const unsigned long long item = *(volatile unsigned long long*)&sh->scan_item[head % FU_SCANQ];
            __syncwarp();
            ++head;
            if (lane == 0) *(volatile unsigned*)&sh->scan_head = head;
            if (err || (item & (3ull << 62)) == FU_ITEM_EXIT) break;
            const int tile = (int)((item >> 16) & 0xffffffull);
            const unsigned kept = (unsigned)(item >> 8) & 0xffu, merged = (unsigned)item & 0xffu;
            const int total = __popc(kept);
            const int excl = tile_lookback(D, tile, total, lane, &err);
            const int r = tile * W + lane;
            const bool mine = lane < W && (kept >> lane & 1u);
            const int d = excl + __popc(kept & ((1u << lane) - 1u));
            if (mine) {
                st_relaxed64(a.fstate + r, state_kept(d));
                a.dst[r] = d;
            } else if (lane < W && (merged >> lane & 1u)) {
                a.dst[r] = -1;
            }
            if (lane == 0) {
                if (aux.n && kept)
                    queue_push(&sh->auxq, FU_ITEM_AUX | ((unsigned long long)tile << 38) | ((unsigned long long)excl << 8) | kept, &err);
                if (tile == a.ntiles - 1) {
                    // the sequence is done: sizes, the speculated branch, the counters of the next call (main.py:112-120)
                    const long long s_keep = excl + total, n_merged = a.S - s_keep;
                    const long long N = a.counters[C_N], n_vis = a.counters[C_NVIS];
                    int e = 0;
                    if (n_vis == 0) e = 1;                  // the reference divides by zero here (main.py:114)
                    else if (!((double)n_merged / (double)n_vis < a.bound)) e = 3;   // top-k branch: the host redoes the call
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
                    a.status[FF_ST_ERROR] = e;
                    a.status[FF_ST_NMERGED] = n_merged;
                    a.status[FF_ST_FUSED] = 1;
                }
            }
            if (mine) {
                // links of the next call: a kept row follows the anchor of its predecessor's run
                const int2 lk = __ldg(a.link + r);
                if (lk.x >= 0) {
                    const unsigned long long st = wait_state(a.fstate, lk.x, &err);
                    if (state_type(st) != 0) {
                        a.link_next[d].x = state_dst(st);
                        a.link_next[state_dst(st)].y = d;
                    }
                } else {
                    a.link_next[d].x = lk.x;                // chain head / not a chain row
                    if (lk.x == -2) a.link_next[d].y = -2;
                }
                if (lk.x != -2 && lk.y < 0) a.link_next[d].y = -1;          // chain tail
            }
            __syncwarp();
        }
        if (lane == 0)
            for (int i = 0; i < n_aux_workers; ++i) queue_push(&sh->auxq, FU_ITEM_EXIT, &err);
        if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
        return;
    }

    // ---- tile warps.  Per warp: slot P (chain predecessor) and slot C (own row), one mbarrier for both
    unsigned char* slot_p = fu_smem + (size_t)(2 * wid) * a.slot_bytes;
    unsigned char* slot_c = slot_p + a.slot_bytes;
    const uint32_t sp32 = smem_u32(slot_p), sc32 = smem_u32(slot_c), bar = smem_u32(&sh->bars[wid]);
    const uint4* pr = reinterpret_cast<const uint4*>(slot_p);
    const uint4* cr = reinterpret_cast<const uint4*>(slot_c);
    int tile = sh->next_tile[0];
    int iter = 0;
    uint32_t phase = 0;
    bool store_pending = false;                             // lane 0: a bulk store may still be reading a slot

    while (tile < a.ntiles) {
        const int par = iter & 1;
        const int r = tile * W + wid;
        const bool valid = r < a.S;
        int2 lk = make_int2(-2, -2);
        if (valid) lk = __ldg(a.link + r);
        const int p = lk.x, sc = lk.y;
        const bool has_pred = valid && p >= 0;
        const bool self_emit = valid && (p == -2 || sc < 0);   // not a chain row, or a chain tail: it writes its own run
        const bool need_row = has_pred || self_emit;
        __syncwarp();                                       // every lane is done with the slots before they are refilled
        if (need_row && lane == 0) {
            if (store_pending) { tma_wait_read_0(); store_pending = false; }
            mbar_expect_tx(bar, (uint32_t)row_bytes * (has_pred ? 2u : 1u));
            tma_load(sc32, a.hidden + (int64_t)r * row_bytes, (uint32_t)row_bytes, bar);
            if (has_pred) tma_load(sp32, a.hidden + (int64_t)p * row_bytes, (uint32_t)row_bytes, bar);
        }
        // the predecessor's state word: its tile is ~P / W tiles back; usually published by the time the rows are here
        unsigned long long st_p = 0;
        if (has_pred) st_p = ld_relaxed64(a.fstate + p);
        if (valid) prefetch_aux(aux, r, lane);
        float s = -2.0f;                                    // IGNORE_TOKEN at chain heads (main.py:225-238)
        int flag = 0;
        if (need_row) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        if (has_pred) {
            float dot = 0.f, na = 0.f, nb = 0.f;
#pragma unroll 4
            for (int vb = 0; vb < nvec; vb += 32)           // lane l sums vectors l, l + 32, ... in this order (as k_similarity)
                if (vb + lane < nvec) acc_pair<DT>(pr[vb + lane], cr[vb + lane], dot, na, nb);
            dot = warp_sum(dot);
            na = warp_sum(na);
            nb = warp_sum(nb);
            s = finish_cosine<DT>(dot, na, nb);
            flag = (s >= a.thr);                            // NaN compares false
        }
        if (lane == 0) {
            if (valid) a.sim_seq[r] = s;
            *(volatile int*)&sh->flags[par][wid] = valid ? flag : 2;
        }
        tile_barrier(W * 32);                               // (A) the tile's flags are in shared memory
        if (threadIdx.x == 0) {
            // the tile's count goes out at once (nobody's look-back waits for more than this tile's rows), the rest of the
            // bookkeeping to the scan warp; the next tile's ticket travels while this tile is finished
            unsigned kept = 0, merged = 0;
#pragma unroll
            for (int w = 0; w < FU_WARPS; ++w)
                if (w < W) {
                    const int f = *(volatile int*)&sh->flags[par][w];
                    kept |= (unsigned)(f == 0) << w;
                    merged |= (unsigned)(f == 1) << w;
                }
            tile_post(D, tile, __popc(kept), 0);
            const int nt = (int)atomicAdd(a.desc, 1ull);
            const unsigned t = sh->scan_tail;
            int spins = 0;
            while ((int)(t - *(volatile unsigned*)&sh->scan_head) >= FU_SCANQ) {
                if (++spins > FU_SPIN_LIMIT) { err = 1; break; }
                __nanosleep(100);
            }
            *(volatile unsigned long long*)&sh->scan_item[t % FU_SCANQ] = FU_ITEM_TILE | ((unsigned long long)tile << 16) | (kept << 8) | merged;
            __threadfence_block();
            *(volatile unsigned*)&sh->scan_tail = t + 1;
            *(volatile int*)&sh->next_tile[par] = nt;
            __threadfence_block();
            *(volatile int*)&sh->next_iter[par] = iter + 1;
        }
        // ---- finish the predecessor's run.  Needs the predecessor's state word, nothing of this tile's own prefix.
        if (has_pred && state_type(st_p) == 0) st_p = wait_state(a.fstate, p, &err);
        const bool p_merged = has_pred && state_type(st_p) == 1;
        if (p_merged) {
            // the predecessor is inside a run: its raw row has served (the similarity); fetch the running sum of the run
            // — the anchor's destination row, written by the predecessor's warp before it published — into slot P
            if (lane == 0) {
                __threadfence();                            // acquire: the state word was read with a relaxed load
                asm volatile("fence.proxy.async;" ::: "memory");   // ... and the row is fetched through the async proxy
                mbar_expect_tx(bar, (uint32_t)row_bytes);
                tma_load(sp32, a.out + (int64_t)state_dst(st_p) * row_bytes, (uint32_t)row_bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        if (has_pred && state_type(st_p) != 0) {
            const int d_a = state_dst(st_p);                // destination row of the run's anchor (the predecessor itself if kept)
            const int L_p = p_merged ? state_len(st_p) : 0;
            char* orow = a.out + (int64_t)d_a * row_bytes;
            if (!flag) {
                // this row ends the run of its predecessor
                if (!p_merged) {                            // a plain kept row: the staged copy goes out as it is
                    if (lane == 0) {
                        tma_store(orow, sp32, (uint32_t)row_bytes);
                        tma_commit();
                        store_pending = true;
                    }
                } else {                                    // T(sum / T(L + 1)), main.py:314-317
                    const Divider<DT> dv(L_p + 1);
#pragma unroll 2
                    for (int vb = 0; vb < nvec; vb += 32)
                        if (vb + lane < nvec) st_stream16(orow + (int64_t)(vb + lane) * 16, dv.vec_fast(pr[vb + lane]));
                }
            } else {
                // this row is merged away: T(sum + row), main.py:304-311; slot P holds the sum so far (the anchor's raw row
                // if this is the first member)
                const int L = L_p + 1;
                if (sc < 0) {                               // ... and the chain ends here: finish the run as well
                    const Divider<DT> dv(L + 1);
#pragma unroll 2
                    for (int vb = 0; vb < nvec; vb += 32)
                        if (vb + lane < nvec)
                            st_stream16(orow + (int64_t)(vb + lane) * 16, dv.vec_fast(Num<DT>::add_vec(pr[vb + lane], cr[vb + lane])));
                    if (lane == 0) a.link_next[d_a].y = -1;
                } else {
#pragma unroll 2
                    for (int vb = 0; vb < nvec; vb += 32)
                        if (vb + lane < nvec)
                            st_stream16(orow + (int64_t)(vb + lane) * 16, Num<DT>::add_vec(pr[vb + lane], cr[vb + lane]));
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence();                        // the sum is visible before the state word that announces it
                    st_relaxed64(a.fstate + r, state_merged(d_a, L));
                }
            }
        }
        if (self_emit && !flag) {
            // an unmerged row nobody comes to close (chain tail, row outside the chains): it writes itself, which takes its
            // own destination — the one wait for this tile's own prefix, on the few rows of this kind
            const unsigned long long st_r = wait_state(a.fstate, r, &err);
            if (lane == 0 && state_type(st_r) == 2) {
                tma_store(a.out + (int64_t)state_dst(st_r) * row_bytes, sc32, (uint32_t)row_bytes);
                tma_commit();
                store_pending = true;
            }
        }
        // the next tile
        // (thread 0 always gets there: its own waits are bounded.  All tile warps must see the same tile: no early exit.)
        while (*(volatile int*)&sh->next_iter[par] != iter + 1) __nanosleep(20);
        __threadfence_block();
        tile = *(volatile int*)&sh->next_tile[par];
        ++iter;
    }

    tile_barrier(W * 32);
    if (threadIdx.x == 0) {
        const unsigned t = sh->scan_tail;
        while ((int)(t - *(volatile unsigned*)&sh->scan_head) >= FU_SCANQ) __nanosleep(100);
        *(volatile unsigned long long*)&sh->scan_item[t % FU_SCANQ] = FU_ITEM_EXIT;
        __threadfence_block();
        *(volatile unsigned*)&sh->scan_tail = t + 1;
    }
    // leave the other bank's state words and descriptors zeroed for the next call of the prefill
    const int64_t n_thr = (int64_t)gridDim.x * W * 32, me = (int64_t)blockIdx.x * W * 32 + threadIdx.x;
    for (int64_t i = me; i < a.S; i += n_thr) a.fstate_clr[i] = 0ull;
    for (int64_t i = me; i <= a.ntiles; i += n_thr) a.desc_clr[i] = 0ull;
    if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
    if (lane == 0) tma_wait_all();
}

// (pred, succ) of every sequence row from the compact by-patch arrays (ff_links.cuh / the scan kernels)
__global__ void __launch_bounds__(256)
k_links_seq(const int* __restrict__ rank, const int* __restrict__ order, const int* __restrict__ chain,
            const int64_t* __restrict__ counters, int S, int2* __restrict__ link) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const int N = (int)counters[C_N];
    const int j = rank[i];
    int2 l = make_int2(-2, -2);
    if (j >= 0) {
        const int c = chain[j];
        l.x = (j > 0 && chain[j - 1] == c) ? order[j - 1] : -1;
        l.y = (j + 1 < N && chain[j + 1] == c) ? order[j + 1] : -1;
    }
    link[i] = l;
}

}  // namespace ff
