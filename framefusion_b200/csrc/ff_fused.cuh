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

// Development aid (tools/trace_fused.py builds a separate library with -DFF_FUSED_TRACE): globaltimer stamps per tile.
#ifdef FF_FUSED_TRACE
#define FU_TRACE_SLOTS 16
#define FU_STAMP(tile, k) do { if (a.trace) a.trace[(size_t)(tile) * FU_TRACE_SLOTS + (k)] = fu_gtime(); } while (0)
#define FU_STAMP_MAX(tile, k) do { if (a.trace) atomicMax((unsigned long long*)&a.trace[(size_t)(tile) * FU_TRACE_SLOTS + (k)], (unsigned long long)fu_gtime()); } while (0)
__device__ __forceinline__ long long fu_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#else
#define FU_STAMP(tile, k) do { } while (0)
#define FU_STAMP_MAX(tile, k) do { } while (0)
#endif

struct FusedArgs {
    long long* trace;                              // FF_FUSED_TRACE builds only
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

// the same with a fixed number of polls: the caller has something else to do if the word is still missing
__device__ __forceinline__ unsigned long long poll_state(const unsigned long long* fstate, int x, int polls) {
    unsigned long long v = ld_relaxed64(fstate + x);
    for (int i = 0; i < polls && (v & 3ull) == 0ull; ++i) {
        __nanosleep(64);
        v = ld_relaxed64(fstate + x);
    }
    return v;
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
constexpr unsigned long long FU_ITEM_TILE = 1ull << 62, FU_ITEM_EXIT = 3ull << 62;
// worker items: bits [63:62] kind
//   POST  (tile << 38) | (exclusive prefix << 8) | kept mask: aux rows and next-call links of the tile's kept rows
//   COPY  (row << 31) | destination: a kept row that writes itself (chain tail, row outside the chains)
//   LINK  (row << 31) | destination: the next-call link of one kept row whose predecessor's state was late
constexpr unsigned long long FU_W_POST = 0ull << 62, FU_W_COPY = 2ull << 62, FU_W_LINK = 3ull << 62;

struct FusedQueue {                                         // several producer lanes, several consumer warps
    unsigned long long item[FU_QSIZE];
    unsigned seq[FU_QSIZE];                                 // slot i: 2n = free for lap n, 2n + 1 = holds the item of lap n
    unsigned tail, head, done;
    int pending;                                            // items pushed and not yet completed
};

struct FusedShared {
    unsigned long long bars[2 * FU_WARPS];                  // two mbarriers per tile warp: row loads, running-sum loads
    int flags[2][FU_WARPS];                                 // per iteration parity: 0 kept, 1 merged away, 2 no row
    int next_tile[2], next_iter[2];                         // the ticket of the next tile, valid once next_iter == iteration + 1
    unsigned long long scan_item[FU_SCANQ];                 // tile warps (thread 0) -> scan warp
    unsigned scan_tail, scan_head;
    FusedQueue auxq;                                        // tile warps / scan warp -> workers
};

// One lane.  Never waits for room: false if the ring is full, and the caller does the work itself — so nobody ever waits
// for something only a later row's progress could provide.
__device__ __forceinline__ bool queue_push(FusedQueue* q, unsigned long long item) {
    unsigned t;
    while (true) {
        t = *(volatile unsigned*)&q->tail;
        if ((int)(t - *(volatile unsigned*)&q->head) >= FU_QSIZE) return false;
        if (atomicCAS(&q->tail, t, t + 1u) == t) break;
    }
    atomicAdd(&q->pending, 1);
    const unsigned want = 2u * (t / FU_QSIZE);
    while (*(volatile unsigned*)&q->seq[t % FU_QSIZE] != want) __nanosleep(20);    // its last reader is just leaving
    *(volatile unsigned long long*)&q->item[t % FU_QSIZE] = item;
    __threadfence_block();
    *(volatile unsigned*)&q->seq[t % FU_QSIZE] = want + 1u;
    return true;
}

// whole warp.  false: nothing left to do (the producers are done and every item has been completed)
__device__ __forceinline__ bool queue_pop(FusedQueue* q, int lane, unsigned long long* out) {
    unsigned h = 0;
    int got = 0;
    if (lane == 0) {
        while (true) {
            h = *(volatile unsigned*)&q->head;
            if (h != *(volatile unsigned*)&q->tail) {
                if (atomicCAS(&q->head, h, h + 1u) == h) { got = 1; break; }
                continue;
            }
            if (*(volatile unsigned*)&q->done && *(volatile int*)&q->pending == 0) break;
            __nanosleep(200);
        }
    }
    got = __shfl_sync(FULL, got, 0);
    if (!got) return false;
    h = __shfl_sync(FULL, h, 0);
    const unsigned want = 2u * (h / FU_QSIZE) + 1u;
    while (*(volatile unsigned*)&q->seq[h % FU_QSIZE] != want) __nanosleep(50);     // claimed by its producer, being written
    __threadfence_block();
    *out = *(volatile unsigned long long*)&q->item[h % FU_QSIZE];
    __syncwarp();
    if (lane == 0) *(volatile unsigned*)&q->seq[h % FU_QSIZE] = want + 1u;          // free for the next lap
    return true;
}

__device__ __forceinline__ void queue_complete(FusedQueue* q, int lane) {
    __syncwarp();
    if (lane == 0) atomicSub(&q->pending, 1);
}

// A worker item of kind POST (aux rows + next-call links of a tile's kept rows; lane w: row w), LINK (the link of one
// row, lane 0) or COPY (a kept row that writes itself).  requeue: a link whose predecessor state is missing goes back to
// the ring if there is room (workers); otherwise the lane waits for it (an earlier row: it always comes).
template <int DT>
__device__ __noinline__ void run_post_item(const FusedArgs& a, const AuxPack& aux, FusedQueue* q, int W, unsigned long long item,
                                           int lane, int* err, bool requeue) {
    const unsigned long long kind = item & (3ull << 62);
    if (kind == FU_W_COPY) {
        const int r = (int)((item >> 31) & 0x3fffffffull), d = (int)(item & 0x7fffffffull);
        copy_row(a.hidden + (int64_t)r * a.row_bytes, a.out + (int64_t)d * a.row_bytes, a.nvec, lane);
        return;
    }
    int r = -1, d = -1;
    if (kind == FU_W_POST) {
        const int tile = (int)((item >> 38) & 0xffffffull), excl = (int)((item >> 8) & 0x3fffffffull);
        const unsigned mask = (unsigned)(item & 0xffull);
        if (aux.n) {
#pragma unroll 1
            for (int w = 0; w < W; ++w)
                if (mask >> w & 1u) gather_aux_rows(aux, tile * W + w, excl + __popc(mask & ((1u << w) - 1u)), lane);
        }
        if (lane < W && (mask >> lane & 1u)) { r = tile * W + lane; d = excl + __popc(mask & ((1u << lane) - 1u)); }
    } else if (lane == 0) {
        r = (int)((item >> 31) & 0x3fffffffull);
        d = (int)(item & 0x7fffffffull);
    }
    // links of the next call: a kept row follows the anchor of its predecessor's run
    if (r >= 0) {
        const int2 lk = __ldg(a.link + r);
        if (kind == FU_W_POST) {
            if (lk.x < 0) {
                a.link_next[d].x = lk.x;                    // chain head / not a chain row
                if (lk.x == -2) a.link_next[d].y = -2;
            }
            if (lk.x != -2 && lk.y < 0) a.link_next[d].y = -1;              // chain tail
        }
        if (lk.x >= 0) {
            unsigned long long st = poll_state(a.fstate, lk.x, 24);
            if (state_type(st) == 0 && !(requeue && queue_push(q, FU_W_LINK | ((unsigned long long)r << 31) | (unsigned long long)d)))
                st = wait_state(a.fstate, lk.x, err);
            if (state_type(st) != 0) {
                a.link_next[d].x = state_dst(st);
                a.link_next[state_dst(st)].y = d;
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void tile_barrier(int n_threads) {                       // the tile warps only (named barrier 1)
    asm volatile("bar.sync 1, %0;" :: "r"(n_threads) : "memory");
}

// CTA = W tile warps (W rows per tile, four shared-memory slots each) + one scan warp + FU_WORKERS - 1 workers.
template <int DT>
__global__ void __launch_bounds__((FU_WARPS + FU_WORKERS) * 32, 2)
k_fused_merge(const __grid_constant__ FusedArgs a, const __grid_constant__ AuxPack aux) {
    extern __shared__ __align__(128) unsigned char fu_smem[];
    pdl_enter();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, W = a.tile_rows;
    FusedShared* sh = reinterpret_cast<FusedShared*>(fu_smem + (size_t)(4 * W) * a.slot_bytes);
    unsigned long long* D = a.desc + 1;
    if (wid < W && lane == 0) {
        mbar_init(smem_u32(&sh->bars[2 * wid]), 1);
        mbar_init(smem_u32(&sh->bars[2 * wid + 1]), 1);
    }
    if (threadIdx.x == 0) {
        sh->next_tile[0] = (int)atomicAdd(a.desc, 1ull);
        sh->next_iter[0] = sh->next_iter[1] = 0;
        sh->scan_tail = sh->scan_head = 0;
        sh->auxq.tail = sh->auxq.head = sh->auxq.done = 0;
        sh->auxq.pending = 0;
    }
    for (int i = threadIdx.x; i < FU_QSIZE; i += blockDim.x) sh->auxq.seq[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    int err = 0;
    const int nvec = a.nvec;
    const int64_t row_bytes = a.row_bytes;

    if (wid > W) {
        // ---- workers: what nobody waits for — aux rows, next-call links, rows that write themselves.  A link whose
        // predecessor state is missing goes back to the end of the ring while there is room.
        FusedQueue* q = &sh->auxq;
        unsigned long long item;
        while (queue_pop(q, lane, &item)) {
            run_post_item<DT>(a, aux, q, W, item, lane, &err, true);
            queue_complete(q, lane);
        }
        if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
        return;
    }

    if (wid == W) {
        // ---- scan warp: destination rows.  For every tile of this CTA, in order: resolve the exclusive prefix (the tile's
        // count was posted by the tile warps), publish the state words of the kept rows, write the links of the next call.
        unsigned head = 0;
        while (true) {
            int spins = 0;
            bool dead = false;
            while (*(volatile unsigned*)&sh->scan_tail == head) {
                if (++spins > (FU_SPIN_LIMIT << 6)) { dead = true; break; }     // seconds without a tile: the tile warps are gone
                __nanosleep(100);
            }
            if (dead) { err = 1; break; }
            __threadfence_block();
            const unsigned long long item = *(volatile unsigned long long*)&sh->scan_item[head % FU_SCANQ];
            __syncwarp();
            ++head;
            if (lane == 0) *(volatile unsigned*)&sh->scan_head = head;
            if ((item & (3ull << 62)) == FU_ITEM_EXIT) break;
            const int tile = (int)((item >> 16) & 0xffffffull);
            const unsigned kept = (unsigned)(item >> 8) & 0xffu, merged = (unsigned)item & 0xffu;
            const int total = __popc(kept);
            if (lane == 0) FU_STAMP(tile, 6);
            const int excl = tile_lookback(D, tile, total, lane, &err);
            if (lane == 0) FU_STAMP(tile, 7);
            const int r = tile * W + lane;
            const bool mine = lane < W && (kept >> lane & 1u);
            const int d = excl + __popc(kept & ((1u << lane) - 1u));
            if (mine) {
                st_relaxed64(a.fstate + r, state_kept(d));
                a.dst[r] = d;
            } else if (lane < W && (merged >> lane & 1u)) {
                a.dst[r] = -1;
            }
            // rows that write themselves (chain tails, rows outside the chains): their destination is known here
            unsigned self = 0;
            if (mine) {
                const int2 lk = __ldg(a.link + r);
                self = (lk.x == -2 || lk.y < 0) ? 1u : 0u;
            }
            unsigned self_mask = __ballot_sync(FULL, self != 0u);
            // hand the rest of the tile's bookkeeping to the workers; with the ring full this warp does it itself
            while (self_mask) {
                const int w = __ffs(self_mask) - 1;
                self_mask &= self_mask - 1;
                const unsigned long long it = FU_W_COPY | ((unsigned long long)(tile * W + w) << 31) |
                                              (unsigned long long)(excl + __popc(kept & ((1u << w) - 1u)));
                int ok = 0;
                if (lane == 0) ok = queue_push(&sh->auxq, it) ? 1 : 0;
                if (!__shfl_sync(FULL, ok, 0)) run_post_item<DT>(a, aux, &sh->auxq, W, it, lane, &err, false);
            }
            if (kept) {
                const unsigned long long it = FU_W_POST | ((unsigned long long)tile << 38) | ((unsigned long long)excl << 8) | kept;
                int ok = 0;
                if (lane == 0) ok = queue_push(&sh->auxq, it) ? 1 : 0;
                if (!__shfl_sync(FULL, ok, 0)) run_post_item<DT>(a, aux, &sh->auxq, W, it, lane, &err, false);
            }
            if (lane == 0) {
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
            __syncwarp();
            if (lane == 0) FU_STAMP(tile, 8);
        }
        if (lane == 0) {
            __threadfence_block();
            *(volatile unsigned*)&sh->auxq.done = 1u;       // every push of the tile warps came before their EXIT item
        }
        if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
        return;
    }

    // ---- tile warps.  Per warp: two slot sets {P: chain predecessor, C: own row}, used alternately, so that the emission
    // step of tile k - 1 runs while the rows of tile k travel — by then the predecessor's state word is an iteration old.
    unsigned char* set0 = fu_smem + (size_t)(4 * wid) * a.slot_bytes;
    const uint32_t barL = smem_u32(&sh->bars[2 * wid]), barS = smem_u32(&sh->bars[2 * wid + 1]);
    int tile = sh->next_tile[0];
    int iter = 0;
    uint32_t phaseL = 0, phaseS = 0;
    bool store_pending = false;                             // lane 0: a bulk store may still be reading a slot
    // the row of the previous tile whose emission step is still to do
    bool prev_todo = false;
    int prev_r = 0, prev_p = 0, prev_flag = 0, prev_sc = 0, prev_tile = 0;
    unsigned long long prev_st = 0;

    // The emission step of row prev_r out of slot set `ps`: its predecessor's run ends, or grows by this row.
    auto emit_prev = [&](unsigned char* ps) {
        const uint4* pr = reinterpret_cast<const uint4*>(ps);
        const uint4* cr = reinterpret_cast<const uint4*>(ps + a.slot_bytes);
        const uint32_t sp32 = smem_u32(ps);
        const bool p_merged = state_type(prev_st) == 1;
        if (p_merged) {
            // the predecessor is inside a run: its raw row has served (the similarity); fetch the running sum of the run — the
            // anchor's destination row, written before the state word was — into slot P
            if (lane == 0) {
                __threadfence();                            // acquire: the state word was read with a relaxed load
                asm volatile("fence.proxy.async;" ::: "memory");   // ... and the row is fetched through the async proxy
                mbar_expect_tx(barS, (uint32_t)row_bytes);
                tma_load(sp32, a.out + (int64_t)state_dst(prev_st) * row_bytes, (uint32_t)row_bytes, barS);
            }
            mbar_wait(barS, phaseS);
            phaseS ^= 1u;
        }
        const int d_a = state_dst(prev_st);                 // destination row of the run's anchor (the predecessor itself if kept)
        const int L_p = p_merged ? state_len(prev_st) : 0;
        char* orow = a.out + (int64_t)d_a * row_bytes;
        if (!prev_flag) {
            // this row ends the run of its predecessor
            if (!p_merged) {                                // a plain kept row: the staged copy goes out as it is
                if (lane == 0) {
                    tma_store(orow, sp32, (uint32_t)row_bytes);
                    tma_commit();
                    store_pending = true;
                }
            } else {                                        // T(sum / T(L + 1)), main.py:314-317
                const Divider<DT> dv(L_p + 1);
#pragma unroll 2
                for (int vb = 0; vb < nvec; vb += 32)
                    if (vb + lane < nvec) st_stream16(orow + (int64_t)(vb + lane) * 16, dv.vec_fast(pr[vb + lane]));
            }
        } else {
            // this row is merged away: T(sum + row), main.py:304-311; slot P holds the sum so far (the anchor's raw row if this
            // is the first member)
            const int L = L_p + 1;
            if (prev_sc < 0) {                              // ... and the chain ends here: finish the run as well
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
                __threadfence();                            // the sum is visible before the state word that announces it
                st_relaxed64(a.fstate + prev_r, state_merged(d_a, L));
            }
        }
        if (lane == 0) FU_STAMP_MAX(prev_tile, 11);
        prev_todo = false;
    };

    while (tile < a.ntiles || prev_todo) {
        const int par = iter & 1;
        const bool live = tile < a.ntiles;                  // false: only the last tile's emission step is left
        unsigned char* cs = set0 + (size_t)(2 * par) * a.slot_bytes;         // slot set of this iteration
        unsigned char* ps = set0 + (size_t)(2 * (par ^ 1)) * a.slot_bytes;   // ... of the previous one
        if (live && threadIdx.x == 0) FU_STAMP(tile, 0);
        // the next tile's ticket travels during this iteration
        int nt = 0;
        if (live && threadIdx.x == 0) nt = (int)atomicAdd(a.desc, 1ull);
        const int r = tile * W + wid;
        const bool valid = live && r < a.S;
        int2 lk = make_int2(-2, -2);
        if (valid) lk = __ldg(a.link + r);
        const int p = lk.x, sc = lk.y;
        const bool has_pred = valid && p >= 0;              // rows without a predecessor in their chain need no data here:
        __syncwarp();                                       // a successor or a worker writes them
        if (has_pred && lane == 0) {
            if (store_pending) { tma_wait_read_0(); store_pending = false; }
            mbar_expect_tx(barL, (uint32_t)row_bytes * 2u);
            tma_load(smem_u32(cs + a.slot_bytes), a.hidden + (int64_t)r * row_bytes, (uint32_t)row_bytes, barL);
            tma_load(smem_u32(cs), a.hidden + (int64_t)p * row_bytes, (uint32_t)row_bytes, barL);
        }
        if (valid) prefetch_aux(aux, r, lane);
        // while the rows travel: the previous tile's emission step, if its predecessor's state word is there (it nearly
        // always is: that row's tile posted an iteration ago at the latest)
        if (prev_todo) {
            if (state_type(prev_st) == 0) prev_st = ld_relaxed64(a.fstate + prev_p);
            if (lane == 0) FU_STAMP_MAX(prev_tile, 9);
            if (state_type(prev_st) != 0) emit_prev(ps);
        }
        float s = -2.0f;                                    // IGNORE_TOKEN at chain heads (main.py:225-238)
        int flag = 0;
        if (has_pred) {
            mbar_wait(barL, phaseL);
            phaseL ^= 1u;
        }
        if (live && threadIdx.x == 0) FU_STAMP(tile, 1);
        if (live && lane == 0) FU_STAMP_MAX(tile, 10);
        if (has_pred) {
            const uint4* pr = reinterpret_cast<const uint4*>(cs);
            const uint4* cr = reinterpret_cast<const uint4*>(cs + a.slot_bytes);
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
        if (live) {
            if (lane == 0) {
                if (valid) a.sim_seq[r] = s;
                *(volatile int*)&sh->flags[par][wid] = valid ? flag : 2;
            }
            tile_barrier(W * 32);                           // (A) the tile's flags are in shared memory
            if (threadIdx.x == 0) {
                FU_STAMP(tile, 2);
                // the tile's count goes out at once: nobody's look-back ever waits for more than the rows of a tile to arrive
                unsigned kept = 0, merged = 0;
#pragma unroll
                for (int w = 0; w < FU_WARPS; ++w)
                    if (w < W) {
                        const int f = *(volatile int*)&sh->flags[par][w];
                        kept |= (unsigned)(f == 0) << w;
                        merged |= (unsigned)(f == 1) << w;
                    }
                tile_post(D, tile, __popc(kept), 0);
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
                FU_STAMP(tile, 3);
            }
        }
        // the previous tile's emission step if it could not be done above: now this warp waits (its count is out already)
        if (prev_todo) {
            prev_st = wait_state(a.fstate, prev_p, &err);
#ifdef FF_FUSED_TRACE
            if (lane == 0 && a.trace) atomicAdd((unsigned long long*)&a.trace[(size_t)prev_tile * FU_TRACE_SLOTS + 12], 1ull);
#endif
            if (state_type(prev_st) != 0) emit_prev(ps);
            prev_todo = false;
        }
        // this tile's row becomes the pending one
        if (has_pred) {
            prev_todo = true;
            prev_r = r; prev_p = p; prev_flag = flag; prev_sc = sc; prev_tile = tile;
            prev_st = ld_relaxed64(a.fstate + p);
        }
        if (live) {
            if (threadIdx.x == 0) FU_STAMP(tile, 5);
            // the next tile (thread 0 always gets there: its own waits are bounded.  All tile warps must see the same tile.)
            while (*(volatile int*)&sh->next_iter[par] != iter + 1) __nanosleep(20);
            __threadfence_block();
            tile = *(volatile int*)&sh->next_tile[par];
        }
        ++iter;
    }

    tile_barrier(W * 32);
    if (threadIdx.x == 0) {
        const unsigned t = sh->scan_tail;
        int spins = 0;
        while ((int)(t - *(volatile unsigned*)&sh->scan_head) >= FU_SCANQ) {
            if (++spins > FU_SPIN_LIMIT) { err = 1; break; }
            __nanosleep(100);
        }
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
