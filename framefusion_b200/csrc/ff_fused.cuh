// Read-once merge kernel: similarity + threshold select + run merge + compaction of hidden_states and of the aux
// tensors (cos / sin / patch_type / position ids) in ONE launch that fetches every row of hidden_states from HBM
// exactly once (main.py:104-138, threshold branch — every merge call of a prefill but possibly the last).
//
// Shape of the problem.  A token is compared with the previous surviving token of the SAME patch id (its chain
// predecessor, main.py:216-238) — 576 rows (4 MB) back on the first call of a uniform video — while the output is
// compacted in SEQUENCE order (main.py:132-138): a row's destination is the number of kept rows before it, known only
// when every earlier row has been compared.  With ~600 tiles in flight that knowledge is always several microseconds
// behind the rows, and an SM cannot park rows that long (latency x bandwidth of the loads alone fills its shared
// memory).  So the kernel is three feed-forward stages, and nothing upstream ever waits for anything downstream:
//
//   FRONT (tile warps).  Tiles of W consecutive rows are handed out in sequence order by a ticket (a dispatcher warp per
//     CTA takes them a few iterations ahead); warp w of the CTA owns row tile * W + w.  Its own row (HBM) and its chain predecessor (pred[r]: an L2 hit, that row was some tile's own row
//     a few microseconds ago) are staged by TMA (cp.async.bulk + mbarrier) into the warp's two shared-memory slots.
//     Three row sums out of shared memory, warp shuffles, the reference's rounding chain -> sim, and the row's flag
//     (kept / merged away) is published.  A front warp waits for its rows and for nothing else.
//   SCAN (one warp per CTA).  Collects the flags of each of the CTA's tiles, posts the tile's kept-row count, resolves
//     the exclusive prefix by decoupled look-back over the tile descriptors, publishes the destinations.
//   WORKERS.  Per tile, out of the L2: a kept row ends the run of its predecessor — the worker walks the flags back to
//     the run's anchor and writes the anchor's destination row: the raw row if the run has no members, else
//     T(T(..T(anchor + m1) + ..) + mL) / T(L+1) with one rounding to T per add in chain order (the sequence torch-CPU
//     index_add_ performs, main.py:304-311) and one division (main.py:314-317).  Chain tails end their own run.  The
//     aux rows (cos / sin / patch_type / position ids) and the links of the next call (pred / succ by destination
//     index) go with it.  Items are independent of each other.
//
// Every wait is for a row with a SMALLER sequence index, tickets are taken in order by CTAs that are running, and every
// spin is bounded: no deadlock whatever is resident, and a kernel that always ends.
//
// The branch decision (main.py:114-116) needs the global count, known only at the end: the kernel speculates on the
// threshold branch, the scan warp of the last tile checks count / n_vis < bound and otherwise reports FF_ST_ERROR = 3;
// the host then redoes the call with the multi-kernel path (top-k branch, at most once per prefill).  The input is
// never modified, so the redo sees the original rows.
#pragma once
#include "ff_common.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FU_WARPS = 8;                        // warps of the two front warpgroups: up to seven tile warps (rows per tile) and the dispatcher
constexpr int FU_WORKERS = 4;                      // warps per CTA besides the tile warps: one scan warp + workers (one warpgroup)
constexpr int FU_WSLOTS = 2 * (FU_WORKERS - 1);    // shared-memory row slots of the workers: two each
constexpr int FU_REGS_LAUNCH = 80, FU_REGS_FRONT = 56, FU_REGS_BACK = 128;   // setmaxnreg: 12 * 32 * 80 = 8 * 32 * 56 + 4 * 32 * 128
constexpr int FU_QSIZE = 32;                       // ring entries between the scan warp and the workers
constexpr int FU_SCANQ = 8;                        // ring entries between the tile warps and the scan warp
constexpr int FU_TICKETS = 4;                      // how many iterations the tile warps of a CTA may drift apart
constexpr int FU_SPIN_LIMIT = 1 << 18;             // polls (~64 ns apart) before a wait gives up and reports FF_ST_INTERNAL

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

// the aux tensors, one entry per (tensor, plane): rows of at most 512 bytes in 16- or 8-byte pieces, one piece per lane
struct AuxFlat {
    int n;                                         // entries; -1: the tensors do not fit this form (gather_aux_rows instead)
    int row_bytes[8];
    int piece[8];                                  // 16 or 8: bytes per lane
    const char* src[8];
    char* dst[8];
};

struct FusedArgs {
    long long* trace;                              // FF_FUSED_TRACE builds only
    AuxFlat auxf;
    const char* hidden;
    char* out;
    int S, nvec, row_bytes, slot_bytes, ntiles, tile_rows;
    const int2* link;                              // [S] (pred, succ): row index, -1 = chain head / tail, -2 = not a chain row
    int2* link_next;                               // [S_keep] the same for the compacted sequence
    unsigned* fflag;                               // [S] front flags, zero on entry
    unsigned* fdst;                                // [S] destination + 1, zero on entry
    unsigned long long* desc;                      // zero on entry, desc_words u64 in all: the ticket, round_base [nrounds + 1],
                                                   // then u32 slot [ntiles_pad], excl [ntiles_pad], round_cnt [nrounds]
    int desc_words, ntiles_pad, nrounds;
    unsigned* fflag_clr;                           // other bank: cleared for the next call
    unsigned* fdst_clr;
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

// 16-byte load that stays in the L2 (rows the front pulled in a few microseconds ago)
__device__ __forceinline__ uint4 ld_cg16(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ unsigned ld_relaxed32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// front flag of a row: 0 = not yet known, 1 = merged away, 2 = kept
// Waits are for rows with a smaller sequence index, and give up after FU_SPIN_LIMIT polls: *err is set, the caller skips
// what depended on the value, the kernel always terminates.
__device__ __forceinline__ unsigned wait_flag(const unsigned* fflag, int x, int* err) {
    unsigned v = ld_relaxed32(fflag + x);
    int spins = 0;
    while (v == 0u) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(64);
        v = ld_relaxed32(fflag + x);
    }
    return v;
}
__device__ __forceinline__ int wait_dst(const unsigned* fdst, int x, int* err) {       // destination row, -1 after a time-out
    unsigned v = ld_relaxed32(fdst + x);
    int spins = 0;
    while (v == 0u) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(64);
        v = ld_relaxed32(fdst + x);
    }
    return (int)v - 1;
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

// ---- the prefix over the tiles, without anybody polling shared descriptors (hundreds of tiles resolve together here: a
// decoupled look-back has every one of them poll the same few cache lines, and the L2 serves a line one request at a
// time).  Tiles are grouped in rounds of FU_ROUND by number.  A tile posts its kept mask into its own slot and bumps the
// round's counter; whoever brings the counter to the round's size scans the round — one warp, eight slots per lane — takes
// the round's base from the previous round's scanner through one word, and writes every tile's exclusive prefix into that
// tile's own word.  A tile only ever polls its own word.
constexpr int FU_ROUND = 256;

struct ScanArrays {
    unsigned* slot;              // [ntiles_pad] 0x100 | kept mask once posted
    unsigned* excl;              // [ntiles_pad] exclusive prefix + 1
    unsigned* round_cnt;         // [nrounds] tiles of the round that have posted
    unsigned long long* round_base;   // [nrounds + 1] kept rows before the round + 1
};

__device__ __forceinline__ void round_scan(const ScanArrays& sa, int round, int ntiles, int lane, int* err) {
    const int t0 = round * FU_ROUND + lane * 8;             // this lane's eight tiles (the arrays are padded to whole rounds)
    __threadfence();                                        // acquire: every slot of the round was written before its bump
    unsigned m[8];
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        m[j] = (t0 + j < ntiles) ? ld_relaxed32(sa.slot + t0 + j) : 0u;
        sum += __popc(m[j] & 0xffu);
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    unsigned long long base = 1ull;                         // round 0 starts at 0 (+ 1)
    if (round > 0) {
        int spins = 0;
        base = ld_relaxed64(sa.round_base + round);
        while (base == 0ull) {                              // the previous round's scanner is still at it
            if (++spins > FU_SPIN_LIMIT) { *err = 1; base = 1ull; break; }
            __nanosleep(100);
            base = ld_relaxed64(sa.round_base + round);
        }
    }
    if (lane == 0) st_relaxed64(sa.round_base + round + 1, base + (unsigned long long)total);
    unsigned e = (unsigned)base + (unsigned)(incl - sum);   // exclusive prefix + 1 of this lane's first tile
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (t0 + j < ntiles) st_relaxed32(sa.excl + t0 + j, e);
        e += __popc(m[j] & 0xffu);
    }
}

// ---- shared memory behind the row slots: the hand-over rings between the three kinds of warps of a CTA
//   tile warps --(tile)--> scan warp --(tile, prefix, kept mask)--> workers
constexpr unsigned long long FU_ITEM_EXIT = ~0ull;

struct FusedQueue {                                         // one producer (the scan warp), several consumer warps
    unsigned long long item[FU_QSIZE];
    unsigned seq[FU_QSIZE];                                 // slot i: 2n = free for lap n, 2n + 1 = holds the item of lap n
    unsigned tail, head, done, pad;
};

struct FusedShared {
    unsigned long long bars[FU_WARPS];                      // one mbarrier per tile warp
    unsigned long long wbars[FU_WSLOTS];                    // one per worker slot
    int next_tile[FU_TICKETS], next_iter[FU_TICKETS];       // the ticket of the next tile, valid once next_iter == iteration + 1
    int progress[FU_WARPS];                                 // iterations each tile warp has finished
    unsigned long long scan_item[FU_SCANQ];                 // tile warps (thread 0) -> scan warp: tile numbers
    unsigned scan_tail, scan_head;
    FusedQueue q;                                           // scan warp -> workers
};

// One lane.  Never waits for room: false if the ring is full, and the caller does the work itself.
__device__ __forceinline__ bool queue_push(FusedQueue* q, unsigned long long item) {
    const unsigned t = *(volatile unsigned*)&q->tail;       // single producer
    if ((int)(t - *(volatile unsigned*)&q->head) >= FU_QSIZE) return false;
    const unsigned want = 2u * (t / FU_QSIZE);
    if (*(volatile unsigned*)&q->seq[t % FU_QSIZE] != want) return false;           // its last reader is just leaving
    *(volatile unsigned long long*)&q->item[t % FU_QSIZE] = item;
    __threadfence_block();
    *(volatile unsigned*)&q->seq[t % FU_QSIZE] = want + 1u;
    *(volatile unsigned*)&q->tail = t + 1u;
    return true;
}

// whole warp.  false: nothing left to do (the producer is done and the ring is empty)
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
            if (*(volatile unsigned*)&q->done) {
                if (*(volatile unsigned*)&q->head == *(volatile unsigned*)&q->tail) break;
                continue;
            }
            __nanosleep(200);
        }
    }
    got = __shfl_sync(FULL, got, 0);
    if (!got) return false;
    h = __shfl_sync(FULL, h, 0);
    __threadfence_block();
    *out = *(volatile unsigned long long*)&q->item[h % FU_QSIZE];
    __syncwarp();
    if (lane == 0) *(volatile unsigned*)&q->seq[h % FU_QSIZE] = 2u * (h / FU_QSIZE) + 2u;   // free for the next lap
    return true;
}

// ---- the workers' side.  A worker owns two shared-memory row slots; rows travel L2 -> slot -> destination by TMA bulk
// copies (no registers, two rows in flight per worker), runs are summed in the slots.
struct WorkerSlots {
    unsigned char* ptr[2];
    uint32_t addr[2], bar[2], phase[2];
    int state[2];                                           // 0 free, 1 a row is arriving (to be stored to dst), 2 a store is reading it
    char* dst[2];
    int next;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// the row that is arriving in slot s goes on to its destination
__device__ __forceinline__ void slot_finish(WorkerSlots& ws, int s, int row_bytes, int lane) {
    if (ws.state[s] != 1) return;
    mbar_wait(ws.bar[s], ws.phase[s]);
    ws.phase[s] ^= 1u;
    if (lane == 0) {
        tma_store(ws.dst[s], ws.addr[s], (uint32_t)row_bytes);
        tma_commit();
    }
    ws.state[s] = 2;
}
// slot s can be overwritten
__device__ __forceinline__ void slot_free(WorkerSlots& ws, int s, int row_bytes, int lane) {
    slot_finish(ws, s, row_bytes, lane);
    if (ws.state[s] == 2) {
        if (lane == 0) tma_wait_read_0();                   // (every store this lane has issued: the other slot's too)
        __syncwarp();
        ws.state[0] = ws.state[0] == 2 ? 0 : ws.state[0];
        ws.state[1] = ws.state[1] == 2 ? 0 : ws.state[1];
    }
}
// src row -> dst row through the next slot; returns with the row (and possibly the one before it) still travelling
__device__ __forceinline__ void worker_copy(WorkerSlots& ws, const char* src, char* dst, int row_bytes, int lane) {
    const int s = ws.next;
    ws.next ^= 1;
    slot_free(ws, s, row_bytes, lane);
    if (lane == 0) {
        mbar_expect_tx(ws.bar[s], (uint32_t)row_bytes);
        tma_load(ws.addr[s], src, (uint32_t)row_bytes, ws.bar[s]);
    }
    ws.state[s] = 1;
    ws.dst[s] = dst;
    slot_finish(ws, s ^ 1, row_bytes, lane);                // meanwhile the previous row has arrived: send it on
}

// A run: anchor row and its L >= 1 members -> destination row d_a as T(T(..T(anchor + m1) ..+ mL) / T(L + 1)): one rounding to
// T per add in chain order (main.py:304-311), one division (main.py:314-317).  `last` = the last member; the members are
// visited front to back (lane k remembers the k-th from the end; runs longer than 32 follow the successor links).
template <int DT>
__device__ __forceinline__ void worker_run(const FusedArgs& a, WorkerSlots& ws, int anchor, int last, int L, int d_a, int lane) {
    const int row_bytes = a.row_bytes, nvec = a.nvec;
    slot_free(ws, 0, row_bytes, lane);
    slot_free(ws, 1, row_bytes, lane);
    int mine = -1;
    {
        int x = last;
        for (int k = 0; k < L && k < 32; ++k) {
            if (lane == k) mine = x;
            x = __ldg(&a.link[x].x);
        }
    }
    uint4* A = reinterpret_cast<uint4*>(ws.ptr[0]);
    const uint4* B = reinterpret_cast<const uint4*>(ws.ptr[1]);
    int walk = anchor;
#pragma unroll 1
    for (int m = L - 1; m >= 0; --m) {                      // m = L - 1: first member behind the anchor ... m = 0: the last
        int idx;
        if (L <= 32) idx = __shfl_sync(FULL, mine, m);
        else { walk = __ldg(&a.link[walk].y); idx = walk; }
        __syncwarp();                                       // every lane has read slot B
        if (lane == 0) {
            if (m == L - 1) {
                mbar_expect_tx(ws.bar[0], (uint32_t)row_bytes);
                tma_load(ws.addr[0], a.hidden + (int64_t)anchor * row_bytes, (uint32_t)row_bytes, ws.bar[0]);
            }
            mbar_expect_tx(ws.bar[1], (uint32_t)row_bytes);
            tma_load(ws.addr[1], a.hidden + (int64_t)idx * row_bytes, (uint32_t)row_bytes, ws.bar[1]);
        }
        if (m == L - 1) { mbar_wait(ws.bar[0], ws.phase[0]); ws.phase[0] ^= 1u; }
        mbar_wait(ws.bar[1], ws.phase[1]);
        ws.phase[1] ^= 1u;
        if (m > 0) {
#pragma unroll 2
            for (int v = lane; v < nvec; v += 32) A[v] = Num<DT>::add_vec(A[v], B[v]);       // T(acc + member)
        } else {
            const Divider<DT> dv(L + 1);
#pragma unroll 2
            for (int v = lane; v < nvec; v += 32) A[v] = dv.vec_fast(Num<DT>::add_vec(A[v], B[v]));
        }
    }
    fence_async_smem();                                     // the sums were written through the generic proxy
    __syncwarp();
    if (lane == 0) {
        tma_store(a.out + (int64_t)d_a * row_bytes, ws.addr[0], (uint32_t)row_bytes);
        tma_commit();
    }
    ws.state[0] = 2;
    ws.next = 1;
}

// the aux rows of sequence row r -> destination row d: one 16- or 8-byte piece per lane and entry, all loads first
__device__ __forceinline__ void worker_aux(const FusedArgs& a, const AuxPack& aux, int r, int d, int lane) {
    const AuxFlat& f = a.auxf;
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

// The runs that end in one tile, L2 -> destination, with the aux rows of its kept rows and the links of the next call.
// item = (tile << 38) | (exclusive prefix << 8) | kept mask.  Lane w looks after row w of the tile:
//   kept, with a predecessor: it ends the predecessor's run -> that run goes out
//   kept, nobody behind it (chain tail, row outside the chains): it goes out itself
//   merged away at the end of its chain: it ends its own run -> that run goes out
// Every flag the walks read belongs to an earlier row than the tile's last: known since the tile's look-back resolved.
template <int DT>
__device__ __forceinline__ void run_tile_item(const FusedArgs& a, const AuxPack& aux, WorkerSlots& ws, int W, unsigned long long item,
                                              int lane, int* err) {
    const int tile = (int)((item >> 38) & 0xffffffull), excl = (int)((item >> 8) & 0x3fffffffull);
    const unsigned kept = (unsigned)(item & 0xffull);
    const int r = tile * W + lane;
    int run_last = -1, run_anchor = -1, run_L = 0, run_dst = -1;    // the run this lane's row ends
    int self_dst = -1, d_r = -1;
    if (lane < W && r < a.S) {
        const int2 lk = __ldg(a.link + r);
        const bool is_kept = kept >> lane & 1u;
        int start = -1;                                     // where the walk back starts
        if (is_kept) {
            d_r = excl + __popc(kept & ((1u << lane) - 1u));
            if (lk.x >= 0) start = lk.x;
            else {
                a.link_next[d_r].x = lk.x;                  // chain head / not a chain row
                if (lk.x == -2) a.link_next[d_r].y = -2;
            }
            if (lk.x == -2 || lk.y < 0) {
                self_dst = d_r;
                if (lk.x != -2) a.link_next[d_r].y = -1;
            }
        } else if (lk.x >= 0 && lk.y < 0) {
            start = r;                                      // merged away, and the chain ends here
        }
        if (start >= 0) {
            int x = start, L = 0;
            while (wait_flag(a.fflag, x, err) == 1u) {      // merged away: one more member, on to its predecessor
                ++L;
                x = __ldg(&a.link[x].x);
                if (x < 0) break;                           // (cannot happen: a chain head is never merged away)
            }
            if (x >= 0) {
                const int d_a = wait_dst(a.fdst, x, err);
                if (d_a >= 0) {
                    run_last = start; run_anchor = x; run_L = L; run_dst = d_a;
                    if (is_kept) { a.link_next[d_r].x = d_a; a.link_next[d_a].y = d_r; }
                    else a.link_next[d_a].y = -1;
                }
            }
        }
    }
    __syncwarp();
    const int row_bytes = a.row_bytes;
#pragma unroll 1
    for (int w = 0; w < W; ++w) {
        const int anchor = __shfl_sync(FULL, run_anchor, w), last = __shfl_sync(FULL, run_last, w);
        const int L = __shfl_sync(FULL, run_L, w), d_a = __shfl_sync(FULL, run_dst, w);
        const int sd = __shfl_sync(FULL, self_dst, w), d_w = __shfl_sync(FULL, d_r, w);
        if (anchor >= 0) {
            if (L == 0) worker_copy(ws, a.hidden + (int64_t)anchor * row_bytes, a.out + (int64_t)d_a * row_bytes, row_bytes, lane);
            else worker_run<DT>(a, ws, anchor, last, L, d_a, lane);
        }
        if (sd >= 0) worker_copy(ws, a.hidden + (int64_t)(tile * W + w) * row_bytes, a.out + (int64_t)sd * row_bytes, row_bytes, lane);
        if (d_w >= 0 && aux.n) worker_aux(a, aux, tile * W + w, d_w, lane);
    }
}

// CTA = two front warpgroups (warps 0 .. 7: the W tile warps, W rows per tile, two shared-memory slots each) + one back
// warpgroup (warp 8: scan, warps 9 .. 11: workers).  The warpgroups trade registers (setmaxnreg): the front needs few, the
// back keeps whole rows in flight.
template <int DT>
__global__ void __launch_bounds__((FU_WARPS + FU_WORKERS) * 32, 2)
k_fused_merge(const __grid_constant__ FusedArgs a, const __grid_constant__ AuxPack aux) {
    extern __shared__ __align__(128) unsigned char fu_smem[];
    pdl_enter();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, W = a.tile_rows;
    unsigned char* wslots = fu_smem + (size_t)(2 * W) * a.slot_bytes;        // behind the front's slots: the workers'
    FusedShared* sh = reinterpret_cast<FusedShared*>(wslots + (size_t)FU_WSLOTS * a.slot_bytes);
    ScanArrays sa;
    sa.round_base = a.desc + 1;
    sa.slot = reinterpret_cast<unsigned*>(a.desc + 1 + a.nrounds + 1);
    sa.excl = sa.slot + a.ntiles_pad;
    sa.round_cnt = sa.excl + a.ntiles_pad;
    if (wid < W && lane == 0) mbar_init(smem_u32(&sh->bars[wid]), 1);
    if (threadIdx.x < FU_WSLOTS) mbar_init(smem_u32(&sh->wbars[threadIdx.x]), 1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < FU_TICKETS; ++i) sh->next_iter[i] = 0;
        for (int i = 0; i < FU_WARPS; ++i) sh->progress[i] = 0;
        sh->scan_tail = sh->scan_head = 0;
        sh->q.tail = sh->q.head = sh->q.done = 0;
    }
    for (int i = threadIdx.x; i < FU_QSIZE; i += blockDim.x) sh->q.seq[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    int err = 0;
    const int nvec = a.nvec;
    const int64_t row_bytes = a.row_bytes;

    // each role's code lives entirely inside its branch: the register budget set at its top holds to its end
    if (wid > FU_WARPS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(FU_REGS_BACK));
        // ---- workers: copies nobody waits for
        WorkerSlots ws;
        const int me = wid - FU_WARPS - 1;
        for (int s = 0; s < 2; ++s) {
            ws.ptr[s] = wslots + (size_t)(2 * me + s) * a.slot_bytes;
            ws.addr[s] = smem_u32(ws.ptr[s]);
            ws.bar[s] = smem_u32(&sh->wbars[2 * me + s]);
            ws.phase[s] = 0;
            ws.state[s] = 0;
            ws.dst[s] = nullptr;
        }
        ws.next = 0;
        unsigned long long item;
        while (queue_pop(&sh->q, lane, &item)) run_tile_item<DT>(a, aux, ws, W, item, lane, &err);
        slot_free(ws, 0, a.row_bytes, lane);
        slot_free(ws, 1, a.row_bytes, lane);
        if (lane == 0) tma_wait_all();
        if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
        return;
    }

    if (wid == FU_WARPS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(FU_REGS_BACK));
        // ---- scan warp: destinations.  Two steps per tile, pipelined: POST as soon as the front has flagged the tile's rows
        // (mask into the tile's slot, bump the round, scan the round if this was its last tile), RESOLVE when the tile's
        // exclusive prefix has been written (publish the destinations, hand the tile to the workers).  Posting never waits for
        // resolving, so rounds fill at the front's pace.
        constexpr int PEND = 8;
        unsigned long long pend[PEND];                      // posted, not yet resolved: (tile << 16) | (merged << 8) | kept
        int n_pend = 0, p_head = 0;
        unsigned head = 0;
        long long cur = -1;                                 // tile being posted
        bool exiting = false;
        int idle = 0;
        while (true) {
            bool progress = false;
            if (cur < 0 && !exiting && n_pend < PEND && *(volatile unsigned*)&sh->scan_tail != head) {
                __threadfence_block();
                const unsigned long long item = *(volatile unsigned long long*)&sh->scan_item[head % FU_SCANQ];
                __syncwarp();
                ++head;
                if (lane == 0) *(volatile unsigned*)&sh->scan_head = head;
                if (item == FU_ITEM_EXIT) exiting = true;
                else cur = (long long)item;
                progress = true;
            }
            // both polls travel together
            unsigned f = 2u;
            if (cur >= 0) {
                const int r = (int)cur * W + lane;
                if (lane < W && r < a.S) f = ld_relaxed32(a.fflag + r);
            }
            unsigned e = 0;
            const int ptile = n_pend ? (int)(pend[p_head] >> 16) : 0;
            if (n_pend) e = ld_relaxed32(sa.excl + ptile);
            if (cur >= 0 && !__ballot_sync(FULL, f == 0u)) {
                // POST
                const int tile = (int)cur, r = tile * W + lane;
                const bool mine = lane < W && r < a.S;
                const unsigned kept = __ballot_sync(FULL, mine && f == 2u), merged = __ballot_sync(FULL, mine && f == 1u);
                if (lane == 0) FU_STAMP(tile, 6);
                const int round = tile / FU_ROUND;
                const int round_size = min(FU_ROUND, a.ntiles - round * FU_ROUND);
                unsigned old = 0;
                if (lane == 0) {
                    st_relaxed32(sa.slot + tile, 0x100u | kept);
                    __threadfence();                        // the slot is visible before the bump that may complete the round
                    old = atomicAdd(sa.round_cnt + round, 1u);
                }
                old = __shfl_sync(FULL, old, 0);
                if ((int)old == round_size - 1) round_scan(sa, round, a.ntiles, lane, &err);
                pend[(p_head + n_pend) % PEND] = ((unsigned long long)tile << 16) | (merged << 8) | kept;
                ++n_pend;
                cur = -1;
                progress = true;
            }
            if (n_pend && e != 0u) {
                // RESOLVE
                const unsigned long long pe = pend[p_head];
                const int tile = ptile, excl = (int)e - 1, r = tile * W + lane;
                const unsigned kept = (unsigned)pe & 0xffu, merged = (unsigned)(pe >> 8) & 0xffu;
                const int total = __popc(kept);
                if (lane == 0) FU_STAMP(tile, 7);
                if (kept >> lane & 1u) {
                    const int d = excl + __popc(kept & ((1u << lane) - 1u));
                    st_relaxed32(a.fdst + r, (unsigned)d + 1u);
                    a.dst[r] = d;
                } else if (merged >> lane & 1u) {
                    a.dst[r] = -1;
                }
                if (kept | merged) {
                    const unsigned long long it = ((unsigned long long)tile << 38) | ((unsigned long long)excl << 8) | kept;
                    if (lane == 0) {                        // ring full: the workers are behind, and so is everything upstream
                        int spins = 0;
                        while (!queue_push(&sh->q, it)) {
                            if (++spins > FU_SPIN_LIMIT) { err = 1; break; }
                            __nanosleep(200);
                        }
                    }
                }
                if (lane == 0 && tile == a.ntiles - 1) {
                    // the sequence is done: sizes, the speculated branch, the counters of the next call (main.py:112-120)
                    const long long s_keep = excl + total, n_merged = a.S - s_keep;
                    const long long N = a.counters[C_N], n_vis = a.counters[C_NVIS];
                    int ec = 0;
                    if (n_vis == 0) ec = 1;                 // the reference divides by zero here (main.py:114)
                    else if (!((double)n_merged / (double)n_vis < a.bound)) ec = 3;   // top-k branch: the host redoes the call
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
                    a.status[FF_ST_FUSED] = 1;
                }
                __syncwarp();
                if (lane == 0) FU_STAMP(tile, 8);
                p_head = (p_head + 1) % PEND;
                --n_pend;
                progress = true;
            }
            if (exiting && cur < 0 && n_pend == 0) break;
            if (progress) idle = 0;
            else {
                if (++idle > (FU_SPIN_LIMIT << 4)) { err = 1; break; }      // seconds without progress: give up
                __nanosleep(100);
            }
        }
        if (lane == 0) {
            __threadfence_block();
            *(volatile unsigned*)&sh->q.done = 1u;
        }
        if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
        return;
    }

    // ---- the front warpgroups: W tile warps and, in warp FU_WARPS - 1, the dispatcher
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(FU_REGS_FRONT));
    if (wid == FU_WARPS - 1) {
        // ---- dispatcher: takes the tiles' tickets a few iterations ahead of the tile warps (an atomic on one hot word is a
        // microsecond away), tells the scan warp about every tile, ends both streams.  Never more than FU_TICKETS - 1 iterations
        // ahead of the slowest tile warp: a slot of the ticket ring is rewritten only when everybody has read it.
        for (int k = 0;; ++k) {
            int spins = 0;
            for (int w = 0; w < W; ++w)
                while (*(volatile int*)&sh->progress[w] < k - (FU_TICKETS - 1)) {
                    if (++spins > (FU_SPIN_LIMIT << 4)) { err = 1; break; }
                    __nanosleep(100);
                }
            int t = 0;
            if (lane == 0) t = (int)atomicAdd(a.desc, 1ull);
            t = __shfl_sync(FULL, t, 0);
            if (lane == 0) {
                const unsigned q = sh->scan_tail;           // the scan warp learns of the tile now: it waits for its flags
                spins = 0;
                while ((int)(q - *(volatile unsigned*)&sh->scan_head) >= FU_SCANQ) {
                    if (++spins > (FU_SPIN_LIMIT << 4)) { err = 1; break; }
                    __nanosleep(100);
                }
                *(volatile unsigned long long*)&sh->scan_item[q % FU_SCANQ] = t < a.ntiles ? (unsigned long long)t : FU_ITEM_EXIT;
                __threadfence_block();
                *(volatile unsigned*)&sh->scan_tail = q + 1;
                if (t < a.ntiles) FU_STAMP(t, 0);
                *(volatile int*)&sh->next_tile[k % FU_TICKETS] = t;
                __threadfence_block();
                *(volatile int*)&sh->next_iter[k % FU_TICKETS] = k + 1;
            }
            if (t >= a.ntiles) break;
        }
        if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
        return;
    }
    if (wid >= W) return;                                   // rows too long for seven slot pairs: fewer tile warps

    // ---- tile warps (the front).  Per warp: slot P (chain predecessor) and slot C (own row).
    unsigned char* slot_p = fu_smem + (size_t)(2 * wid) * a.slot_bytes;
    unsigned char* slot_c = slot_p + a.slot_bytes;
    const uint32_t sp32 = smem_u32(slot_p), sc32 = smem_u32(slot_c), bar = smem_u32(&sh->bars[wid]);
    const uint4* pr = reinterpret_cast<const uint4*>(slot_p);
    const uint4* cr = reinterpret_cast<const uint4*>(slot_c);
    uint32_t phase = 0;

    for (int iter = 0;; ++iter) {
        const int tk = iter % FU_TICKETS;
        {
            int spins = 0;
            while (*(volatile int*)&sh->next_iter[tk] != iter + 1) {          // (the dispatcher's own waits are bounded)
                if (++spins > (FU_SPIN_LIMIT << 5)) { err = 1; break; }
                __nanosleep(50);
            }
        }
        __threadfence_block();
        const int tile = err ? a.ntiles : *(volatile int*)&sh->next_tile[tk];
        if (tile >= a.ntiles) break;
        const int r = tile * W + wid;
        const bool valid = r < a.S;
#ifdef FF_FUSED_TRACE
        if (wid == W - 1 && lane == 0 && a.trace) { FU_STAMP(tile, 2); a.trace[(size_t)tile * FU_TRACE_SLOTS + 13] = blockIdx.x + 1; }
#endif
        int2 lk = make_int2(-2, -2);
        if (valid) lk = __ldg(a.link + r);
        const int p = lk.x;
        const bool has_pred = valid && p >= 0;
        __syncwarp();                                       // every lane is done with the slots before they are refilled
        if (has_pred && lane == 0) {
            mbar_expect_tx(bar, (uint32_t)row_bytes * 2u);
            tma_load(sc32, a.hidden + (int64_t)r * row_bytes, (uint32_t)row_bytes, bar);
            tma_load(sp32, a.hidden + (int64_t)p * row_bytes, (uint32_t)row_bytes, bar);
        }
        if (wid == W - 1 && lane == 0) FU_STAMP(tile, 4);
        if (valid) prefetch_aux(aux, r, lane);
        if (has_pred) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        if (wid == W - 1 && lane == 0) FU_STAMP(tile, 5);
        if (lane == 0) FU_STAMP_MAX(tile, 10);
        if (has_pred) {
            // three row sums in float32 (the reference's reductions accumulate in float32; their order is torch's own and
            // unknown, which the oracle brackets).  One warp has the whole row to itself, so the sums are split over four
            // independent chains per lane — packed pairs of even / odd elements, alternating vectors — or the dependent adds
            // alone would take microseconds.
            float2 d0 = make_float2(0.f, 0.f), d1 = d0, a0 = d0, a1 = d0, b0 = d0, b1 = d0;
            const int nfull = nvec >> 5;
            int v = 0;
#pragma unroll 2
            for (; v + 2 <= nfull; v += 2) {
                acc_pair2<DT>(pr[v * 32 + lane], cr[v * 32 + lane], d0, a0, b0);
                acc_pair2<DT>(pr[v * 32 + 32 + lane], cr[v * 32 + 32 + lane], d1, a1, b1);
            }
            if (v < nfull) acc_pair2<DT>(pr[v * 32 + lane], cr[v * 32 + lane], d0, a0, b0);
            if (nfull * 32 + lane < nvec) acc_pair2<DT>(pr[nfull * 32 + lane], cr[nfull * 32 + lane], d1, a1, b1);
            const float dot = warp_sum((d0.x + d0.y) + (d1.x + d1.y));
            const float na = warp_sum((a0.x + a0.y) + (a1.x + a1.y));
            const float nb = warp_sum((b0.x + b0.y) + (b1.x + b1.y));
            const float s = finish_cosine<DT>(dot, na, nb);
            if (lane == 0) {
                a.sim_seq[r] = s;
                st_relaxed32(a.fflag + r, (s >= a.thr) ? 1u : 2u);          // NaN compares false: kept
            }
        } else if (valid && lane == 0) {
            a.sim_seq[r] = -2.0f;                           // IGNORE_TOKEN at chain heads (main.py:225-238)
            st_relaxed32(a.fflag + r, 2u);
        }
        if (wid == W - 1 && lane == 0) FU_STAMP(tile, 14);
        if (lane == 0) {
            FU_STAMP_MAX(tile, 11);
            *(volatile int*)&sh->progress[wid] = iter + 1;
        }
    }
    if (lane == 0) *(volatile int*)&sh->progress[wid] = 0x7fffffff;
    // leave the other bank's flags, destinations and descriptors zeroed for the next call of the prefill
    const int64_t n_thr = (int64_t)gridDim.x * W * 32, me = (int64_t)blockIdx.x * W * 32 + threadIdx.x;
    for (int64_t i = me; i < a.S; i += n_thr) { a.fflag_clr[i] = 0ull; a.fdst_clr[i] = 0u; }
    for (int64_t i = me; i < a.desc_words; i += n_thr) a.desc_clr[i] = 0ull;
    if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
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
