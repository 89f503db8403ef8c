// Read-once merge kernel: similarity + threshold select + run merge + compaction of hidden_states and of the aux
// tensors (cos / sin / patch_type / position ids) in ONE launch that fetches every row of hidden_states from HBM
// once (main.py:104-138, threshold branch — every merge call of a prefill but possibly the last).
//
// Shape of the problem.  A token is compared with the previous surviving token of the SAME patch id (its chain
// predecessor, main.py:216-238) — 576 rows (4 MB) back on the first call of a uniform video — while the output is
// compacted in SEQUENCE order (main.py:132-138): a row's destination is the number of kept rows before it, known only
// when every earlier row has been compared.  So every row is visited twice by the warps of one persistent grid — once
// for its similarity (from HBM), once for the gather (from the L2, a few thousand rows later) — and the warps are
// independent workers that take units of work, one row each, from ONE ticket:
//
//   S unit (row r).  The row (HBM, plain loads: `ld.global.nc.L1::no_allocate` data is dropped from the L2 first, and
//     the gather would find nothing) and its chain predecessor (an L2 hit: some S unit's own row a moment ago), three
//     row sums, the reference's rounding chain, sim >= thr -> merged away.  ONE relaxed 64-bit reduction into
//     word[r / 32] carries both facts the gather needs: (1 << 32) | (kept << (r % 32)) — rows reported in the high
//     half, kept mask in the low half.  An S unit waits for nothing.
//   G unit (row r), `lag` rows behind the S units.  It waits until the words up to its own are complete — then every
//     flag of every earlier row is known — and counts the kept rows before r: the prefix of its band (1024 rows: one
//     word per lane), carried from band to band in registers, plus the masks of the band.  A kept row ends the run
//     of its chain predecessor: the warp walks the masks back to the run's anchor and writes the anchor's destination
//     row: the raw row if the run has no members, else T(T(..T(anchor + m1) + ..) + mL) / T(L+1), one rounding to T per
//     add in chain order (the sequence torch-CPU index_add_ performs, main.py:304-311) and one division
//     (main.py:314-317).  Chain tails end their own run, rows outside the chains are copied.  The aux rows and the
//     (pred, succ) links of the next call go with it.
//
// Tickets run in a fixed order — S(0 .. lag-1), then 32 S units and 32 G units alternating — and a CTA draws them eight
// at a time (one global atomic per eight units; its warps pick them from shared memory in order, the next eight already
// requested).  Every wait is for work with a smaller ticket that a running warp has picked or will pick before anything
// larger: no deadlock whatever is resident; every spin is bounded all the same (FF_ST_INTERNAL).
//
// The branch decision (main.py:114-116) needs the global count, known only at the end: the kernel speculates on the
// threshold branch, the G unit of the last row checks count / n_vis < bound and otherwise reports FF_ST_ERROR = 3; the
// host then redoes the call with the multi-kernel path (top-k branch, at most once per prefill).  The input is never
// modified, so the redo sees the original rows.
#pragma once
#include "ff_common.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FU_WARPS = 8;                        // warps per CTA (independent workers; a CTA shares its ticket ring)
constexpr int FU_MIN_CTAS = 4;                     // per SM: 32 warps of at most 64 registers per thread
constexpr int FU_BAND = 1024;                      // rows per band: 32 words of 32 rows, one word per lane
constexpr int FU_LAG = 8192;                       // S units run this many rows ahead of the G units (FF_FUSED_LAG)
constexpr int FU_SFRAMES = 2;                      // frames whose S units are taken patch by patch (FF_FUSED_SFRAMES; 1: sequence order)
constexpr int FU_BATCH = 8;                        // tickets per draw
#ifndef FU_PREFETCH_SLOT
#define FU_PREFETCH_SLOT (FU_BATCH / 2)
#endif
constexpr int FU_RING = 16;                        // draws a CTA remembers
constexpr int FU_SPIN_LIMIT = 1 << 20;             // polls (~100 ns apart) before a wait gives up and reports FF_ST_INTERNAL

// the aux tensors, one entry per (tensor, plane): rows of at most 512 bytes in 16- or 8-byte pieces, one piece per lane
struct AuxFlat {
    int n;                                         // entries; -1: the tensors do not fit this form (gather_aux_rows instead)
    int row_bytes[8];
    int piece[8];                                  // 16 or 8: bytes per lane
    const char* src[8];
    char* dst[8];
};

struct FusedArgs {
    AuxFlat auxf;
    const char* hidden;
    char* out;
    int S, nvec, row_bytes, nwords, nbands, lag, n_tickets;
    int perm_P, perm_F;                            // order of the S units on a uniform video (0: sequence order): see s_row()
    const int2* link;                              // [S] (pred, succ): row index, -1 = chain head / tail, -2 = not a chain row
    int2* link_next;                               // [S_keep] the same for the compacted sequence
    unsigned long long* desc;                      // zero on entry, desc_words u64 in all: the ticket, word [nwords]
                                                   // ((rows reported << 32) | kept mask of rows 32 i .. 32 i + 31), then u32
                                                   // band_excl [nbands] (kept rows before the band + 1)
    int desc_words;
    unsigned long long* desc_clr;                  // other bank: cleared for the next call
    float* sim_seq;                                // [S] similarity with the chain predecessor (introspection)
    int* dst;                                      // [S] destination row or -1
    int64_t* counters;
    int64_t* counters_next;
    int64_t* status;
    float thr;
    double bound;
    long long seq;                                 // number of this call: status[FF_ST_SEQ] once the block is complete
};

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

// ---- tickets: a CTA draws FU_BATCH at a time; its warps pick them in order from shared memory -----------------------
struct TicketRing {
    unsigned taken;                                         // picks of this CTA so far
    int base[FU_RING];                                      // first ticket of draw i (slot i % FU_RING)
    unsigned ready[FU_RING];                                // i + 1 once base is there
};

__device__ __forceinline__ void ticket_init(TicketRing* tr, unsigned long long* gticket) {
    if (threadIdx.x < FU_RING) tr->ready[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        tr->taken = 0u;
        tr->base[0] = (int)atomicAdd(gticket, (unsigned long long)FU_BATCH);
        tr->ready[0] = 1u;
    }
    __syncthreads();
}

// whole warp; n_tickets (or more) when the work is exhausted or the ring timed out
__device__ __forceinline__ int ticket_take(TicketRing* tr, unsigned long long* gticket, int n_tickets, int lane, int* err) {
    unsigned n = 0;
    if (lane == 0) n = atomicAdd(&tr->taken, 1u);
    n = __shfl_sync(FULL, n, 0);
    const unsigned draw = n / FU_BATCH, slot = n % FU_BATCH;
    if (slot == FU_PREFETCH_SLOT && lane == 0) {
        // half of this draw is gone: request the next one (this lane alone waits for the answer)
        const int b = (int)atomicAdd(gticket, (unsigned long long)FU_BATCH);
        *(volatile int*)&tr->base[(draw + 1) % FU_RING] = b;
        __threadfence_block();
        *(volatile unsigned*)&tr->ready[(draw + 1) % FU_RING] = draw + 2u;
    }
    __syncwarp();
    int spins = 0;
    while (*(volatile unsigned*)&tr->ready[draw % FU_RING] != draw + 1u) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; return n_tickets; }
        __nanosleep(40);
    }
    __threadfence_block();
    return *(volatile int*)&tr->base[draw % FU_RING] + (int)slot;
}

// ---- flags and prefixes ------------------------------------------------------------------------------------------
struct FusedDesc {
    unsigned long long* ticket;
    unsigned long long* word;                               // [nwords] (rows reported << 32) | kept mask
    unsigned* band_excl;                                    // [nbands] kept rows before the band + 1
    int S, nwords;
};

// the 32 words of band b, one per lane (words behind the sequence read as 0); returns once the first `upto` of them are
// complete — every row of theirs has been compared, by S units with smaller tickets
__device__ __forceinline__ unsigned long long band_words(const FusedDesc& d, int b, int upto, int lane, int* err) {
    const int i = b * 32 + lane;
    const bool have = i < d.nwords;
    const unsigned need = have ? (unsigned)min(32, d.S - i * 32) : 0u;
    unsigned long long w = 0ull;
    int spins = 0;
    while (true) {
        if (have) w = ld_relaxed64(d.word + i);
        const bool ok = lane >= upto || (unsigned)(w >> 32) == need;
        if (__all_sync(FULL, ok)) break;
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(100);
    }
    return w;
}

// What a warp remembers between its G units: the band it last worked in, the kept rows before that band and before the
// band in front of it.  A warp enters band b only when all earlier bands are complete, so the prefix of the next band
// follows from the cached one and the band's words — nobody waits for a publisher.
struct BandCache {
    int cb, ce, pe;                                         // band, kept rows before it, kept rows before band cb - 1
};

__device__ __forceinline__ int wait_band_excl(const FusedDesc& d, int b, int* err) {
    unsigned v = ld_relaxed32(d.band_excl + b);
    int spins = 0;
    while (v == 0u) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(100);
        v = ld_relaxed32(d.band_excl + b);
    }
    return (int)v - 1;
}

// kept rows before row x, x in an earlier row than the caller's and therefore flagged (whole warp); *mx = the mask of x's word
__device__ __forceinline__ int prefix_at(const FusedDesc& d, const BandCache& bc, int x, int lane, unsigned* mx, int* err) {
    const int b = x >> 10, wi = (x >> 5) & 31;
    const unsigned long long w = ld_relaxed64(d.word + min(b * 32 + lane, d.nwords - 1));
    const int e = b == bc.cb ? bc.ce : (b == bc.cb - 1 ? bc.pe : wait_band_excl(d, b, err));
    const int cnt = warp_sum_int(lane < wi ? __popc((unsigned)w) : 0);
    const unsigned m = __shfl_sync(FULL, (unsigned)w, wi);
    *mx = m;
    return e < 0 ? -1 : e + cnt + __popc(m & ((1u << (x & 31)) - 1u));
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

// the row an S unit reads for the first time: it must stay in the L2 for its successor's S unit and for the gather
#ifdef FU_EVICT_LAST
__device__ __forceinline__ uint4 ld_keep16(const void* p) {
    uint4 r;
    asm volatile("{\n.reg .b64 pol;\ncreatepolicy.fractional.L2::evict_last.b64 pol, 1.0;\n"
                 "ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], pol;\n}"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
#else
__device__ __forceinline__ uint4 ld_keep16(const void* p) { return ldg16(p); }
#endif

struct SimAcc {
    float2 d0, d1, a0, a1, b0, b1;
};

template <int DT, int N>
__device__ __forceinline__ void sim_piece(const char* __restrict__ pr, const char* __restrict__ cr, SimAcc& s) {
    uint4 p[N], c[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
        c[q] = ld_keep16(cr + q * 512);
        p[q] = ldg16(pr + q * 512);
    }
#pragma unroll
    for (int q = 0; q < N; ++q) {
        if (q & 1) acc_pair2<DT>(p[q], c[q], s.d1, s.a1, s.b1);
        else acc_pair2<DT>(p[q], c[q], s.d0, s.a0, s.b0);
    }
}

// cosine similarity of rows pr (predecessor) and cr with the reference's rounding chain (main.py:345-349).  The three
// row sums are float32 (ATen accumulates the reductions in float32; their order is torch's own and unknown, which the
// oracle brackets), split over four independent chains per lane; four + four vectors in flight.
template <int DT>
__device__ __forceinline__ float row_similarity(const char* __restrict__ pr, const char* __restrict__ cr, int nvec, int lane) {
    SimAcc s;
    s.d0 = s.d1 = s.a0 = s.a1 = s.b0 = s.b1 = make_float2(0.f, 0.f);
    int v = 0;
    pr += lane * 16;
    cr += lane * 16;
#pragma unroll 1
    for (; v + 128 <= nvec; v += 128) sim_piece<DT, 4>(pr + (int64_t)v * 16, cr + (int64_t)v * 16, s);
    if (v + 64 <= nvec) { sim_piece<DT, 2>(pr + (int64_t)v * 16, cr + (int64_t)v * 16, s); v += 64; }
    if (v + 32 <= nvec) { sim_piece<DT, 1>(pr + (int64_t)v * 16, cr + (int64_t)v * 16, s); v += 32; }
    if (v + lane < nvec) sim_piece<DT, 1>(pr + (int64_t)v * 16, cr + (int64_t)v * 16, s);
    const float dot = warp_sum((s.d0.x + s.d0.y) + (s.d1.x + s.d1.y));
    const float na = warp_sum((s.a0.x + s.a0.y) + (s.a1.x + s.a1.y));
    const float nb = warp_sum((s.b0.x + s.b0.y) + (s.b1.x + s.b1.y));
    return finish_cosine<DT>(dot, na, nb);
}

// the members of a run in chain order: lane k of `mine` holds the k-th member from the END (runs of at most 32 members);
// longer runs follow the successor links
struct RunWalk {
    const int2* link;
    int L, mine, anchor;
    __device__ __forceinline__ int first() const { return L <= 32 ? __shfl_sync(FULL, mine, L - 1) : __ldg(&link[anchor].y); }
    __device__ __forceinline__ int next(int prev, int m) const {            // the m-th member, m >= 2
        return L <= 32 ? __shfl_sync(FULL, mine, L - m) : __ldg(&link[prev].y);
    }
};

// N vectors per lane of a run: T(T(..T(anchor + m1) ..+ mL) / T(L + 1)), one rounding to T per add in chain order
template <int DT, int N>
__device__ __forceinline__ void run_piece(const char* __restrict__ hidden, int64_t row_bytes, int64_t off, const RunWalk& w,
                                          int first, const Divider<DT>& dv, char* __restrict__ orow) {
    uint4 acc[N], x[N];
    const char* ar = hidden + (int64_t)w.anchor * row_bytes + off;
    const char* mr = hidden + (int64_t)first * row_bytes + off;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        acc[q] = ld_stream16(ar + q * 512);
        x[q] = ld_stream16(mr + q * 512);
    }
    int walk = first;
#pragma unroll 1
    for (int m = 1; m <= w.L; ++m) {
        if (m > 1) {
            walk = w.next(walk, m);
            mr = hidden + (int64_t)walk * row_bytes + off;
#pragma unroll
            for (int q = 0; q < N; ++q) x[q] = ld_stream16(mr + q * 512);
        }
#pragma unroll
        for (int q = 0; q < N; ++q) acc[q] = Num<DT>::add_vec(acc[q], x[q]);              // T(acc + member), main.py:304
    }
    if (dv.pow2) {
#pragma unroll
        for (int q = 0; q < N; ++q) st_stream16(orow + off + q * 512, Num<DT>::scale_vec(acc[q], dv.rcp));
    } else if (dv.by_rcp) {
#pragma unroll
        for (int q = 0; q < N; ++q) st_stream16(orow + off + q * 512, dv.vec_rcp(acc[q]));
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) st_stream16(orow + off + q * 512, dv.vec(acc[q]));     // IEEE division, out of line
    }
}

// A run: anchor row and its L >= 1 members -> destination row (main.py:304-317)
template <int DT>
__device__ __forceinline__ void sum_run(const char* __restrict__ hidden, int nvec, int64_t row_bytes, const RunWalk& w,
                                        char* __restrict__ orow, int lane) {
    const Divider<DT> dv(w.L + 1);
    const int first = w.first();
    int v = 0;
    const int64_t lo = lane * 16;
#pragma unroll 1
    for (; v + 128 <= nvec; v += 128) run_piece<DT, 4>(hidden, row_bytes, lo + (int64_t)v * 16, w, first, dv, orow);
    if (v + 64 <= nvec) { run_piece<DT, 2>(hidden, row_bytes, lo + (int64_t)v * 16, w, first, dv, orow); v += 64; }
    if (v + 32 <= nvec) { run_piece<DT, 1>(hidden, row_bytes, lo + (int64_t)v * 16, w, first, dv, orow); v += 32; }
    // (the shuffles inside a piece are warp-wide: the partial piece is walked by every lane, out-of-range lanes re-read
    // their first vector and store nothing)
    if (v < nvec) {
        const bool in = v + lane < nvec;
        const int64_t off = in ? lo + (int64_t)v * 16 : lo;
        uint4 acc = ld_stream16(hidden + (int64_t)w.anchor * row_bytes + off);
        int walk = first;
#pragma unroll 1
        for (int m = 1; m <= w.L; ++m) {
            if (m > 1) walk = w.next(walk, m);
            acc = Num<DT>::add_vec(acc, ld_stream16(hidden + (int64_t)walk * row_bytes + off));
        }
        if (in) st_stream16(orow + off, dv.vec_fast(acc));
    }
}

// the aux rows of sequence row r -> destination row d: one 16- or 8-byte piece per lane and entry, all loads first
__device__ __forceinline__ void fused_aux(const FusedArgs& a, const AuxPack& aux, int r, int d, int lane) {
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

// the sequence is done: sizes, the speculated branch, the counters of the next call (main.py:112-120)
__device__ __forceinline__ void fused_finish(const FusedArgs& a, long long s_keep) {
    const long long n_merged = a.S - s_keep;
    const long long N = a.counters[C_N], n_vis = a.counters[C_NVIS];
    int ec = 0;
    if (n_vis == 0) ec = 1;                                 // the reference divides by zero here (main.py:114)
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
    __threadfence_system();
    *(volatile int64_t*)&a.status[FF_ST_SEQ] = a.seq;
}

// Order of the S units.  On the first call of a uniform video (frames x P vision rows in one span, chains = patches) the
// S units of F consecutive frames are taken patch by patch — (f, p), (f + 1, p), .., (f, p + 1) .. — so that the unit of
// (f + 1, p) runs next to the unit of (f, p), on the same SM at the same time: its predecessor row is the neighbour's own
// row (one request to the L2 instead of two that miss together), and the predecessor of (f, p) was fetched F frames of
// units ago and has arrived.  Any layout is served correctly (the map is a bijection of the rows whatever the links
// say); other layouts keep the sequence order.
struct SOrder {
    int first, span, P, F, frames;                          // span = frames * P rows from `first`; F = 0: identity
    __device__ __forceinline__ int row(int s) const {
        const int local = s - first;
        if (F == 0 || local < 0 || local >= span) return s;
        const int bi = local / (F * P), w = local - bi * (F * P), f0 = bi * F;
        const int fb = min(F, frames - f0);
        const int p = w / fb;
        return first + (f0 + (w - p * fb)) * P + p;
    }
};

// S unit of row r (see the head of the file)
template <int DT>
__device__ __forceinline__ void s_unit(const FusedArgs& a, const FusedDesc& d, int r, int lane) {
    const int2 lk = __ldg(a.link + r);
    float s = -2.0f;                                        // IGNORE_TOKEN at chain heads (main.py:225-238)
    if (lk.x >= 0) s = row_similarity<DT>(a.hidden + (int64_t)lk.x * a.row_bytes, a.hidden + (int64_t)r * a.row_bytes, a.nvec, lane);
    const unsigned kept = !(lk.x >= 0 && s >= a.thr);       // NaN compares false: kept
    if (lane == 0) {
        a.sim_seq[r] = s;
        red_add64(d.word + (r >> 5), (1ull << 32) | ((unsigned long long)kept << (r & 31)));
    }
}

// G unit of row r (see the head of the file)
template <int DT>
__device__ __forceinline__ void g_unit(const FusedArgs& a, const AuxPack& aux, const FusedDesc& d, BandCache& bc, int r, int lane, int* err) {
    const int2 lk = __ldg(a.link + r);
    const int b = r >> 10, wi = (r >> 5) & 31;
    const int nvec = a.nvec;
    const int64_t row_bytes = a.row_bytes;
    if (b != bc.cb) {
        // entering a band: the prefix moves along with the words of the bands in between, all complete by now
        if (bc.cb < 0) bc.cb = 0;
        while (bc.cb < b) {
            const unsigned long long w = band_words(d, bc.cb, 32, lane, err);
            bc.pe = bc.ce;
            bc.ce += warp_sum_int(__popc((unsigned)w));
            ++bc.cb;
        }
    }
    if (*err) return;
    if ((r & 1023) == 0 && lane == 0) st_relaxed32(d.band_excl + b, (unsigned)bc.ce + 1u);   // for walks that end in older bands
    const unsigned long long wown = band_words(d, b, wi + 1, lane, err);
    if (*err) return;
    const unsigned m = __shfl_sync(FULL, (unsigned)wown, wi);
    const int ex = bc.ce + warp_sum_int(lane < wi ? __popc((unsigned)wown) : 0) + __popc(m & ((1u << (r & 31)) - 1u));
    const bool is_kept = m >> (r & 31) & 1u;
    const int d_r = is_kept ? ex : -1;
    if (lane == 0) {
        a.dst[r] = d_r;
        if (r == a.S - 1) fused_finish(a, (long long)ex + (is_kept ? 1 : 0));
    }
    int start = -1;                                         // where the walk back starts
    bool self = false;
    if (is_kept) {
        if (lk.x >= 0) start = lk.x;
        else if (lane == 0) {
            a.link_next[d_r].x = lk.x;                      // chain head / not a chain row
            if (lk.x == -2) a.link_next[d_r].y = -2;
        }
        if (lk.x == -2 || lk.y < 0) {
            self = true;
            if (lk.x != -2 && lane == 0) a.link_next[d_r].y = -1;
        }
    } else if (lk.x >= 0 && lk.y < 0) {
        start = r;                                          // merged away, and the chain ends here
    }
    if (start >= 0) {
        // every lane walks (uniform loads); lane k remembers the k-th member from the end
        int x = start, L = 0, mine = -1;
        unsigned mx;
        while (true) {                                      // (the words of this band are at hand: no load)
            mx = (x >> 10) == b ? __shfl_sync(FULL, (unsigned)wown, (x >> 5) & 31) : (unsigned)ld_relaxed64(d.word + (x >> 5));
            if (mx >> (x & 31) & 1u) break;                 // kept: the anchor
            if (lane == L) mine = x;
            ++L;
            x = __ldg(&a.link[x].x);
            if (x < 0) break;                               // (cannot happen: a chain head is never merged away)
        }
        if (x >= 0) {
            int d_a;
            if ((x >> 10) == b) {
                const int xi = (x >> 5) & 31;
                d_a = bc.ce + warp_sum_int(lane < xi ? __popc((unsigned)wown) : 0) + __popc(mx & ((1u << (x & 31)) - 1u));
            } else {
                d_a = prefix_at(d, bc, x, lane, &mx, err);
            }
            if (d_a >= 0) {
                if (lane == 0) {
                    if (is_kept) { a.link_next[d_r].x = d_a; a.link_next[d_a].y = d_r; }
                    else a.link_next[d_a].y = -1;
                }
                char* orow = a.out + (int64_t)d_a * row_bytes;
                if (L == 0) copy_row(a.hidden + (int64_t)x * row_bytes, orow, nvec, lane);
                else {
                    RunWalk rw;
                    rw.link = a.link; rw.L = L; rw.mine = mine; rw.anchor = x;
                    sum_run<DT>(a.hidden, nvec, row_bytes, rw, orow, lane);
                }
            }
        }
    }
    if (self) copy_row(a.hidden + (int64_t)r * row_bytes, a.out + (int64_t)d_r * row_bytes, nvec, lane);
    if (is_kept && aux.n) fused_aux(a, aux, r, d_r, lane);
}

template <int DT>
__global__ void __launch_bounds__(FU_WARPS * 32, FU_MIN_CTAS)
k_fused_merge(const __grid_constant__ FusedArgs a, const __grid_constant__ AuxPack aux) {
    __shared__ TicketRing ring;
    pdl_enter();
    const int lane = threadIdx.x & 31;
    FusedDesc d;
    d.ticket = a.desc;
    d.word = a.desc + 1;
    d.band_excl = reinterpret_cast<unsigned*>(a.desc + 1 + a.nwords);
    d.S = a.S;
    d.nwords = a.nwords;
    const int lag = a.lag, n_tickets = a.n_tickets;
    SOrder so;
    so.F = 0;
    if (a.perm_F > 1 && a.perm_P > 0) {
        const long long inv = a.counters[C_FIRSTINV], nvis = a.counters[C_NVIS];
        if (inv > 0 && nvis > 0 && nvis % a.perm_P == 0 && (long long)a.S - inv + nvis <= a.S) {
            so.first = (int)(a.S - inv);
            so.span = (int)nvis;
            so.P = a.perm_P;
            so.F = a.perm_F;
            so.frames = (int)(nvis / a.perm_P);
        }
    }
    int err = 0;
    BandCache bc;
    bc.cb = -1; bc.ce = 0; bc.pe = 0;
    ticket_init(&ring, d.ticket);
#pragma unroll 1
    while (true) {
        const int k = ticket_take(&ring, d.ticket, n_tickets, lane, &err);
        if (k >= n_tickets) break;
        // ticket -> unit: S(0 .. lag-1), then 32 S units (rows lag + 32 i ..) and 32 G units (rows 32 i ..) alternating
        bool is_g = false;
        int r = k;
        if (k >= lag) {
            const int j = k - lag, blk = j >> 6, o = j & 63;
            is_g = o >= 32;
            r = is_g ? (blk << 5) + o - 32 : lag + (blk << 5) + o;
        }
        if (r < a.S) {
            if (is_g) g_unit<DT>(a, aux, d, bc, r, lane, &err);
            else s_unit<DT>(a, d, so.row(r), lane);
        }
    }
    // leave the other bank's ticket, words and prefixes zeroed for the next call of the prefill
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.desc_words; i += (int64_t)gridDim.x * blockDim.x)
        a.desc_clr[i] = 0ull;
    if (err && lane == 0) a.status[FF_ST_INTERNAL] = 1;
}

// (pred, succ) of every sequence row from the compact by-patch arrays (ff_links.cuh / the scan kernels).  On the first
// call of a prefill (first_inv != null) it also says whether the sequence is a uniform video — ONE span of chain rows
// whose patch ids run 0, 1, .., n_ids - 1, 0, 1, .. — which is what the frame-pipelined kernel (ff_frame.cuh) serves.
__global__ void __launch_bounds__(256)
k_links_seq(const int* __restrict__ rank, const int* __restrict__ order, const int* __restrict__ chain,
            int64_t* __restrict__ counters, int S, int n_ids, int2* __restrict__ link, unsigned long long* first_inv) {
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
        if (first_inv) {
            const int jp = i > 0 ? rank[i - 1] : -1;
            if (jp < 0) {                                       // start of a span of chain rows
                atomicMax(first_inv, (unsigned long long)(S - i));
                atomicAdd((unsigned long long*)&counters[C_SPANS], 1ull);
                if (c != 0) counters[C_NONUNI] = 1;
            } else if (c != (chain[jp] + 1) % n_ids) {
                counters[C_NONUNI] = 1;
            }
        }
    }
    link[i] = l;
}

}  // namespace ff
