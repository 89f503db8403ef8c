// Read-once merge kernel: similarity + threshold select + run merge + compaction of hidden_states and of the aux
// tensors (cos / sin / patch_type / position ids) in ONE launch that fetches every row of hidden_states from HBM
// once (main.py:104-138, threshold branch — every merge call of a prefill but possibly the last).
//
// Shape of the problem.  A token is compared with the previous surviving token of the SAME patch id (its chain
// predecessor, main.py:216-238) — 576 rows (4 MB) back on the first call of a uniform video — while the output is
// compacted in SEQUENCE order (main.py:132-138): a row's destination is the number of kept rows before it, known only
// when every earlier row has been compared.  So the rows are visited twice, by two kinds of tiles of one persistent
// grid, and the second visit is served by the L2:
//
//   S tiles (similarity).  Tile t = rows [t * W, (t + 1) * W), one warp per row: the row (HBM) and its chain predecessor
//     (an L2 hit: it was some S tile's own row a few microseconds ago), three row sums, the reference's rounding chain,
//     sim >= thr -> merged away.  The tile's kept mask goes to tile_mask[t], and the round the tile belongs to (RT tiles
//     by number) receives one atomic: (1 << 32) | kept rows.  An S tile waits for nothing.
//   G tiles (gather).  Tile t again, `lag` tiles behind the S tiles.  It waits until every round up to its own is
//     complete — then every flag of every earlier row is known — sums the kept counts (rounds, then the masks of its own
//     round) into the tile's exclusive prefix, and publishes it (tile_excl[t]).  Per row, one warp: a kept row ends the run
//     of its chain predecessor — the warp walks the masks back to the run's anchor and writes the anchor's destination
//     row: the raw row if the run has no members, else T(T(..T(anchor + m1) + ..) + mL) / T(L+1), one rounding to T per
//     add in chain order (the sequence torch-CPU index_add_ performs, main.py:304-311) and one division
//     (main.py:314-317).  Chain tails end their own run, rows outside the chains are copied.  The aux rows and the
//     (pred, succ) links of the next call go with it.  All of these rows were read by S tiles at most `lag` + a run's
//     length ago: they come out of the L2.
//
// Tiles are handed out by ONE ticket in a fixed order — S(0 .. lag-1), then S(lag + i), G(i) alternating — so every
// wait is for work with a smaller ticket, held by a CTA that is running: no deadlock whatever is resident; every spin is
// bounded all the same (FF_ST_INTERNAL).
//
// The branch decision (main.py:114-116) needs the global count, known only at the end: the kernel speculates on the
// threshold branch, the last G tile checks count / n_vis < bound and otherwise reports FF_ST_ERROR = 3; the host then
// redoes the call with the multi-kernel path (top-k branch, at most once per prefill).  The input is never modified,
// so the redo sees the original rows.
#pragma once
#include "ff_common.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FU_WARPS = 8;                        // warps per CTA = rows per tile
constexpr int FU_MIN_CTAS = 2;                     // per SM: up to 128 registers per thread, sixteen 16-byte vectors per lane in flight
constexpr int FU_ROUND_TILES = 32;                 // tiles per round (completion is tracked per round)
constexpr int FU_LAG_TILES = 512;                  // S tiles run this far ahead of the G tiles (default; FF_FUSED_LAG)
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
    int S, nvec, row_bytes, ntiles, nrounds, lag;
    const int2* link;                              // [S] (pred, succ): row index, -1 = chain head / tail, -2 = not a chain row
    int2* link_next;                               // [S_keep] the same for the compacted sequence
    unsigned long long* desc;                      // zero on entry, desc_words u64 in all: the ticket, round words [nrounds],
                                                   // then u32 tile_excl [ntiles] (exclusive prefix + 1)
    unsigned* tile_mask;                           // [ntiles] kept mask of the tile's rows (needs no clearing)
    int desc_words;
    unsigned long long* desc_clr;                  // other bank: cleared for the next call
    float* sim_seq;                                // [S] similarity with the chain predecessor (introspection)
    int* dst;                                      // [S] destination row or -1
    int64_t* counters;
    int64_t* counters_next;
    int64_t* status;
    float thr;
    double bound;
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

// exclusive prefix of tile x (published by its G tile, which holds a smaller ticket); -1 after a time-out
__device__ __forceinline__ int wait_excl(const unsigned* tile_excl, int x, int* err) {
    unsigned v = ld_relaxed32(tile_excl + x);
    int spins = 0;
    while (v == 0u) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(100);
        v = ld_relaxed32(tile_excl + x);
    }
    return (int)v - 1;
}

// ---- row movers: 16-byte vectors; a row is cut into pieces of N vectors per lane (N = 8, 4, 2, 1: 256 ... 32 vectors)
// and a last partial piece, so that every piece issues its N (or 2 N) loads back to back with no predicate in between —
// predicated loads are not batched by ptxas, and a warp with one or two loads in flight is latency bound.
template <int N>
__device__ __forceinline__ void copy_piece(const char* __restrict__ src, char* __restrict__ dst) {
    uint4 x[N];
#pragma unroll
    for (int q = 0; q < N; ++q) x[q] = ld_stream16(src + q * 512);
#pragma unroll
    for (int q = 0; q < N; ++q) st_stream16(dst + q * 512, x[q]);
}

// src row -> dst row
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

struct SimAcc {
    float2 d0, d1, a0, a1, b0, b1;
};

template <int DT, int N>
__device__ __forceinline__ void sim_piece(const char* __restrict__ pr, const char* __restrict__ cr, SimAcc& s) {
    uint4 p[N], c[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
        c[q] = ld_stream16(cr + q * 512);
        p[q] = ld_stream16(pr + q * 512);
    }
#pragma unroll
    for (int q = 0; q < N; ++q) {
        if (q & 1) acc_pair2<DT>(p[q], c[q], s.d1, s.a1, s.b1);
        else acc_pair2<DT>(p[q], c[q], s.d0, s.a0, s.b0);
    }
}

// cosine similarity of rows pr (predecessor) and cr with the reference's rounding chain (main.py:345-349).  The three
// row sums are float32 (ATen accumulates the reductions in float32; their order is torch's own and unknown, which the
// oracle brackets), split over four independent chains per lane.
template <int DT>
__device__ __forceinline__ float row_similarity(const char* __restrict__ pr, const char* __restrict__ cr, int nvec, int lane) {
    SimAcc s;
    s.d0 = s.d1 = s.a0 = s.a1 = s.b0 = s.b1 = make_float2(0.f, 0.f);
    int v = 0;
    pr += lane * 16;
    cr += lane * 16;
#pragma unroll 1
    for (; v + 256 <= nvec; v += 256) sim_piece<DT, 8>(pr + (int64_t)v * 16, cr + (int64_t)v * 16, s);
    if (v + 128 <= nvec) { sim_piece<DT, 4>(pr + (int64_t)v * 16, cr + (int64_t)v * 16, s); v += 128; }
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
    for (; v + 256 <= nvec; v += 256) run_piece<DT, 8>(hidden, row_bytes, lo + (int64_t)v * 16, w, first, dv, orow);
    if (v + 128 <= nvec) { run_piece<DT, 4>(hidden, row_bytes, lo + (int64_t)v * 16, w, first, dv, orow); v += 128; }
    if (v + 64 <= nvec) { run_piece<DT, 2>(hidden, row_bytes, lo + (int64_t)v * 16, w, first, dv, orow); v += 64; }
    if (v + 32 <= nvec) { run_piece<DT, 1>(hidden, row_bytes, lo + (int64_t)v * 16, w, first, dv, orow); v += 32; }
    // (the shuffles inside a piece are warp-wide: the partial piece is walked by every lane, out-of-range lanes re-read
    // their last full vector and store nothing)
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
}

template <int DT>
__global__ void __launch_bounds__(FU_WARPS * 32, FU_MIN_CTAS)
k_fused_merge(const __grid_constant__ FusedArgs a, const __grid_constant__ AuxPack aux) {
    __shared__ int s_ticket[2];
    __shared__ int s_kept[FU_WARPS];
    __shared__ int s_excl;
    __shared__ unsigned s_mask;
    pdl_enter();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int W = FU_WARPS;
    unsigned long long* const ticket = a.desc;
    unsigned long long* const round_word = a.desc + 1;
    unsigned* const tile_excl = reinterpret_cast<unsigned*>(a.desc + 1 + a.nrounds);
    const int ntiles = a.ntiles, nvec = a.nvec, lag = a.lag;
    const int64_t row_bytes = a.row_bytes;
    const int n_tickets = 2 * ntiles, n_pairs = ntiles - lag;         // lag <= ntiles (host)
    int err = 0;
    int rc = 0, bc = 0;                                     // warp 0: rounds < rc are complete and hold bc kept rows

    if (threadIdx.x == 0) s_ticket[0] = (int)atomicAdd(ticket, 1ull);
    __syncthreads();
#pragma unroll 1
    for (int it = 0;; ++it) {
        const int k = s_ticket[it & 1];
        if (k >= n_tickets) break;
        int next = 0;
        if (threadIdx.x == 0) next = (int)atomicAdd(ticket, 1ull);    // the next tile's ticket travels while this tile is done
        // ticket -> (kind, tile): S(0 .. lag-1), then S(lag + i), G(i) alternating, then the remaining G tiles
        bool is_g;
        int tile;
        if (k < lag) { is_g = false; tile = k; }
        else {
            const int j = k - lag;
            if (j < 2 * n_pairs) { is_g = j & 1; tile = is_g ? (j >> 1) : lag + (j >> 1); }
            else { is_g = true; tile = n_pairs + (j - 2 * n_pairs); }
        }
        const int r = tile * W + wid;
        const bool valid = r < a.S;
        int2 lk = make_int2(-2, -2);
        if (valid) lk = __ldg(a.link + r);

        if (!is_g) {
            // ---- S tile
            int kept = 0;
            if (valid) {
                float s = -2.0f;                            // IGNORE_TOKEN at chain heads (main.py:225-238)
                if (lk.x >= 0)
                    s = row_similarity<DT>(a.hidden + (int64_t)lk.x * row_bytes, a.hidden + (int64_t)r * row_bytes, nvec, lane);
                kept = !(lk.x >= 0 && s >= a.thr);          // NaN compares false: kept
                if (lane == 0) a.sim_seq[r] = s;
            }
            if (lane == 0) s_kept[wid] = kept;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned m = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) m |= (unsigned)s_kept[w] << w;
                st_relaxed32(a.tile_mask + tile, m);
                __threadfence();                            // the mask is visible before the round learns of the tile
                atomicAdd(round_word + tile / FU_ROUND_TILES, (1ull << 32) | (unsigned long long)__popc(m));
            }
        } else {
            // ---- G tile
            if (wid == 0) {
                const int rt = tile / FU_ROUND_TILES;
                int add = 0;
#pragma unroll 1
                for (int r0 = rc; r0 <= rt; r0 += 32) {
                    const int rr = r0 + lane;
                    if (rr <= rt) {
                        const unsigned need = (unsigned)min(FU_ROUND_TILES, ntiles - rr * FU_ROUND_TILES);
                        unsigned long long w = ld_relaxed64(round_word + rr);
                        int spins = 0;
                        while ((unsigned)(w >> 32) != need) {
                            if (++spins > FU_SPIN_LIMIT) { err = 1; break; }
                            __nanosleep(100);
                            w = ld_relaxed64(round_word + rr);
                        }
                        if (rr < rt) add += (int)(unsigned)w;
                    }
                }
                bc += warp_sum_int(add);
                rc = rt;
                __threadfence();                            // acquire: the masks of the complete rounds
                int within = 0;
                for (int t2 = rt * FU_ROUND_TILES + lane; t2 < tile; t2 += 32) within += __popc(ld_relaxed32(a.tile_mask + t2));
                within = warp_sum_int(within);
                const unsigned m = ld_relaxed32(a.tile_mask + tile);
                if (lane == 0) {
                    s_excl = bc + within;
                    s_mask = m;
                    st_relaxed32(tile_excl + tile, (unsigned)(bc + within) + 1u);
                    if (tile == ntiles - 1) fused_finish(a, (long long)(bc + within) + __popc(m));
                }
            }
            __syncthreads();
            const int excl = s_excl;
            const unsigned mask = s_mask;
            if (valid) {
                const bool is_kept = mask >> wid & 1u;
                const int d_r = is_kept ? excl + __popc(mask & ((1u << wid) - 1u)) : -1;
                if (lane == 0) a.dst[r] = d_r;
                int start = -1;                             // where the walk back starts
                bool self = false;
                if (is_kept) {
                    if (lk.x >= 0) start = lk.x;
                    else if (lane == 0) {
                        a.link_next[d_r].x = lk.x;          // chain head / not a chain row
                        if (lk.x == -2) a.link_next[d_r].y = -2;
                    }
                    if (lk.x == -2 || lk.y < 0) {
                        self = true;
                        if (lk.x != -2 && lane == 0) a.link_next[d_r].y = -1;
                    }
                } else if (lk.x >= 0 && lk.y < 0) {
                    start = r;                              // merged away, and the chain ends here
                }
                if (start >= 0) {
                    // every lane walks (uniform loads); lane k remembers the k-th member from the end
                    int x = start, L = 0, mine = -1;
                    unsigned mx;
                    while (true) {
                        mx = (x / W == tile) ? mask : ld_relaxed32(a.tile_mask + x / W);
                        if (mx >> (x % W) & 1u) break;      // kept: the anchor
                        if (lane == L) mine = x;
                        ++L;
                        x = __ldg(&a.link[x].x);
                        if (x < 0) break;                   // (cannot happen: a chain head is never merged away)
                    }
                    if (x >= 0) {
                        const int ex = (x / W == tile) ? excl : wait_excl(tile_excl, x / W, &err);
                        if (ex >= 0) {
                            const int d_a = ex + __popc(mx & ((1u << (x % W)) - 1u));
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
        }
        if (threadIdx.x == 0) s_ticket[(it + 1) & 1] = next;
        __syncthreads();
    }
    // leave the other bank's ticket, round words and prefixes zeroed for the next call of the prefill
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.desc_words; i += (int64_t)gridDim.x * blockDim.x)
        a.desc_clr[i] = 0ull;
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
