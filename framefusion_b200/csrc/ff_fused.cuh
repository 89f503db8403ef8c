// Read-once merge kernel: similarity + threshold select + run merge + compaction of hidden_states and of the aux
// tensors (cos / sin / patch_type / position ids) in ONE launch that fetches every row of hidden_states from HBM
// once (main.py:104-138, threshold branch — every merge call of a prefill but possibly the last).
//
// Shape of the problem.  A token is compared with the previous surviving token of the SAME patch id (its chain
// predecessor, main.py:216-238) — 576 rows (4 MB) back on the first call of a uniform video — while the output is
// compacted in SEQUENCE order (main.py:132-138): a row's destination is the number of kept rows before it, known only
// when every earlier row has been compared.  So every row is visited twice by the warps of one persistent grid — once
// for its similarity (from HBM), once for the gather (from the L2) — and the warps are independent workers that take
// units of work from ONE ticket:
//
//   S unit (row r).  The sequence is cut into bands of B rows.  If r is the first row of its chain inside its band, the
//     warp walks the chain through the band: the row arrives from HBM in registers and stays in the warp's shared-memory
//     slot as the next step's predecessor, so that a row is fetched once although it takes part in two similarities
//     (the head of a segment fetches its predecessor — a row of the previous band — from the L2).  Three row sums, the
//     reference's rounding chain, sim >= thr -> merged away; kept rows set their bit in mask[r / 32].  At the end of the
//     segment the band receives ONE release-atomic: (rows << 32) | kept rows.  Any other row: nothing to do.  An S unit
//     waits for nothing.
//   G unit (row r), `lag` rows behind the S units.  It waits until its band is complete and the band's exclusive prefix
//     is published (by the G unit of the band's first row, from the previous band's prefix and count) — then every flag
//     of every earlier row is known.  A kept row ends the run of its chain predecessor: the warp walks the masks back to
//     the run's anchor and writes the anchor's destination row: the raw row if the run has no members, else
//     T(T(..T(anchor + m1) + ..) + mL) / T(L+1), one rounding to T per add in chain order (the sequence torch-CPU
//     index_add_ performs, main.py:304-311) and one division (main.py:314-317).  Chain tails end their own run, rows
//     outside the chains are copied.  The aux rows and the (pred, succ) links of the next call go with it.  All of these
//     rows were read by S units at most `lag` + a run's length ago: they come out of the L2.
//
// Tickets are handed out in a fixed order — S(0 .. lag-1), then 32 S units and 32 G units alternating — so every
// wait is for work with a smaller ticket, held by a warp that is running: no deadlock whatever is resident; every spin is
// bounded all the same (FF_ST_INTERNAL).
//
// The branch decision (main.py:114-116) needs the global count, known only at the end: the kernel speculates on the
// threshold branch, the G unit of the last row checks count / n_vis < bound and otherwise reports FF_ST_ERROR = 3; the
// host then redoes the call with the multi-kernel path (top-k branch, at most once per prefill).  The input is never
// modified, so the redo sees the original rows.
#pragma once
#include "ff_common.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FU_WARPS = 8;                        // warps per CTA (independent workers; a CTA only shares its shared memory)
constexpr int FU_MIN_CTAS = 2;                     // per SM: up to 128 registers per thread
constexpr int FU_BAND = 2048;                      // rows per band (multiple of 32; FF_FUSED_BAND)
constexpr int FU_LAG = 4096;                       // S units run this many rows ahead of the G units (>= band; FF_FUSED_LAG)
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
    int S, nvec, row_bytes, slot_vecs, band, nbands, lag, n_tickets;
    const int2* link;                              // [S] (pred, succ): row index, -1 = chain head / tail, -2 = not a chain row
    int2* link_next;                               // [S_keep] the same for the compacted sequence
    unsigned long long* desc;                      // zero on entry, desc_words u64 in all: the ticket, band words [nbands]
                                                   // ((rows done << 32) | kept rows), then u32 band_excl [nbands] (exclusive
                                                   // prefix + 1) and u32 mask [ceil(S / 32)] (bit = kept)
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
__device__ __forceinline__ unsigned long long ld_acquire64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_or32(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// the arrays behind the ticket
struct FusedDesc {
    unsigned long long* ticket;
    unsigned long long* band_word;                          // [nbands] (rows done << 32) | kept rows
    unsigned* band_excl;                                    // [nbands] kept rows before the band + 1
    unsigned* mask;                                         // [ceil(S / 32)] bit = kept
    int S, band;
    __device__ __forceinline__ int band_rows(int b) const { return min(band, S - b * band); }
};

// every row of band b has been compared (the S units that do it hold smaller tickets); the kept rows of the band, -1 after a time-out
__device__ __forceinline__ int wait_band_done(const FusedDesc& d, int b, int* err) {
    const unsigned need = (unsigned)d.band_rows(b);
    unsigned long long w = ld_acquire64(d.band_word + b);
    int spins = 0;
    while ((unsigned)(w >> 32) != need) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; return -1; }
        __nanosleep(100);
        w = ld_acquire64(d.band_word + b);
    }
    return (int)(unsigned)w;
}
// kept rows before band b (published by the G unit of the band's first row, which holds a smaller ticket); -1 after a time-out
__device__ __forceinline__ int wait_band_excl(const FusedDesc& d, int b, int* err) {
    unsigned v = ld_acquire32(d.band_excl + b);
    int spins = 0;
    while (v == 0u) {
        if (++spins > FU_SPIN_LIMIT) { *err = 1; break; }
        __nanosleep(100);
        v = ld_acquire32(d.band_excl + b);
    }
    return (int)v - 1;
}
// What a warp remembers between its G units: the band it last worked in, the kept rows before that band and before the
// band in front of it.  Bands complete in order for a warp's purposes (it enters band b only when all bands up to b are
// complete), so the prefix of the next band follows from the cached one and the band word — nobody waits for a publisher.
struct BandCache {
    int cb, ce, pe;                                         // band, kept rows before it, kept rows before band cb - 1
};

// kept rows before row x of a complete band (whole warp); *mx = the mask word of x
__device__ __forceinline__ int prefix_at(const FusedDesc& d, const BandCache& bc, int x, int lane, unsigned* mx, int* err) {
    const int b = x / d.band, w0 = b * (d.band >> 5), wx = x >> 5;
    int cnt = 0;
    for (int i = w0 + lane; i < wx; i += 32) cnt += __popc(ld_relaxed32(d.mask + i));
    const unsigned m = ld_relaxed32(d.mask + wx);
    const int e = b == bc.cb ? bc.ce : (b == bc.cb - 1 ? bc.pe : wait_band_excl(d, b, err));
    cnt = warp_sum_int(cnt);
    *mx = m;
    return e < 0 ? -1 : e + cnt + __popc(m & ((1u << (x & 31)) - 1u));
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

// ---- the walk of an S unit: the current row against the row in the warp's shared-memory slot, which it then replaces
// dot += T(p * c), nb += c * c on one 16-byte vector pair (cf. acc_pair2: the same arithmetic without the sum of p * p)
template <int DT>
__device__ __forceinline__ void acc_dot_nb(const uint4& vp, const uint4& vc, float2& dot, float2& nb) {
    if (DT == FF_BF16) {
        const uint32_t pw[4] = {vp.x, vp.y, vp.z, vp.w}, cw[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __nv_bfloat162 pp = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&pw[q]), *reinterpret_cast<const __nv_bfloat162*>(&cw[q]));
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

struct StepAcc {
    float2 d0, d1, b0, b1;
};

// N vectors per lane of the current row: loads back to back, then per vector the slot's vector (predecessor), the sums,
// and the current vector takes its place.  DOT = false: no predecessor (the row is only parked and its norm taken).
template <int DT, int N, bool DOT>
__device__ __forceinline__ void step_piece(const char* __restrict__ cr, uint4* __restrict__ sl, StepAcc& s) {
    uint4 c[N];
#pragma unroll
    for (int q = 0; q < N; ++q) c[q] = ld_stream16(cr + q * 512);
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const uint4 p = DOT ? sl[q * 32] : c[q];
        if (q & 1) acc_dot_nb<DT>(p, c[q], s.d1, s.b1);
        else acc_dot_nb<DT>(p, c[q], s.d0, s.b0);
        sl[q * 32] = c[q];
    }
}

// row cr against the slot; returns (dot, |cr|^2) summed over the warp
template <int DT, bool DOT>
__device__ __forceinline__ float2 row_step(const char* __restrict__ cr, uint4* __restrict__ slot, int nvec, int lane) {
    StepAcc s;
    s.d0 = s.d1 = s.b0 = s.b1 = make_float2(0.f, 0.f);
    int v = 0;
    cr += lane * 16;
    uint4* sl = slot + lane;
#pragma unroll 1
    for (; v + 256 <= nvec; v += 256) step_piece<DT, 8, DOT>(cr + (int64_t)v * 16, sl + v, s);
    if (v + 128 <= nvec) { step_piece<DT, 4, DOT>(cr + (int64_t)v * 16, sl + v, s); v += 128; }
    if (v + 64 <= nvec) { step_piece<DT, 2, DOT>(cr + (int64_t)v * 16, sl + v, s); v += 64; }
    if (v + 32 <= nvec) { step_piece<DT, 1, DOT>(cr + (int64_t)v * 16, sl + v, s); v += 32; }
    if (v + lane < nvec) step_piece<DT, 1, DOT>(cr + (int64_t)v * 16, sl + v, s);
    float2 r;
    r.x = DOT ? warp_sum((s.d0.x + s.d0.y) + (s.d1.x + s.d1.y)) : 0.f;
    r.y = warp_sum((s.b0.x + s.b0.y) + (s.b1.x + s.b1.y));
    return r;
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

// S unit of row r (see the head of the file)
template <int DT>
__device__ __forceinline__ void s_unit(const FusedArgs& a, const FusedDesc& d, int r, uint4* slot, int lane) {
    int2 lk = __ldg(a.link + r);
    const int b = r / a.band, band_end = min(b * a.band + a.band, a.S);
    if (lk.x >= b * a.band) return;                         // its chain predecessor walks the band
    int rows = 0, kept_rows = 0;
    if (lk.x == -2) {                                       // outside the chains: kept (main.py:132)
        if (lane == 0) { a.sim_seq[r] = -2.0f; red_or32(d.mask + (r >> 5), 1u << (r & 31)); }
        rows = kept_rows = 1;
    } else {
        const int nvec = a.nvec;
        const int64_t row_bytes = a.row_bytes;
        float na = 0.f;
        bool have_prev = lk.x >= 0;
        if (have_prev) na = row_step<DT, false>(a.hidden + (int64_t)lk.x * row_bytes, slot, nvec, lane).y;
        int cur = r;
#pragma unroll 1
        while (true) {
            float s = -2.0f;                                // IGNORE_TOKEN at chain heads (main.py:225-238)
            float2 t;
            if (have_prev) {
                t = row_step<DT, true>(a.hidden + (int64_t)cur * row_bytes, slot, nvec, lane);
                s = finish_cosine<DT>(t.x, na, t.y);
            } else {
                t = row_step<DT, false>(a.hidden + (int64_t)cur * row_bytes, slot, nvec, lane);
            }
            const int kept = !(have_prev && s >= a.thr);    // NaN compares false: kept
            if (lane == 0) {
                a.sim_seq[cur] = s;
                if (kept) red_or32(d.mask + (cur >> 5), 1u << (cur & 31));
            }
            ++rows;
            kept_rows += kept;
            na = t.y;
            have_prev = true;
            const int nxt = lk.y;
            if (nxt < 0 || nxt >= band_end) break;
            cur = nxt;
            lk = __ldg(a.link + cur);
        }
    }
    if (lane == 0) red_release_add64(d.band_word + b, ((unsigned long long)rows << 32) | (unsigned long long)kept_rows);
}

// G unit of row r (see the head of the file)
template <int DT>
__device__ __forceinline__ void g_unit(const FusedArgs& a, const AuxPack& aux, const FusedDesc& d, BandCache& bc, int r, int lane, int* err) {
    const int2 lk = __ldg(a.link + r);
    const int b = r / a.band;
    const int nvec = a.nvec;
    const int64_t row_bytes = a.row_bytes;
    if (b != bc.cb) {
        // entering a band: every band up to b must be complete; the prefix moves along with the band words
        if (bc.cb < 0) bc.cb = 0;
        while (bc.cb < b) {
            const int kp = wait_band_done(d, bc.cb, err);
            if (kp < 0) return;
            bc.pe = bc.ce;
            bc.ce += kp;
            ++bc.cb;
        }
        if (wait_band_done(d, b, err) < 0) return;
    }
    if (r == b * a.band && lane == 0) st_release32(d.band_excl + b, (unsigned)bc.ce + 1u);   // for walks that end in older bands
    unsigned m;
    const int ex = prefix_at(d, bc, r, lane, &m, err);
    if (ex < 0) return;
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
        while (true) {
            const unsigned mx = ld_relaxed32(d.mask + (x >> 5));
            if (mx >> (x & 31) & 1u) break;                 // kept: the anchor
            if (lane == L) mine = x;
            ++L;
            x = __ldg(&a.link[x].x);
            if (x < 0) break;                               // (cannot happen: a chain head is never merged away)
        }
        if (x >= 0) {
            unsigned mx;
            const int d_a = prefix_at(d, bc, x, lane, &mx, err);
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
    extern __shared__ uint4 fu_slots[];                     // one row per warp
    pdl_enter();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint4* const slot = fu_slots + (size_t)wid * a.slot_vecs;
    FusedDesc d;
    d.ticket = a.desc;
    d.band_word = a.desc + 1;
    d.band_excl = reinterpret_cast<unsigned*>(a.desc + 1 + a.nbands);
    d.mask = d.band_excl + a.nbands;
    d.S = a.S;
    d.band = a.band;
    const int lag = a.lag, n_tickets = a.n_tickets;
    int err = 0;
    BandCache bc;
    bc.cb = -1; bc.ce = 0; bc.pe = 0;
    int k = 0;
    if (lane == 0) k = (int)atomicAdd(d.ticket, 1ull);
    k = __shfl_sync(FULL, k, 0);
#pragma unroll 1
    while (k < n_tickets) {
        int next = 0;
        if (lane == 0) next = (int)atomicAdd(d.ticket, 1ull);         // the next ticket travels while this unit is done
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
            else s_unit<DT>(a, d, r, slot, lane);
        }
        k = __shfl_sync(FULL, next, 0);
    }
    // leave the other bank's ticket, band words, prefixes and masks zeroed for the next call of the prefill
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
