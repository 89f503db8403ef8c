// Single-pass fused merge kernel: similarity + threshold select + run merge + compaction in ONE sweep over
// hidden_states (main.py:104-138 for the threshold branch, which is every merge call of a prefill but
// possibly the last).
//
// Why one pass is possible.  In the threshold branch a token's fate depends only on its own similarity
// (sim >= T(slb), main.py:113), and the compacted position of a kept token is the number of kept tokens
// before it in sequence order — a prefix sum.  So the sweep runs over tiles of FUSED_ROWS consecutive
// sequence rows, one warp per row, tiles handed out by an atomic ticket so that tile order is start order:
//
//   phase 1  the warp of row i loads row i and its chain successor succ(i) (16-byte loads, both rows stay in
//            registers), computes sim(i, succ) with the reference's rounding chain and PUBLISHES the flag of
//            succ(i) (one byte, tagged with the call's epoch) plus the similarity.
//   phase 2  it reads its own flag — published by the warp of its predecessor, which sits in an earlier tile —
//            and the tile's kept-row count goes through a decoupled look-back (single-pass prefix scan over
//            tiles) to give every kept row its output position.
//   phase 3  a kept row whose successor is flagged is an anchor: it adds the successor (already in registers),
//            then walks the chain, recomputing each next similarity itself with the same code and the same
//            summation order (so it agrees bit for bit with the flag the owner of that pair publishes), adding
//            members one at a time with a rounding to T per add, and finally divides by T(L+1).  The row and its
//            cos / sin / patch_type / position-id entries are written straight to the compacted position.
//
// Every wait is on a tile with a smaller ticket (already running, and phase 1 never waits), so the sweep is
// deadlock free.  HBM traffic is one read of hidden + one write of the kept rows: a row is loaded again as
// "successor" or as "run member" by other warps, but within a few MB of its first touch, i.e. from L2.
//
// The branch decision needs the global count (main.py:114-116), known only at the end, so the kernel
// SPECULATES on the threshold branch: the last tile checks r = count / n_vis < bound and otherwise reports
// FF_ST_ERROR = 3; the host then redoes the call with the generic multi-kernel path (top-k branch, at most
// once per prefill).  The input is never modified, so the redo sees the original rows.
//
// Chain links live in the index space of the PREVIOUS call: link[i] = (has_pred << 31) | term, where term is
// the old index of the next kept token of the chain; `map` (the previous call's destination array) translates
// it.  That removes a separate remap kernel from the critical path.
#pragma once
#include "ff_common.cuh"
#include "ff_links.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FUSED_ROWS = 4;                      // rows (= warps) per tile
constexpr int FUSED_MIN_ROWS = FUSED_ROWS;
constexpr int FUSED_THREADS = FUSED_ROWS * 32;
constexpr int FUSED_MAX_VPL = 16;                  // 16-byte vectors per lane: rows up to 8 KB

struct FusedArgs {
    const char* hidden;
    char* out;
    int S;
    int nvec;                                      // 16-byte vectors per row
    int64_t row_bytes;
    const int* link;
    const int* map;                                // previous call's dst[] or null (links already in this index space)
    int* link_next;
    uint8_t* state;                                // per sequence row: (tag << 1) | flag
    float* sim_seq;                                // per sequence row: similarity with the chain predecessor
    int* dst;
    unsigned long long* tiles;                     // look-back descriptors
    int64_t* counters;
    int64_t* counters_next;
    int64_t* status;
    float thr;
    double bound;
    unsigned tag;                                  // 1..127
    unsigned long long epoch;                      // 1..65535, in bits 48..63 of a descriptor
    int n_tiles;
    AuxPack aux;
};

__device__ __forceinline__ void aux_copy_row(const AuxPack& p, int i, int d, int lane) {
    for (int q = 0; q < p.n; ++q) {
        const ff_aux& a = p.a[q];
        const int64_t rb = a.row_bytes;
        for (int64_t pl = 0; pl < a.planes; ++pl) {
            const char* s = (const char*)a.src + pl * a.src_plane_stride + (int64_t)i * rb;
            char* o = (char*)a.dst + pl * a.dst_plane_stride + (int64_t)d * rb;
            if (((rb | (int64_t)(uintptr_t)s | (int64_t)(uintptr_t)o) & 15) == 0) {
                for (int64_t v = lane; v < rb / 16; v += 32)
                    reinterpret_cast<uint4*>(o)[v] = __ldg(reinterpret_cast<const uint4*>(s) + v);
            } else if (((rb | (int64_t)(uintptr_t)s | (int64_t)(uintptr_t)o) & 7) == 0) {
                for (int64_t v = lane; v < rb / 8; v += 32)
                    reinterpret_cast<uint2*>(o)[v] = __ldg(reinterpret_cast<const uint2*>(s) + v);
            } else {
                for (int64_t v = lane; v < rb; v += 32) o[v] = s[v];
            }
        }
    }
}

template <int DT>
__device__ __forceinline__ uint4 add_round(const uint4& a, const uint4& b) {      // T(a + b), elementwise
    float x[Num<DT>::EPV], y[Num<DT>::EPV];
    Num<DT>::unpack(a, x);
    Num<DT>::unpack(b, y);
#pragma unroll
    for (int e = 0; e < Num<DT>::EPV; ++e) x[e] = x[e] + y[e];
    return Num<DT>::pack(x);
}

template <int DT>
__device__ __forceinline__ uint4 div_round(const uint4& a, float div) {           // T(a / div)
    float x[Num<DT>::EPV];
    Num<DT>::unpack(a, x);
#pragma unroll
    for (int e = 0; e < Num<DT>::EPV; ++e) x[e] = x[e] / div;
    return Num<DT>::pack(x);
}

constexpr unsigned long long DESC_AGG = 1ull << 32, DESC_INCL = 2ull << 32;

template <int DT, int VPL>
__global__ void __launch_bounds__(FUSED_THREADS, (VPL <= 4 ? 8 : (VPL <= 8 ? 4 : 3)))
k_fused_merge(const FusedArgs a) {
    __shared__ int s_tile;
    __shared__ int s_keep[FUSED_ROWS];
    __shared__ int s_hit[FUSED_ROWS];
    __shared__ int s_excl;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint4 zero = make_uint4(0, 0, 0, 0);

    for (;;) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd((unsigned long long*)&a.counters[C_TICKET], 1ull);
        __syncthreads();
        const int t = s_tile;
        if (t >= a.n_tiles) break;
        const int i = t * FUSED_ROWS + w;
        const bool valid = i < a.S;

        // ---------------- phase 1: own row + successor row, sim(i, succ), publish succ's flag
        uint4 A[VPL], B[VPL];
        int succ = -1;
        bool haspred = false;
        int hit = 0;
        if (valid) {
            const uint32_t l = (uint32_t)__ldg(a.link + i);
            haspred = (l >> 31) != 0;
            const uint32_t s = l & LINK_NONE;
            if (s != LINK_NONE) succ = a.map ? __ldg(a.map + s) : (int)s;
            const char* ra = a.hidden + (int64_t)i * a.row_bytes;
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int v = k * 32 + lane;
                A[k] = v < a.nvec ? ld_stream16(ra + (int64_t)v * 16) : zero;
            }
            if (succ >= 0) {
                const char* rb = a.hidden + (int64_t)succ * a.row_bytes;
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    const int v = k * 32 + lane;
                    B[k] = v < a.nvec ? ld_stream16(rb + (int64_t)v * 16) : zero;
                }
                float dot = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
                for (int k = 0; k < VPL; ++k) acc_pair<DT>(A[k], B[k], dot, na, nb);
                dot = warp_sum(dot);
                na = warp_sum(na);
                nb = warp_sum(nb);
                const float sim = finish_cosine<DT>(dot, na, nb);
                hit = sim >= a.thr;
                if (lane == 0) {
                    a.sim_seq[succ] = sim;
                    *(volatile uint8_t*)(a.state + succ) = (uint8_t)((a.tag << 1) | (unsigned)hit);
                }
            }
        }

        // ---------------- phase 2: own flag (from the predecessor's warp), tile count, look-back
        int flag_i = 0;
        if (valid && haspred) {
            if (lane == 0) {
                unsigned b;
                while ((((b = *(volatile uint8_t*)(a.state + i))) >> 1) != a.tag) __nanosleep(32);
                flag_i = (int)(b & 1u);
            }
            flag_i = __shfl_sync(FULL, flag_i, 0);
        }
        const int keep = valid && !flag_i;
        if (lane == 0) { s_keep[w] = keep; s_hit[w] = hit; }
        __syncthreads();
        if (w == 0) {
            int agg = 0, hits = 0;
#pragma unroll
            for (int r = 0; r < FUSED_ROWS; ++r) { agg += s_keep[r]; hits += s_hit[r]; }
            if (lane == 0) {
                if (hits) atomicAdd((unsigned long long*)&a.counters[C_COUNT], (unsigned long long)hits);
                __threadfence();
                if (t > 0) *(volatile unsigned long long*)(a.tiles + t) = (a.epoch << 48) | DESC_AGG | (unsigned)agg;
            }
            int excl = 0;
            int look = t - 1;
            while (look >= 0) {
                const int idx = look - lane;
                unsigned long long d;
                int ok;
                do {
                    d = idx >= 0 ? *(volatile unsigned long long*)(a.tiles + idx) : ((a.epoch << 48) | DESC_INCL);
                    ok = ((d >> 48) == a.epoch) && ((d >> 32) & 3ull) != 0;
                } while (!__all_sync(FULL, ok));
                const unsigned incl_mask = __ballot_sync(FULL, ((d >> 32) & 3ull) == 2ull);
                int v = (int)(unsigned)(d & 0xffffffffull);
                if (incl_mask) {
                    const int first = __ffs(incl_mask) - 1;
                    v = lane <= first ? v : 0;
                    excl += warp_sum_int(v);
                    break;
                }
                excl += warp_sum_int(v);
                look -= 32;
            }
            if (lane == 0) {
                const int incl = excl + agg;
                __threadfence();
                *(volatile unsigned long long*)(a.tiles + t) = (a.epoch << 48) | DESC_INCL | (unsigned)incl;
                s_excl = excl;
                if (t == a.n_tiles - 1) {
                    // every tile has published: count is final.  Decide the branch (main.py:112-127).
                    __threadfence();
                    const long long count = *(volatile long long*)&a.counters[C_COUNT];
                    const long long n_vis = a.counters[C_NVIS], N = a.counters[C_N];
                    int err = 0;
                    if (n_vis == 0) err = 1;
                    else if (!((double)count / (double)n_vis < a.bound)) err = 3;   // top-k branch: not ours
                    a.counters[C_SKEEP] = incl;
                    a.counters[C_NMERGED] = count;
                    a.counters_next[C_N] = N - count;
                    a.counters_next[C_NVIS] = n_vis - count;
                    a.counters_next[C_COUNT] = 0;
                    a.counters_next[C_TICKET] = 0;
                    a.counters_next[C_TICKET2] = 0;
                    a.status[FF_ST_SEQ_KEEP] = incl;
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
        }
        __syncthreads();
        int d = s_excl;
        for (int r = 0; r < w; ++r) d += s_keep[r];
        if (valid && lane == 0) a.dst[i] = keep ? d : -1;

        // ---------------- phase 3: kept rows: merge the run that follows, write to the compacted position
        if (keep) {
            char* orow = a.out + (int64_t)d * a.row_bytes;
            uint32_t term = LINK_NONE;
            if (succ >= 0) {
                if (!hit) {
                    term = (uint32_t)succ;
                } else {
                    int L = 1;
#pragma unroll
                    for (int k = 0; k < VPL; ++k) A[k] = add_round<DT>(A[k], B[k]);
                    int cur = succ;
                    for (;;) {
                        const uint32_t l = (uint32_t)__ldg(a.link + cur);
                        const uint32_t s = l & LINK_NONE;
                        if (s == LINK_NONE) break;
                        const int nxt = a.map ? __ldg(a.map + s) : (int)s;
                        const char* rn = a.hidden + (int64_t)nxt * a.row_bytes;
                        float dot = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
                        for (int k = 0; k < VPL; ++k) {
                            const int v = k * 32 + lane;
                            const uint4 x = v < a.nvec ? ldg16(rn + (int64_t)v * 16) : zero;
                            acc_pair<DT>(B[k], x, dot, na, nb);
                        }
                        dot = warp_sum(dot);
                        na = warp_sum(na);
                        nb = warp_sum(nb);
                        const float sim = finish_cosine<DT>(dot, na, nb);
                        if (!(sim >= a.thr)) { term = (uint32_t)nxt; break; }
#pragma unroll
                        for (int k = 0; k < VPL; ++k) {
                            const int v = k * 32 + lane;
                            const uint4 x = v < a.nvec ? ldg16(rn + (int64_t)v * 16) : zero;
                            A[k] = add_round<DT>(A[k], x);
                            B[k] = x;
                        }
                        ++L;
                        cur = nxt;
                    }
                    const float div = Num<DT>::rnd((float)(L + 1));
#pragma unroll
                    for (int k = 0; k < VPL; ++k) A[k] = div_round<DT>(A[k], div);
                }
            }
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int v = k * 32 + lane;
                if (v < a.nvec) st_stream16(orow + (int64_t)v * 16, A[k]);
            }
            aux_copy_row(a.aux, i, d, lane);
            if (lane == 0) a.link_next[d] = (int)((haspred ? LINK_HASPRED : 0u) | term);
        }
    }
}

template <int DT, int VPL>
inline int launch_fused_t(int sm_count, const FusedArgs& a, cudaStream_t st) {
    static int occ_cache = 0;
    if (occ_cache == 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused_merge<DT, VPL>, FUSED_THREADS, 0) != cudaSuccess || occ < 1)
            occ = 1;
        occ_cache = occ;
    }
    int grid = sm_count * occ_cache;
    if (grid > a.n_tiles) grid = a.n_tiles;
    k_fused_merge<DT, VPL><<<grid, FUSED_THREADS, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? FF_OK : FF_E_CUDA;
}

template <int DT>
inline int launch_fused_dt(int sm_count, const FusedArgs& a, int vpl, cudaStream_t st) {
    if (vpl <= 2) return launch_fused_t<DT, 2>(sm_count, a, st);
    if (vpl <= 4) return launch_fused_t<DT, 4>(sm_count, a, st);
    if (vpl <= 8) return launch_fused_t<DT, 8>(sm_count, a, st);
    if (vpl <= 12) return launch_fused_t<DT, 12>(sm_count, a, st);
    if (vpl <= 14) return launch_fused_t<DT, 14>(sm_count, a, st);
    return launch_fused_t<DT, 16>(sm_count, a, st);
}

// Returns FF_E_UNSUPPORTED when the shape is outside the fused kernel (caller falls back to the generic path).
inline int launch_fused(int sm_count, int dtype, int64_t S, int64_t H, double thr, FusedArgs a, cudaStream_t st) {
    const int64_t eb = dtype == FF_F32 ? 4 : 2;
    const int64_t row_bytes = H * eb;
    if (row_bytes % 16 != 0 || (((uintptr_t)a.hidden | (uintptr_t)a.out) & 15) != 0) return FF_E_UNSUPPORTED;
    const int64_t nvec = row_bytes / 16;
    const int vpl = (int)((nvec + 31) / 32);
    if (vpl > FUSED_MAX_VPL) return FF_E_UNSUPPORTED;
    if (!(thr > -2.0)) return FF_E_UNSUPPORTED;            // chain heads (sim = -2) must never be flagged
    a.S = (int)S;
    a.nvec = (int)nvec;
    a.row_bytes = row_bytes;
    a.thr = (float)thr;
    a.n_tiles = (int)((S + FUSED_ROWS - 1) / FUSED_ROWS);
    switch (dtype) {
        case FF_BF16: return launch_fused_dt<FF_BF16>(sm_count, a, vpl, st);
        case FF_F16: return launch_fused_dt<FF_F16>(sm_count, a, vpl, st);
        case FF_F32: return launch_fused_dt<FF_F32>(sm_count, a, vpl, st);
    }
    return FF_E_UNSUPPORTED;
}

}  // namespace ff
