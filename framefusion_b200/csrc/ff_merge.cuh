// Merge + compaction (main.py:285-317 and 132-138, 161-178), generic path: one warp per kept
// sequence row.  A kept row whose by-patch successor positions are flagged is an anchor: it adds the run
// members one at a time in chain order, every add rounded to T (the order torch-CPU index_add_ uses),
// divides once by T(L+1), and writes the result straight to its compacted position — the reference's
// separate index_add_ / div / boolean-mask passes (3 extra trips over the rows) collapse into one.
#pragma once
#include "ff_common.cuh"

namespace ff {

// run length starting after by-patch position j (flags are contiguous by-patch, chain boundaries are not
// special: the reference's run detection works on the flat by-patch array, main.py:276)
__device__ __forceinline__ int run_after(const uint8_t* __restrict__ flag, int j, int N) {
    int m = j + 1;
    while (m < N && flag[m]) ++m;
    return m - j - 1;
}

template <int DT, bool VEC, bool INPLACE>
__device__ __forceinline__ void merge_row(const void* hidden, void* out_row, const char* src_row, int H, int lane,
                                          const int* __restrict__ order, int j, int L, int wrapL) {
    typedef typename Num<DT>::store_t st;
    const int64_t row_bytes = (int64_t)H * sizeof(st);
    const float div = Num<DT>::rnd((float)(L + wrapL + 1));   // the divisor tensor is cast to T (main.py:317)
    if (VEC) {
        const int nvec = H / Num<DT>::EPV;
        for (int v = lane; v < nvec; v += 32) {
            float acc[Num<DT>::EPV], x[Num<DT>::EPV];
            Num<DT>::unpack(ldg16(src_row + (int64_t)v * 16), acc);
            for (int m = 1; m <= L; ++m) {
                const char* mr = (const char*)hidden + (int64_t)order[j + m] * row_bytes;
                Num<DT>::unpack(ldg16(mr + (int64_t)v * 16), x);
#pragma unroll
                for (int e = 0; e < Num<DT>::EPV; ++e) acc[e] = Num<DT>::rnd(acc[e] + x[e]);
            }
            for (int m = 0; m < wrapL; ++m) {
                const char* mr = (const char*)hidden + (int64_t)order[m] * row_bytes;
                Num<DT>::unpack(ldg16(mr + (int64_t)v * 16), x);
#pragma unroll
                for (int e = 0; e < Num<DT>::EPV; ++e) acc[e] = Num<DT>::rnd(acc[e] + x[e]);
            }
#pragma unroll
            for (int e = 0; e < Num<DT>::EPV; ++e) acc[e] = Num<DT>::rnd(acc[e] / div);
            *reinterpret_cast<uint4*>((char*)out_row + (int64_t)v * 16) = Num<DT>::pack(acc);
        }
    } else {
        for (int e = lane; e < H; e += 32) {
            float acc = Num<DT>::load(src_row, e);
            for (int m = 1; m <= L; ++m)
                acc = Num<DT>::rnd(acc + Num<DT>::load((const char*)hidden + (int64_t)order[j + m] * row_bytes, e));
            for (int m = 0; m < wrapL; ++m)
                acc = Num<DT>::rnd(acc + Num<DT>::load((const char*)hidden + (int64_t)order[m] * row_bytes, e));
            acc = Num<DT>::rnd(acc / div);
            Num<DT>::store(out_row, e, acc);
        }
    }
}

template <int DT, bool VEC>
__global__ void __launch_bounds__(256)
k_merge_compact(const void* __restrict__ hidden, void* __restrict__ out, int S, int H, const int* __restrict__ rank,
                const int* __restrict__ order, const uint8_t* __restrict__ flag, const int* __restrict__ dst,
                const int64_t* __restrict__ counters, int use_flags) {
    typedef typename Num<DT>::store_t st;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= S) return;
    const int d = dst[i];
    if (d < 0) return;
    const int64_t row_bytes = (int64_t)H * sizeof(st);
    const char* src = (const char*)hidden + (int64_t)i * row_bytes;
    char* orow = (char*)out + (int64_t)d * row_bytes;
    int L = 0, wrapL = 0, j = -1;
    if (use_flags) {
        j = rank[i];
        if (j >= 0) {
            const int N = (int)counters[C_N];
            L = run_after(flag, j, N);
            // a flagged run at by-patch position 0 has anchor index -1, which the reference's advanced
            // indexing wraps to the last by-patch position (main.py:290, 306)
            if (j == N - 1 && N > 1 && flag[0]) wrapL = run_after(flag, -1, N - 1);
        }
    }
    if (L + wrapL == 0) {
        if (VEC) {
            const int nvec = H / Num<DT>::EPV;
            int v = lane;
            for (; v + 96 < nvec; v += 128) {
                uint4 a0 = ld_stream16(src + (int64_t)v * 16), a1 = ld_stream16(src + (int64_t)(v + 32) * 16);
                uint4 a2 = ld_stream16(src + (int64_t)(v + 64) * 16), a3 = ld_stream16(src + (int64_t)(v + 96) * 16);
                st_stream16(orow + (int64_t)v * 16, a0);
                st_stream16(orow + (int64_t)(v + 32) * 16, a1);
                st_stream16(orow + (int64_t)(v + 64) * 16, a2);
                st_stream16(orow + (int64_t)(v + 96) * 16, a3);
            }
            for (; v < nvec; v += 32) st_stream16(orow + (int64_t)v * 16, ld_stream16(src + (int64_t)v * 16));
        } else {
            const st* s = (const st*)src;
            st* o = (st*)orow;
            for (int e = lane; e < H; e += 32) o[e] = s[e];
        }
        return;
    }
    merge_row<DT, VEC, false>(hidden, orow, src, H, lane, order, j, L, wrapL);
}

// ---- aux tensors (cos, sin, patch_type, position ids): row copy src[i] -> dst[dst_row[i]] ----------
struct AuxPack {
    ff_aux a[FF_MAX_AUX];
    int n;
};

// ---- the same, tuned for the two-pass path at full size.  One warp per KEPT row, driven by the 16-byte records the
// scan kernel leaves: (source row, destination row, by-patch position, run length) for the rows of the chains in
// by-patch order from the front of rec[], for the rows outside the chains from its end (merge stage); srcidx[d] in the
// prune stage.  A warp requests its record before anything else, so the row data is one dependent load away and no
// warp is launched for a row that was merged away.  Thin warps: eight 16-byte vectors per lane in flight for a copied
// row; an anchor requests the same four vectors of the anchor and of the next run member together (main.py:304-311
// order of the adds).  The aux rows (cos, sin, patch_type, position ids) are copied by warps of their own (odd
// blocks) with every load of a row in flight at once, so their latency never sits behind a hidden_states row.
constexpr int GATHER_WARPS = 4;                  // small blocks: a slow anchor warp holds up only three others
constexpr int AUX_SLOTS = 8;                     // (tensor, plane) pairs an aux warp keeps in flight

__device__ __noinline__ void gather_aux_rows(const AuxPack& aux, int i, int d, int lane) {
    // fast path: every (tensor, plane) row is at most 32 pieces of 16 or 8 bytes -> one piece per lane, all loads first
    uint4 buf[AUX_SLOTS];
    int k = 0;
    bool fast = true;
#pragma unroll 1
    for (int q = 0; q < aux.n; ++q) {
        const ff_aux& x = aux.a[q];
        const int64_t al = x.row_bytes | (int64_t)(uintptr_t)x.src | (int64_t)(uintptr_t)x.dst | x.src_plane_stride | x.dst_plane_stride;
        const int piece = (al & 15) == 0 ? 16 : ((al & 7) == 0 ? 8 : 0);
        if (piece == 0 || x.row_bytes > 32 * piece) fast = false;
        k += (int)x.planes;
    }
    if (fast && k <= AUX_SLOTS) {
        k = 0;
#pragma unroll 1
        for (int q = 0; q < aux.n; ++q) {
            const ff_aux& x = aux.a[q];
            const bool wide = (x.row_bytes & 15) == 0;      // alignment of the bases was checked above
            for (int64_t pl = 0; pl < x.planes; ++pl, ++k) {
                const char* s = (const char*)x.src + pl * x.src_plane_stride + (int64_t)i * x.row_bytes;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (wide) {
                    if (lane * 16 < x.row_bytes) v = __ldg(reinterpret_cast<const uint4*>(s) + lane);
                } else if (lane * 8 < x.row_bytes) {
                    const uint2 t = __ldg(reinterpret_cast<const uint2*>(s) + lane);
                    v.x = t.x; v.y = t.y;
                }
#pragma unroll
                for (int e = 0; e < AUX_SLOTS; ++e)
                    if (e == k) buf[e] = v;                 // static register indexing
            }
        }
        k = 0;
#pragma unroll 1
        for (int q = 0; q < aux.n; ++q) {
            const ff_aux& x = aux.a[q];
            const bool wide = (x.row_bytes & 15) == 0;
            for (int64_t pl = 0; pl < x.planes; ++pl, ++k) {
                char* o = (char*)x.dst + pl * x.dst_plane_stride + (int64_t)d * x.row_bytes;
                uint4 v = buf[0];
#pragma unroll
                for (int e = 1; e < AUX_SLOTS; ++e)
                    if (e == k) v = buf[e];
                if (wide) {
                    if (lane * 16 < x.row_bytes) reinterpret_cast<uint4*>(o)[lane] = v;
                } else if (lane * 8 < x.row_bytes) {
                    reinterpret_cast<uint2*>(o)[lane] = make_uint2(v.x, v.y);
                }
            }
        }
        return;
    }
#pragma unroll 1
    for (int q = 0; q < aux.n; ++q) {                       // any shape
        const ff_aux& x = aux.a[q];
        for (int64_t pl = 0; pl < x.planes; ++pl) {
            const char* s = (const char*)x.src + pl * x.src_plane_stride + (int64_t)i * x.row_bytes;
            char* o = (char*)x.dst + pl * x.dst_plane_stride + (int64_t)d * x.row_bytes;
            if (((x.row_bytes | (int64_t)(uintptr_t)s | (int64_t)(uintptr_t)o) & 15) == 0) {
                for (int64_t v = lane; v < x.row_bytes / 16; v += 32)
                    reinterpret_cast<uint4*>(o)[v] = __ldg(reinterpret_cast<const uint4*>(s) + v);
            } else if (((x.row_bytes | (int64_t)(uintptr_t)s | (int64_t)(uintptr_t)o) & 7) == 0) {
                for (int64_t v = lane; v < x.row_bytes / 8; v += 32)
                    reinterpret_cast<uint2*>(o)[v] = __ldg(reinterpret_cast<const uint2*>(s) + v);
            } else {
                for (int64_t v = lane; v < x.row_bytes; v += 32) o[v] = s[v];
            }
        }
    }
}

template <int DT>
__device__ __noinline__ void gather_anchor_ieee(const void* hidden, void* out_row, const char* src_row, int nvec, int lane,
                                                const int* __restrict__ order, int j, int L, int wrapL) {
    merge_row<DT, true, false>(hidden, out_row, src_row, nvec * Num<DT>::EPV, lane, order, j, L, wrapL);
}

// One record -> (source row, destination row, by-patch position, run length); false when unit u is past the kept rows.
// The record is requested before the counters it is checked against: one dependent load less before the row data.
__device__ __forceinline__ bool gather_unit(int u, int S, const int* __restrict__ srcidx, const int4* __restrict__ rec,
                                            const int64_t* __restrict__ counters, const int64_t* __restrict__ counters_next,
                                            int4* out) {
    if (u >= S) return false;
    int4 r = rec ? __ldg(rec + u) : make_int4(__ldg(srcidx + u), u, -1, 0);
    if (u >= (int)counters[C_SKEEP]) return false;
    if (rec) {
        const int n_chain = (int)counters_next[C_N];
        if (u >= n_chain) r = __ldg(rec + (S - 1 - (u - n_chain)));
    }
    *out = r;
    return true;
}

// grid: 5 * ceil(ceil(S / GATHER_WARPS) / 4) blocks when there are aux tensors — every fifth block copies the aux rows
// of the sixteen units of its four neighbours, four rows per warp — else ceil(S / GATHER_WARPS)
template <int DT>
__global__ void __launch_bounds__(GATHER_WARPS * 32, 32 / GATHER_WARPS)
k_merge_gather(const void* __restrict__ hidden, void* __restrict__ out, int nvec, int S, const int* __restrict__ srcidx,
               const int4* __restrict__ rec, const int* __restrict__ order, const uint8_t* __restrict__ flag,
               const int64_t* __restrict__ counters, const int64_t* __restrict__ counters_next,
               const __grid_constant__ AuxPack aux) {
    pdl_enter();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (rec && counters[C_NMERGED] == 0) return;            // nothing merged: the host hands the input tensors back
    int4 r;
    if (aux.n > 0 && blockIdx.x % 5 == 4) {
        const int u0 = ((int)blockIdx.x / 5) * (4 * GATHER_WARPS) + wid * 4;
#pragma unroll 1
        for (int e = 0; e < 4; ++e)
            if (gather_unit(u0 + e, S, srcidx, rec, counters, counters_next, &r)) gather_aux_rows(aux, r.x, r.y, lane);
        return;
    }
    const int mb = aux.n > 0 ? (int)blockIdx.x - (int)blockIdx.x / 5 : (int)blockIdx.x;
    if (!gather_unit(mb * GATHER_WARPS + wid, S, srcidx, rec, counters, counters_next, &r)) return;
    const int i = r.x, d = r.y, j = r.z, L = r.w;
    const int64_t row_bytes = (int64_t)nvec * 16;
    const char* src = (const char*)hidden + (int64_t)i * row_bytes;
    char* orow = (char*)out + (int64_t)d * row_bytes;
    int wrapL = 0;
    if (j >= 0) {
        const int N = (int)counters[C_N];
        if (j == N - 1 && N > 1 && flag[0]) wrapL = run_after(flag, -1, N - 1);       // main.py:290 wrap-around
    }
    if (L + wrapL == 0) {
        for (int v0 = lane; v0 < nvec; v0 += 256) {         // eight 16-byte vectors per lane in flight
            uint4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (v0 + 32 * q < nvec) x[q] = ld_stream16(src + (int64_t)(v0 + 32 * q) * 16);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (v0 + 32 * q < nvec) st_stream16(orow + (int64_t)(v0 + 32 * q) * 16, x[q]);
        }
    } else {
        const Divider<DT> dv(L + wrapL + 1);
        if (!dv.by_rcp) {                                   // needs the IEEE division: the generic row routine, out of line
            gather_anchor_ieee<DT>(hidden, orow, src, nvec, lane, order, j, L, wrapL);
            return;
        }
        // the first member's row is addressed once, outside the loop: runs of one merged token are the common case
        const char* mr0 = (const char*)hidden + (int64_t)order[L > 0 ? j + 1 : 0] * row_bytes;
        for (int v0 = lane; v0 < nvec; v0 += 128) {
            uint4 acc[4], x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (v0 + 32 * q < nvec) {
                    acc[q] = ld_stream16(src + (int64_t)(v0 + 32 * q) * 16);
                    x[q] = ldg16(mr0 + (int64_t)(v0 + 32 * q) * 16);
                }
            for (int m = 0; m < L + wrapL; ++m) {
                if (m > 0) {
                    const int jm = m < L ? j + 1 + m : m - L;
                    const char* mr = (const char*)hidden + (int64_t)order[jm] * row_bytes;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (v0 + 32 * q < nvec) x[q] = ldg16(mr + (int64_t)(v0 + 32 * q) * 16);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (v0 + 32 * q < nvec) acc[q] = Num<DT>::add_vec(acc[q], x[q]);      // T(acc + member), main.py:304
            }
            if (dv.pow2) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (v0 + 32 * q < nvec) st_stream16(orow + (int64_t)(v0 + 32 * q) * 16, Num<DT>::scale_vec(acc[q], dv.rcp));
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (v0 + 32 * q < nvec) st_stream16(orow + (int64_t)(v0 + 32 * q) * 16, dv.vec_rcp(acc[q]));
            }
        }
    }
}

// ---- in-place variant behind the reference's static merge_tokens_and_get_mask (main.py:243-319):
// one warp per by-patch position; anchors are rewritten in place, nothing is compacted.
template <int DT, bool VEC>
__global__ void __launch_bounds__(256)
k_merge_inplace(void* hidden, int H, int N, const int* __restrict__ order, const uint8_t* __restrict__ flag) {
    typedef typename Num<DT>::store_t st;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= N || flag[j]) return;
    const int L = run_after(flag, j, N);
    int wrapL = 0;
    if (j == N - 1 && N > 1 && flag[0]) wrapL = run_after(flag, -1, N - 1);
    if (L + wrapL == 0) return;
    char* row = (char*)hidden + (int64_t)order[j] * H * sizeof(st);
    merge_row<DT, VEC, true>(hidden, row, row, H, lane, order, j, L, wrapL);
}

// flags from an explicit index list (static API): flag[merge_index[m]] = 1, keep[order[merge_index[m]]] = 0
__global__ void k_flags_from_index(const int64_t* __restrict__ merge_index, int M, int N, const int64_t* __restrict__ order64,
                                   int* __restrict__ order32, uint8_t* __restrict__ flag, uint8_t* __restrict__ keep) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N) order32[t] = (int)order64[t];
    if (t < M) {
        int64_t j = merge_index[t];
        if (j < 0) j += N;
        if (j >= 0 && j < N) {
            flag[j] = 1;
            keep[order64[j]] = 0;
        }
    }
}

__global__ void __launch_bounds__(256)
k_aux_compact(AuxPack p, int S, const int* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= S) return;
    const int d = dst[i];
    if (d < 0) return;
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

// ---- mask[keep][:, keep]  (main.py:100, 138) ---------------------------------------------------------
template <typename E>
__global__ void k_compact_mask(const E* __restrict__ mask, E* __restrict__ out, int S, int S_keep,
                               const int* __restrict__ srcidx) {
    const int a = blockIdx.x;
    if (a >= S_keep) return;
    const E* row = mask + (int64_t)srcidx[a] * S;
    E* orow = out + (int64_t)a * S_keep;
    for (int b = threadIdx.x; b < S_keep; b += blockDim.x) orow[b] = row[srcidx[b]];
}

}  // namespace ff
