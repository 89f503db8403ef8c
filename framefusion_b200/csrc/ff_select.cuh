// Selection, run bookkeeping and compaction offsets — the integer part of a merge / prune call.
// These arrays are 1-8 bytes per token (tens of KB): the cost is latency, not bytes.  Short sequences (< 2 048 rows)
// take ONE 1024-thread block with 8 items per thread and no grid-wide synchronisation; long ones a grid of co-resident
// blocks with two (merge) or three to five (prune) grid barriers.
//
//   k_decide_scan / k_keep_scan    main.py:112-127 (branch), 269-301 (runs), 132 (keep mask) + links of the next call
//                                  + the records k_merge_gather is driven by
//   k_prune_scan / k_prune_select  main.py:69-92
#pragma once
#include "ff_common.cuh"

namespace ff {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;

// ---- top-k flags over vals[0..n): selected = the k largest, NaN largest, ties at the k-th value go to
// the lowest indices (the rule oracle/ff_oracle.py:topk_lowest_index states).  One block.  out[j] in {0,1}.
// s_hist: 256 ints, s_scan: 33 ints, s_misc: 4 uint32.
__device__ void block_topk_flags(const float* __restrict__ vals, int n, long long k, uint8_t* __restrict__ out,
                                 int* s_hist, int* s_scan, uint32_t* s_misc) {
    const int t = threadIdx.x;
    if (k <= 0) {
        for (int j = t; j < n; j += blockDim.x) out[j] = 0;
        return;
    }
    if (k >= n) {
        for (int j = t; j < n; j += blockDim.x) out[j] = 1;
        return;
    }
    // radix select, 8 bits per pass from the top: find the key of the k-th largest element
    uint32_t prefix = 0, mask = 0;
    long long need = k;                                    // how many still to take among keys matching prefix
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = t; b < 256; b += blockDim.x) s_hist[b] = 0;
        __syncthreads();
        for (int j = t; j < n; j += blockDim.x) {
            const uint32_t key = float_key(vals[j]);
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255], 1);
        }
        __syncthreads();
        if (t == 0) {
            long long acc = 0;
            int b = 255;
            for (; b > 0; --b) {                           // walk from the largest digit down
                if (acc + s_hist[b] >= need) break;
                acc += s_hist[b];
            }
            s_misc[0] = (uint32_t)b;
            s_misc[1] = (uint32_t)(need - acc);
        }
        __syncthreads();
        prefix |= s_misc[0] << shift;
        mask |= 255u << shift;
        need = (long long)s_misc[1];
        __syncthreads();
    }
    const uint32_t kth = prefix;                           // exact key of the k-th largest
    // ordered pass: everything above kth, and the first `need` elements equal to kth
    int carry = 0;
    for (int base = 0; base < n; base += SEL_TILE) {
        const int j0 = base + t * SEL_ITEMS;
        uint32_t key[SEL_ITEMS];
        int eq = 0;
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            key[e] = (j0 + e < n) ? float_key(vals[j0 + e]) : 0u;
            eq += (j0 + e < n) && (key[e] == kth);
        }
        int tot;
        int ex = carry + block_exclusive_scan(eq, s_scan, &tot);
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            if (j0 + e < n) {
                uint8_t f = key[e] > kth;
                if (key[e] == kth) { f = (ex < need); ++ex; }
                out[j0 + e] = f;
            }
        }
        carry += tot;
    }
    __syncthreads();
}

struct DecideArgs {
    int64_t* counters;
    int64_t* status;            // pinned, device-mapped
    double bound;
    int S;
    const float* sim;
    uint8_t* flag;              // by-patch merge flags (threshold flags on entry)
    const int* order;
    const int* chain;
    const int* rank;
    int* dst;                   // [S] destination row or -1
    int* srcidx;                // [S_keep] source row of every destination row
    int4* rec;                  // [N_next] per kept chain row, by-patch order: (source row, destination row, by-patch position, run length)
    int* order_next;
    int* chain_next;
    int* rank_next;
    int64_t* counters_next;     // counter bank of the next call: N, n_vis carried over, count reset
    int force_branch;           // -1 = decide from count (main.py:116); 0/1 = flags are given (static API)
    long long seq;              // number of this call: status[FF_ST_SEQ] once every other slot is written (ff_status_wait)
};

// the status block is complete: publish the call's number behind a system-wide fence (the host polls this word)
__device__ __forceinline__ void publish_status(int64_t* status, long long seq) {
    __threadfence_system();
    *(volatile int64_t*)&status[FF_ST_SEQ] = seq;
}

// a kept row outside the chains (text): its record goes to the END of rec[], growing downwards, in no particular order
// (the gather treats every record independently); the rows of the chains fill rec[] from the front in by-patch order
__device__ __forceinline__ void text_record(const DecideArgs& a, int src_row, int dst_row) {
    const int slot = (int)atomicAdd((unsigned long long*)&a.counters[C_TICKET2], 1ull);
    a.rec[a.S - 1 - slot] = make_int4(src_row, dst_row, -1, 0);
}

// the whole selection + scan by ONE block of SEL_THREADS threads (short sequences, the top-k branch, the static API)
__device__ __forceinline__ void decide_scan_block(const DecideArgs& a) {
    __shared__ int s_hist[256];
    __shared__ int s_scan[33];
    __shared__ uint32_t s_misc[4];
    __shared__ long long s_k;
    __shared__ int s_branch, s_err;
    const int t = threadIdx.x;
    const int N = (int)a.counters[C_N];
    const int S = a.S;

    if (t == 0) {
        const long long count = a.counters[C_COUNT], n_vis = a.counters[C_NVIS];
        int branch = 0, err = 0;
        long long k = 0;
        if (a.force_branch >= 0) {
            branch = a.force_branch;
        } else if (n_vis == 0) {
            err = 1;                                       // the reference divides by zero here (main.py:114)
        } else {
            const double r = (double)count / (double)n_vis;            // above_k_ratio, Python float division
            if (!(r < a.bound)) {
                branch = 1;
                k = (long long)(a.bound * (double)n_vis);  // int(sparsity_upper_bound * frame_token_num)
                if (k > N) { err = 2; k = N; }             // torch.topk would raise
                if (k < 0) k = 0;
            }
        }
        s_branch = branch; s_k = k; s_err = err;
    }
    __syncthreads();
    const int branch = s_branch;
    if (branch == 1 && a.force_branch < 0)
        block_topk_flags(a.sim, N, s_k, a.flag, s_hist, s_scan, s_misc);
    __syncthreads();

    // ---- sequence-order scan of the keep mask -> destination rows
    int carry = 0;
    for (int base = 0; base < S; base += SEL_TILE) {
        const int i0 = base + t * SEL_ITEMS;
        int r[SEL_ITEMS];
        int keep[SEL_ITEMS];
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) r[e] = (i0 + e < S) ? a.rank[i0 + e] : -1;
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            keep[e] = (i0 + e < S) && !(r[e] >= 0 && a.flag[r[e]]);
            cnt += keep[e];
        }
        int tot;
        int ex = carry + block_exclusive_scan(cnt, s_scan, &tot);
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            if (i0 + e < S) {
                if (keep[e]) {
                    a.dst[i0 + e] = ex;
                    a.srcidx[ex] = i0 + e;
                    if (r[e] < 0) {
                        a.rank_next[ex] = -1;
                        text_record(a, i0 + e, ex);
                    }
                    ++ex;
                } else {
                    a.dst[i0 + e] = -1;
                }
            }
        }
        carry += tot;
    }
    const int s_keep = carry;
    __syncthreads();                                       // dst[] visible to the whole block below

    // ---- by-patch-order compaction: the chain links the next call will use
    carry = 0;
    for (int base = 0; base < N; base += SEL_TILE) {
        const int j0 = base + t * SEL_ITEMS;
        int keep[SEL_ITEMS];
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            keep[e] = (j0 + e < N) && !a.flag[j0 + e];
            cnt += keep[e];
        }
        int tot;
        int ex = carry + block_exclusive_scan(cnt, s_scan, &tot);
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            if (keep[e]) {
                const int d = a.dst[a.order[j0 + e]];
                a.order_next[ex] = d;
                a.chain_next[ex] = a.chain[j0 + e];
                a.rank_next[d] = ex;
                int L = 0;
                while (j0 + e + 1 + L < N && a.flag[j0 + e + 1 + L]) ++L;
                a.rec[ex] = make_int4(a.order[j0 + e], d, j0 + e, L);
                ++ex;
            }
        }
        carry += tot;
    }
    if (t == 0) {
        const int n_next = carry;
        a.counters[C_NNEXT] = n_next;
        a.counters[C_SKEEP] = s_keep;
        a.counters[C_BRANCH] = branch;
        a.counters[C_K] = s_k;
        a.counters[C_NMERGED] = N - n_next;
        a.counters_next[C_N] = n_next;
        a.counters_next[C_NVIS] = a.counters[C_NVIS] - (N - n_next);   // only chain tokens are ever merged away
        a.counters_next[C_COUNT] = 0;
        a.counters_next[C_TICKET] = 0;
        a.counters_next[C_TICKET2] = 0;
        a.status[FF_ST_SEQ_KEEP] = s_keep;
        a.status[FF_ST_COUNT] = a.counters[C_COUNT];
        a.status[FF_ST_NVIS] = a.counters[C_NVIS];
        a.status[FF_ST_NCHAIN] = N;
        a.status[FF_ST_BRANCH] = branch;
        a.status[FF_ST_TOPK] = s_k;
        a.status[FF_ST_ERROR] = s_err;
        a.status[FF_ST_NMERGED] = N - n_next;
        a.status[FF_ST_FUSED] = 0;
        publish_status(a.status, a.seq);
    }
}

__global__ void __launch_bounds__(SEL_THREADS)
k_decide_scan(DecideArgs a) {
    pdl_enter();
    decide_scan_block(a);
}

// ---- threshold branch, many blocks: the same outputs as k_decide_scan (destination rows in sequence order, the
// by-patch arrays of the next call, counters, status) from a grid of G co-resident blocks (G <= SM count) in three
// phases separated by grid barriers.  One 1024-thread block needs ~100 us for 37 k tokens (latency bound); spread
// over 74 blocks it is a few microseconds.  If the count says top-k (main.py:121-127) block 0 runs the single-block
// routine and the others return at once.
struct ScanArgs {
    DecideArgs d;
    int* part;                   // [2 * gridDim.x] kept rows per block (sequence order, by-patch order)
    unsigned* barrier;           // counts up across the calls of one prefill: every block adds 2 per launch
    unsigned bar_base;           // its value when this launch starts (host bookkeeping, ff_api.cu)
};

__device__ __forceinline__ bool threshold_branch(const int64_t* counters, double bound) {
    const long long count = counters[C_COUNT], n_vis = counters[C_NVIS];
    if (n_vis == 0) return false;                          // k_decide_scan reports the division by zero
    return (double)count / (double)n_vis < bound;
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while ((int)(*(volatile unsigned*)bar - target) < 0) __nanosleep(20);
        __threadfence();
    }
    __syncthreads();
}


// sum of part[0 .. n) by the whole block (n <= blockDim.x): one parallel load instead of n dependent ones
__device__ __forceinline__ int block_sum_prefix(const int* part, int n, int* s_scan) {
    const int v = (int)threadIdx.x < n ? __ldcg(&part[threadIdx.x]) : 0;
    int tot;
    block_exclusive_scan(v, s_scan, &tot);
    return tot;
}

__global__ void __launch_bounds__(SEL_THREADS)
k_keep_scan(ScanArgs a) {
    pdl_enter();
    __shared__ int s_scan[33];
    const DecideArgs& d = a.d;
    const int t = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    if (!threshold_branch(d.counters, d.bound)) {
        if (t == 0) atomicAdd(a.barrier, 2u);              // keep the barrier word in step with the host's count
        if (b == 0) decide_scan_block(d);                  // top-k branch (rare: at most once per prefill): one block does it all
        return;
    }
    const int N = (int)d.counters[C_N], S = d.S;
    // block b owns sequence rows [s0, s1) and by-patch positions [n0, n1), both multiples of the scan tile
    const int per_s = ((S + G - 1) / G + SEL_THREADS - 1) / SEL_THREADS * SEL_THREADS;
    const int per_n = ((N + G - 1) / G + SEL_THREADS - 1) / SEL_THREADS * SEL_THREADS;
    const int s0 = min(b * per_s, S), s1 = min(s0 + per_s, S);
    const int n0 = min(b * per_n, N), n1 = min(n0 + per_n, N);

    // ---- phase 1: kept rows of the block, in both orders.  The first tile of each range (usually the only one) stays in
    // registers for the later phases, and what phase 3 needs of it that no phase produces — source row, chain id, run
    // length — is fetched here, ahead of both barriers.
    const int i0 = s0 + t, j0 = n0 + t;
    const int r0 = i0 < s1 ? d.rank[i0] : -1;
    const int keep_s0 = i0 < s1 && !(r0 >= 0 && d.flag[r0]);
    const int keep_n0 = j0 < n1 && !d.flag[j0];
    int oj0 = 0, cj0 = 0, L0 = 0;
    if (keep_n0) {
        oj0 = d.order[j0];
        cj0 = d.chain[j0];
        while (j0 + 1 + L0 < N && d.flag[j0 + 1 + L0]) ++L0;
    }
    int c_seq = keep_s0, c_bp = keep_n0;
    for (int i = i0 + SEL_THREADS; i < s1; i += SEL_THREADS) {
        const int r = d.rank[i];
        c_seq += !(r >= 0 && d.flag[r]);
    }
    for (int j = j0 + SEL_THREADS; j < n1; j += SEL_THREADS) c_bp += !d.flag[j];
    int tot;
    block_exclusive_scan(c_seq, s_scan, &tot);
    if (t == 0) a.part[b] = tot;
    block_exclusive_scan(c_bp, s_scan, &tot);
    if (t == 0) a.part[G + b] = tot;
    grid_barrier(a.barrier, a.bar_base + (unsigned)G);

    // every block count is known now: the three prefix sums of the block, and the totals
    int carry = block_sum_prefix(a.part, b, s_scan);
    int carry_n = block_sum_prefix(a.part + G, b, s_scan);
    const int n_next = block_sum_prefix(a.part + G, G, s_scan);
    const int s_keep = b == G - 1 ? block_sum_prefix(a.part, G, s_scan) : 0;

    // ---- phase 2: destination rows in sequence order
    for (int base = s0; base < s1; base += SEL_THREADS) {
        const int i = base + t;
        const bool first = base == s0;
        const int r = first ? r0 : (i < s1 ? d.rank[i] : -1);
        const int keep = first ? keep_s0 : (i < s1 && !(r >= 0 && d.flag[r]));
        const int ex = carry + block_exclusive_scan(keep, s_scan, &tot);
        if (i < s1) {
            if (keep) {
                d.dst[i] = ex;
                d.srcidx[ex] = i;
                if (r < 0) {
                    d.rank_next[ex] = -1;
                    text_record(d, i, ex);
                }
            } else {
                d.dst[i] = -1;
            }
        }
        carry += tot;
    }
    grid_barrier(a.barrier, a.bar_base + (unsigned)(2 * G));

    // ---- phase 3: the by-patch arrays of the next call and the gather's records.  The records go in last-to-first: the
    // gather then starts with the rows the similarity pass read last (still in the L2).
    for (int base = n0; base < n1; base += SEL_THREADS) {
        const int j = base + t;
        const bool first = base == n0;
        const int keep = first ? keep_n0 : (j < n1 && !d.flag[j]);
        const int ex = carry_n + block_exclusive_scan(keep, s_scan, &tot);
        if (keep) {
            int oj = oj0, cj = cj0, L = L0;
            if (!first) {
                oj = d.order[j];
                cj = d.chain[j];
                L = 0;
                while (j + 1 + L < N && d.flag[j + 1 + L]) ++L;
            }
            const int dd = __ldcg(&d.dst[oj]);
            d.order_next[ex] = dd;
            d.chain_next[ex] = cj;
            d.rank_next[dd] = ex;
            d.rec[n_next - 1 - ex] = make_int4(oj, dd, j, L);
        }
        carry_n += tot;
    }
    if (b == G - 1 && t == 0) {
        d.counters[C_NNEXT] = n_next;
        d.counters[C_SKEEP] = s_keep;
        d.counters[C_BRANCH] = 0;
        d.counters[C_K] = 0;
        d.counters[C_NMERGED] = N - n_next;
        d.counters_next[C_N] = n_next;
        d.counters_next[C_NVIS] = d.counters[C_NVIS] - (N - n_next);
        d.counters_next[C_COUNT] = 0;
        d.counters_next[C_TICKET] = 0;
        d.counters_next[C_TICKET2] = 0;
        d.status[FF_ST_SEQ_KEEP] = s_keep;
        d.status[FF_ST_COUNT] = d.counters[C_COUNT];
        d.status[FF_ST_NVIS] = d.counters[C_NVIS];
        d.status[FF_ST_NCHAIN] = N;
        d.status[FF_ST_BRANCH] = 0;
        d.status[FF_ST_TOPK] = 0;
        d.status[FF_ST_ERROR] = 0;
        d.status[FF_ST_NMERGED] = N - n_next;
        d.status[FF_ST_FUSED] = 0;
        publish_status(d.status, d.seq);
    }
}

// ---- prune stage: importance = T(mean over rows) is computed by k_row_mean; this block selects and scans.
struct PruneArgs {
    int64_t* counters;
    int64_t* status;
    const float* imp;           // [S] float32 holding T values
    uint8_t* sel;               // [length] scratch: 1 = kept by top-k
    int* dst;
    int* srcidx;
    int S, start, length;
    long long k;
    long long seq;              // see DecideArgs
};

__global__ void __launch_bounds__(SEL_THREADS)
k_prune_scan(PruneArgs a) {
    pdl_enter();
    __shared__ int s_hist[256];
    __shared__ int s_scan[33];
    __shared__ uint32_t s_misc[4];
    const int t = threadIdx.x;
    block_topk_flags(a.imp + a.start, a.length, a.k, a.sel, s_hist, s_scan, s_misc);
    __syncthreads();
    int carry = 0;
    for (int base = 0; base < a.S; base += SEL_TILE) {
        const int i0 = base + t * SEL_ITEMS;
        int keep[SEL_ITEMS];
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            const int i = i0 + e;
            int kp = 0;
            if (i < a.S) kp = (i < a.start || i >= a.start + a.length) ? 1 : a.sel[i - a.start];
            keep[e] = kp;
            cnt += kp;
        }
        int tot;
        int ex = carry + block_exclusive_scan(cnt, s_scan, &tot);
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            const int i = i0 + e;
            if (i < a.S) {
                if (keep[e]) { a.dst[i] = ex; a.srcidx[ex] = i; ++ex; } else a.dst[i] = -1;
            }
        }
        carry += tot;
    }
    if (t == 0) {
        a.counters[C_SKEEP] = carry;
        a.status[FF_ST_SEQ_KEEP] = carry;
        a.status[FF_ST_TOPK] = a.k;
        a.status[FF_ST_ERROR] = 0;
        publish_status(a.status, a.seq);
    }
}

// ---- prune stage on a grid of co-resident blocks: radix select of the k-th largest importance (4 passes of 8 bits,
// block histograms merged with global atomics, a grid barrier per pass, one more for the tie / greater prefixes), ties at the k-th value to the lowest
// indices, then the sequence-order scan of the keep flags.  Same outputs as k_prune_scan, a few microseconds
// instead of ~75 us for 22 k tokens in one block.
struct PruneGridArgs {
    PruneArgs p;
    int* hist;                  // [4][256], zeroed by the host
    int* part;                  // [2 * gridDim.x]
    unsigned* barrier;          // zeroed by the host
    int n_passes;               // 8-bit radix passes that can differ: 2 for bf16 values, 3 for f16, 4 for f32
};

__global__ void __launch_bounds__(SEL_THREADS)
k_prune_select(PruneGridArgs a) {
    pdl_enter();
    __shared__ int s_hist[256];
    __shared__ int s_scan[33];
    __shared__ uint32_t s_digit, s_need;
    const PruneArgs& p = a.p;
    const int t = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    const int n = p.length, S = p.S;
    const float* vals = p.imp + p.start;
    const int per_n = ((n + G - 1) / G + SEL_THREADS - 1) / SEL_THREADS * SEL_THREADS;
    const int n0 = min(b * per_n, n), n1 = min(n0 + per_n, n);
    unsigned bar_target = 0;

    uint32_t kth = 0, need_eq = 0;
    bool all = p.k >= n, none = p.k <= 0;
    if (!all && !none) {
        uint32_t prefix = 0, mask = 0;
        long long need = p.k;
        for (int pass = 0; pass < a.n_passes; ++pass) {        // the remaining low bits are zero in every key
            const int shift = 24 - 8 * pass;
            for (int q = t; q < 256; q += SEL_THREADS) s_hist[q] = 0;
            __syncthreads();
            for (int j = n0 + t; j < n1; j += SEL_THREADS) {
                const uint32_t key = float_key(vals[j]);
                if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255], 1);
            }
            __syncthreads();
            for (int q = t; q < 256; q += SEL_THREADS)
                if (s_hist[q]) atomicAdd(&a.hist[pass * 256 + q], s_hist[q]);
            bar_target += G;
            grid_barrier(a.barrier, bar_target);
            for (int q = t; q < 256; q += SEL_THREADS) s_hist[q] = __ldcg(&a.hist[pass * 256 + q]);
            __syncthreads();
            if (t < 32) {
                // the digit of the k-th value: walk from the largest digit down until `need` elements are covered.  One warp:
                // lane l owns digits [8l, 8l + 8); a suffix sum over the lanes finds the lane the walk stops in.
                int loc[8];
                long long mine = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) { loc[e] = s_hist[8 * t + e]; mine += loc[e]; }
                long long above = mine;                    // inclusive suffix sum: this lane's digits and all larger ones
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long v = __shfl_down_sync(FULL, above, o);
                    if (t + o < 32) above += v;
                }
                above -= mine;                             // elements in digits above this lane's range
                if (t == 0) { s_digit = 0; s_need = (uint32_t)(need - above - (mine - loc[0])); }   // the walk never stops before digit 0
                __syncwarp();
                if (above < need && above + mine >= need && !(t == 0 && above + mine - loc[0] < need)) {
                    long long acc = above;
                    int e = 7;
                    for (; e > 0; --e) {
                        if (acc + loc[e] >= need) break;
                        acc += loc[e];
                    }
                    s_digit = (uint32_t)(8 * t + e);
                    s_need = (uint32_t)(need - acc);
                }
            }
            __syncthreads();
            prefix |= s_digit << shift;
            mask |= 255u << shift;
            need = (long long)s_need;
            __syncthreads();
        }
        kth = prefix;
        need_eq = (uint32_t)need;
    }

    // ---- how many elements equal to / greater than the k-th value sit before this block.  With those two prefixes a
    // row's destination follows directly: kept vision rows before it = (greater before) + min(equal before, need_eq)
    // — ties go to the lowest indices — so neither the selection flags nor a second scan have to cross blocks.
    // keys are compared under the mask of the digits the passes resolved: the low bits of a bf16 / f16 value are zero,
    // but float_key() turns them into ones for negative values and NaN
    const uint32_t kmask = a.n_passes >= 4 ? 0xffffffffu : ~(0xffffffffu >> (8 * a.n_passes));
    int c_eq = 0, c_gt = 0;
    if (!all && !none)
        for (int j = n0 + t; j < n1; j += SEL_THREADS) {
            const uint32_t key = float_key(vals[j]) & kmask;
            c_eq += key == kth;
            c_gt += key > kth;
        }
    int tot;
    block_exclusive_scan(c_eq, s_scan, &tot);
    if (t == 0) a.part[b] = tot;
    block_exclusive_scan(c_gt, s_scan, &tot);
    if (t == 0) a.part[G + b] = tot;
    bar_target += G;
    grid_barrier(a.barrier, bar_target);
    int carry = block_sum_prefix(a.part, b, s_scan);
    int carry_gt = block_sum_prefix(a.part + G, b, s_scan);
    const int k_sel = all ? n : (none ? 0 : (int)p.k);      // vision rows that stay
    for (int base = n0; base < n1; base += SEL_THREADS) {
        const int j = base + t;
        uint32_t key = 0;
        int eq = 0, gt = 0;
        if (j < n1 && !all && !none) { key = float_key(vals[j]) & kmask; eq = key == kth; gt = key > kth; }
        const int ex_eq = carry + block_exclusive_scan(eq, s_scan, &tot);
        carry += tot;
        const int ex_gt = carry_gt + block_exclusive_scan(gt, s_scan, &tot);
        carry_gt += tot;
        if (j < n1) {
            const int keep = all ? 1 : (none ? 0 : (int)(gt || (eq && (uint32_t)ex_eq < need_eq)));
            p.sel[j] = (uint8_t)keep;
            const int i = p.start + j;
            if (keep) {
                const int d = p.start + (all ? j : ex_gt + (int)min((uint32_t)ex_eq, need_eq));
                p.dst[i] = d;
                p.srcidx[d] = i;
            } else {
                p.dst[i] = -1;
            }
        }
    }
    // the rows around the vision span always stay: before it in place, behind it shifted by the rows that went
    for (int i = b * SEL_THREADS + t; i < S; i += G * SEL_THREADS) {
        if (i < p.start) {
            p.dst[i] = i;
            p.srcidx[i] = i;
        } else if (i >= p.start + n) {
            const int d = i - (n - k_sel);
            p.dst[i] = d;
            p.srcidx[d] = i;
        }
    }
    const int s_keep = S - (n - k_sel);
    if (b == G - 1 && t == 0) {
        p.counters[C_SKEEP] = s_keep;
        p.status[FF_ST_SEQ_KEEP] = s_keep;
        p.status[FF_ST_TOPK] = p.k;
        p.status[FF_ST_ERROR] = 0;
        publish_status(p.status, p.seq);
    }
}

// importance[s] = T( sum_rows attn[row][s] / n_rows )      (torch.mean(dim=(1,2)), main.py:70)
template <int DT>
__global__ void k_row_mean(const void* __restrict__ attn, int n_rows, int S, float* __restrict__ imp) {
    pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    float acc = 0.f;
    for (int r = 0; r < n_rows; ++r) acc += Num<DT>::load(attn, (int64_t)r * S + s);
    imp[s] = Num<DT>::rnd(acc / (float)n_rows);
}

}  // namespace ff
