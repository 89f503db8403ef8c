// Chain links: the by-patch order of the reference (main.py:208-214) as a stable counting sort.
//
// The reference builds a [patch_num, S] boolean matrix (patch_type == arange(P)[:, None]) and takes
// nonzero() of it: 21 MB written and read back at 64x576 tokens plus a host sync.  Here the same
// permutation comes from a three-kernel counting sort over 8 bytes per token:
//   hist     per 512-token chunk, count tokens per patch id          (global atomics on a [C, n_ids] table)
//   colscan  per patch id, exclusive scan of the chunk counts; the last block to finish scans the
//            per-id totals into bucket bases
//   scatter  by-patch position = base[id] + chunk offset + stable rank inside the chunk
// Outputs (by-patch position j, sequence position i):
//   order[j] = i,  chain[j] = patch id,  rank[i] = j or -1 (text / ids outside [0, n_ids))
#pragma once
#include "ff_common.cuh"

namespace ff {

constexpr int LINK_CHUNK = 512;

__global__ void __launch_bounds__(LINK_CHUNK)
k_links_hist(const int64_t* __restrict__ pt, int S, int n_ids, int* __restrict__ hist, int64_t* counters) {
    pdl_trigger();
    __shared__ int s_nv[LINK_CHUNK / 32];
    const int i = blockIdx.x * LINK_CHUNK + threadIdx.x;
    int vis = 0;
    if (i < S) {
        const int64_t id = pt[i];
        vis = (id != -1);
        // bucket n_ids collects every row outside the chains (text, ids >= n_ids): the single-pass kernel walks it
        // like a chain that never merges
        const int b = (id >= 0 && id < n_ids) ? (int)id : n_ids;
        atomicAdd(&hist[(int64_t)blockIdx.x * (n_ids + 1) + b], 1);
    }
    vis = warp_sum_int(vis);
    if ((threadIdx.x & 31) == 0) s_nv[threadIdx.x >> 5] = vis;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < LINK_CHUNK / 32; ++w) t += s_nv[w];
        if (t) atomicAdd((unsigned long long*)&counters[C_NVIS], (unsigned long long)t);
    }
}

// one warp per bucket (patch ids, then the non-chain bucket): exclusive scan of hist[:, id] over chunks;
// totals -> total[id]; the last block scans total[] into base[] and writes N (chain tokens only).
// n_b = n_ids + 1 buckets.
__global__ void __launch_bounds__(256)
k_links_colscan(int* __restrict__ hist, int n_chunks, int n_b, int* __restrict__ total, int* __restrict__ base,
                int64_t* counters, int64_t* status) {
    pdl_enter();
    const int n_ids = n_b;                                 // (name kept below: columns of hist)
    __shared__ int s_scan[33];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31;
    const int id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id < n_ids) {
        int running = 0;
        for (int c0 = 0; c0 < n_chunks; c0 += 32) {
            const int c = c0 + lane;
            const int v = c < n_chunks ? hist[(int64_t)c * n_ids + id] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (c < n_chunks) hist[(int64_t)c * n_ids + id] = running + incl - v;
            running += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) total[id] = running;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = atomicAdd((unsigned long long*)&counters[C_TICKET], 1ull);
        s_last = (t == (unsigned long long)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int carry = 0;
    for (int b0 = 0; b0 < n_ids; b0 += blockDim.x) {
        const int k = b0 + threadIdx.x;
        const int v = k < n_ids ? __ldcg(&total[k]) : 0;
        int tot;
        const int ex = block_exclusive_scan(v, s_scan, &tot);
        if (k < n_ids) base[k] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        counters[C_N] = carry - __ldcg(&total[n_b - 1]);   // the last bucket is not a chain
        counters[C_TICKET] = 0;
        status[FF_ST_NCHAIN] = counters[C_N];              // NVIS is complete: k_links_hist ran before this kernel
        status[FF_ST_NVIS] = counters[C_NVIS];
        status[FF_ST_ERROR] = 0;
    }
}

__global__ void __launch_bounds__(LINK_CHUNK)
k_links_scatter(const int64_t* __restrict__ pt, int S, int n_ids, const int* __restrict__ hist,
                const int* __restrict__ base, int* __restrict__ order, int* __restrict__ chain,
                int* __restrict__ rank) {
    pdl_enter();
    __shared__ __align__(16) int s_key[LINK_CHUNK];
    const int t = threadIdx.x;
    const int i = blockIdx.x * LINK_CHUNK + t;
    int key = -1;
    if (i < S) {
        const int64_t id = pt[i];
        key = (id >= 0 && id < n_ids) ? (int)id : n_ids;
    }
    s_key[t] = key;
    __syncthreads();
    if (i >= S) return;
    // stable rank inside the chunk: earlier tokens of the chunk with the same id (brute force, <= 511 compares)
    int local = 0;
    const int4* k4 = reinterpret_cast<const int4*>(s_key);
    const int full = t >> 2;
    for (int q = 0; q < full; ++q) {
        const int4 v = k4[q];
        local += (v.x == key) + (v.y == key) + (v.z == key) + (v.w == key);
    }
    for (int u = full * 4; u < t; ++u) local += (s_key[u] == key);
    const int pos = base[key] + hist[(int64_t)blockIdx.x * (n_ids + 1) + key] + local;
    order[pos] = i;
    chain[pos] = key;
    rank[i] = key < n_ids ? pos : -1;                      // by-patch position, chain rows only
}

// What the frame-pipelined merge kernel needs of the links, and nothing else: N (chain tokens), n_vis, the first chain
// row, and whether the sequence is a uniform video — ONE span of chain rows whose patch ids run 0, 1, .., n_ids - 1, 0, ..
// (the same counters k_links_hist / k_links_colscan / k_links_seq leave; no order, no ranks).
__global__ void __launch_bounds__(LINK_CHUNK)
k_links_uniform(const int64_t* __restrict__ pt, int S, int n_ids, int64_t* counters, int64_t* status) {
    pdl_trigger();
    __shared__ int s_nv[LINK_CHUNK / 32], s_nc[LINK_CHUNK / 32];
    const int i = blockIdx.x * LINK_CHUNK + threadIdx.x;
    int vis = 0, chain = 0;
    if (i < S) {
        const int64_t id = pt[i], pid = i > 0 ? pt[i - 1] : -1;
        vis = (id != -1);
        chain = (id >= 0 && id < n_ids);
        if (chain) {
            if (!(pid >= 0 && pid < n_ids)) {                   // start of a span of chain rows
                atomicMax((unsigned long long*)&counters[C_FIRSTINV], (unsigned long long)(S - i));
                atomicAdd((unsigned long long*)&counters[C_SPANS], 1ull);
                if (id != 0) counters[C_NONUNI] = 1;
            } else if (id != (pid + 1) % n_ids) {
                counters[C_NONUNI] = 1;
            }
        }
    }
    vis = warp_sum_int(vis);
    chain = warp_sum_int(chain);
    if ((threadIdx.x & 31) == 0) { s_nv[threadIdx.x >> 5] = vis; s_nc[threadIdx.x >> 5] = chain; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tv = 0, tc = 0;
        for (int w = 0; w < LINK_CHUNK / 32; ++w) { tv += s_nv[w]; tc += s_nc[w]; }
        if (tv) atomicAdd((unsigned long long*)&counters[C_NVIS], (unsigned long long)tv);
        if (tc) atomicAdd((unsigned long long*)&counters[C_N], (unsigned long long)tc);
        if (blockIdx.x == 0) status[FF_ST_ERROR] = 0;
    }
}

// Is the sequence a uniform video — ONE span of chain rows whose patch ids run 0, 1, .., n_ids - 1, 0, 1, .. ?  That is the
// layout the frame-pipelined kernel (ff_frame.cuh) serves; counted from the compact by-patch arrays after the counting sort
// (first call of a prefill): C_FIRSTINV = S - first chain row, C_SPANS, C_NONUNI.
__global__ void __launch_bounds__(256)
k_links_seq(const int* __restrict__ rank, const int* __restrict__ order, const int* __restrict__ chain,
            int64_t* __restrict__ counters, int S, int n_ids, unsigned long long* first_inv) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const int j = rank[i];
    if (j >= 0) {
        const int c = chain[j];
        const int jp = i > 0 ? rank[i - 1] : -1;
        if (jp < 0) {                                           // start of a span of chain rows
            atomicMax(first_inv, (unsigned long long)(S - i));
            atomicAdd((unsigned long long*)&counters[C_SPANS], 1ull);
            if (c != 0) counters[C_NONUNI] = 1;
        } else if (c != (chain[jp] + 1) % n_ids) {
            counters[C_NONUNI] = 1;
        }
    }
}

}  // namespace ff
