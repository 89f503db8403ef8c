// C ABI of libframefusion_b200.so — see include/framefusion_b200.h for the contract of every entry point.
// Host side only: argument checks, workspace carving, kernel launches on the caller's stream.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ff_common.cuh"
#include "ff_frame.cuh"
#include "ff_importance.cuh"
#include "ff_links.cuh"
#include "ff_merge.cuh"
#include "ff_select.cuh"
#include "ff_similarity.cuh"

using namespace ff;

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};      // kernels this library has launched (ff_launch_count)

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define FF_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(FF_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));    \
    } while (0)

#define FF_LAUNCH_CHECK(name)                                                                      \
    do {                                                                                           \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess) return fail(FF_E_CUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

// Launch with programmatic stream serialisation: the grid may become resident while the previous kernel of the
// stream drains (it has to pass pdl_wait() before touching that kernel's results), which hides the launch latency
// between the short dependent kernels of one call.  Every kernel launched through here starts with pdl_wait().
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define FF_LAUNCH(name, kernel, grid, block, smem, st, ...)                                        \
    do {                                                                                           \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
        cudaError_t e_ = launch_pdl(kernel, dim3(grid), dim3(block), smem, st, __VA_ARGS__);       \
        if (e_ != cudaSuccess) return fail(FF_E_CUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

struct ff_ctx {
    int device;
    int64_t* h_status;
    int64_t* d_status;
    int parity;          // which links / counter bank describes the current sequence
    int last_parity;     // bank the last merge call read (ff_debug_read)
    int64_t links_S;     // sequence length the links describe (-1: none, -2: S_keep of the last merge call)
    int64_t cap;         // capacity the workspace is carved for (set by ff_build_links)
    int64_t n_ids;
    int have_order;      // compact by-patch order / chain / rank arrays valid for `parity` (multi-kernel path)
    int fresh_links;     // ... and they are the ones ff_build_links made (counters[C_FIRSTINV / C_SPANS / C_NONUNI] valid)
    int words_clean[2];  // the flag words of the bank (frame-pipelined kernel) are known to be zero
    int sm_count;
    int max_smem;        // opt-in dynamic shared memory per block
    int smem_per_sm, smem_reserved;   // shared memory of an SM / what the system keeps per resident block
    cudaEvent_t ev_start, ev_stop;   // ff_ctx_timing
    int count_clean[2];  // counters[bank][C_COUNT] is known to be zero (set by the kernel that decided the previous call)
    unsigned bar_base;   // value of the grid-barrier word before the next k_keep_scan (it is not reset between merge calls)
    int bar_dirty;       // the prune stage left the barrier word at an unknown value
    long long seq;       // number of the last reducing call (status[FF_ST_SEQ] when its results are in the status block)
    int links_lite;      // ff_build_links_for left only the counters: the next merge call must be the frame-pipelined kernel
    int links_lite_last; // ... and the last merge call ran on such links (no by-patch order to read back)
    int frame_smem[9];   // dynamic shared memory the frame-pipelined kernel of each (dtype, build) is opted in for
    long long* frame_trace;          // ff_debug_frame_trace: device buffer for time stamps of the next frame-kernel launch
    int64_t frame_trace_bytes;
};

// every entry point runs on the context's device and leaves the caller's current device as it found it (PyTorch
// tracks the current device through the runtime: a layer-split model calls in from several devices)
struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int device) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define FF_DEVICE(ctx)                                                                             \
    DeviceGuard guard_((ctx)->device);                                                             \
    if (!guard_.ok) return fail(FF_E_CUDA, "cudaSetDevice(%d) failed", (ctx)->device)

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
namespace {

struct Ws {
    int64_t* counters[2];
    int* order[2];
    int* chain[2];
    int* rank[2];
    int* len[2];        // [n_ids + 1] rows per chain; bucket n_ids = rows outside the chains
    float* sim;
    uint8_t* flag;
    unsigned long long* desc[2];     // frame-pipelined kernel: barrier / abort words, then (reported << 32 | kept) per 32 rows
    int* dst[2];
    int* srcidx;
    int4* rec;          // [cap] per kept chain row in by-patch order: (source row, destination row, by-patch position, run length)
    int* hist;
    int* part;          // [2 * 160] per-block counts of the multi-block scan
    unsigned* barrier;  // its grid barrier
    int* sel_hist;      // [4][256] digit histograms of the multi-block radix select (right after `barrier`)
    int* base;          // [n_ids + 1] start of every chain list inside order[]
    char* zero_begin;   // region ff_build_links clears
    size_t zero_bytes;
    size_t bytes;
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

Ws carve(void* base_ptr, int64_t cap, int64_t n_ids) {
    Ws w;
    char* p = (char*)base_ptr;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    const int64_t n_chunks = (cap + LINK_CHUNK - 1) / LINK_CHUNK;
    // one region zeroed by ff_build_links with a single memset: counters, chain lengths, histogram, flag bytes
    char* zero_begin = p ? p + off : nullptr;
    w.counters[0] = (int64_t*)take(C_SLOTS * 8);
    w.counters[1] = (int64_t*)take(C_SLOTS * 8);
    w.len[0] = (int*)take((size_t)(n_ids + 1) * 4);
    w.hist = (int*)take((size_t)n_chunks * (size_t)(n_ids + 1) * 4);
    w.barrier = (unsigned*)take(256);
    w.sel_hist = (int*)take(4 * 256 * 4);
    for (int b = 0; b < 2; ++b) {
        w.desc[b] = (unsigned long long*)take(16 + ((size_t)cap / 32 + 2) * 8);   // two words in front, u64 per 32 rows
    }
    w.zero_begin = zero_begin;
    w.zero_bytes = (size_t)((p ? p + off : (char*)nullptr) - zero_begin);
    w.len[1] = (int*)take((size_t)(n_ids + 1) * 4);
    for (int b = 0; b < 2; ++b) {
        w.order[b] = (int*)take(cap * 4);
        w.chain[b] = (int*)take(cap * 4);
        w.rank[b] = (int*)take(cap * 4);
    }
    w.sim = (float*)take(cap * 4);
    w.flag = (uint8_t*)take(cap);
    w.dst[0] = (int*)take(cap * 4);
    w.dst[1] = (int*)take(cap * 4);
    w.srcidx = (int*)take(cap * 4);
    w.rec = (int4*)take(cap * 16);
    w.base = (int*)take((size_t)(n_ids + 1) * 4);
    w.part = (int*)take(2 * 160 * 4);
    w.bytes = off;
    return w;
}

int check_ws(const ff_ctx* ctx, const void* ws, int64_t ws_bytes, int64_t S, Ws* out) {
    if (!ctx || !ws) return fail(FF_E_BADARG, "null ctx / workspace");
    if (((uintptr_t)ws & 255) != 0) return fail(FF_E_BADARG, "workspace must be 256-byte aligned");
    if (S > ctx->cap) return fail(FF_E_WORKSPACE, "sequence length %lld exceeds the capacity %lld the workspace was laid out for", (long long)S, (long long)ctx->cap);
    *out = carve((void*)ws, ctx->cap, ctx->n_ids);
    if ((int64_t)out->bytes > ws_bytes)
        return fail(FF_E_WORKSPACE, "workspace too small: need %zu bytes, have %lld", out->bytes, (long long)ws_bytes);
    return FF_OK;
}

inline bool vec_ok(const void* a, const void* b, int64_t H, int dtype) {
    const int64_t eb = dtype == FF_F32 ? 4 : 2;
    return ((H * eb) % 16 == 0) && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

__global__ void k_store_sim(const float* __restrict__ sim, const int64_t* __restrict__ counters, int dtype, void* out,
                            const int* __restrict__ order, int64_t* order_out) {
    const int N = (int)counters[C_N];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        if (out) {
            if (dtype == FF_BF16) Num<FF_BF16>::store(out, j, sim[j]);
            else if (dtype == FF_F16) Num<FF_F16>::store(out, j, sim[j]);
            else ((float*)out)[j] = sim[j];
        }
        if (order_out) order_out[j] = order[j];
    }
}

__global__ void k_store_vals(const float* __restrict__ v, int n, int dtype, void* out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (dtype == FF_BF16) Num<FF_BF16>::store(out, j, v[j]);
    else if (dtype == FF_F16) Num<FF_F16>::store(out, j, v[j]);
    else ((float*)out)[j] = v[j];
}

__global__ void k_store_order(const int* __restrict__ order, int n, int64_t* out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = order[j];
}

__global__ void k_keep_from_dst(const int* __restrict__ dst, int n, uint8_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = dst[i] >= 0;
}

__global__ void k_srcidx_from_dst(const int* __restrict__ dst, int n, int* __restrict__ srcidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && dst[i] >= 0) srcidx[dst[i]] = i;
}

__global__ void k_copy_u8(const uint8_t* __restrict__ src, int n, uint8_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[i];
}

__global__ void k_links_status(const int64_t* counters, int64_t* status) {
    status[FF_ST_NCHAIN] = counters[C_N];
    status[FF_ST_NVIS] = counters[C_NVIS];
    status[FF_ST_ERROR] = 0;
}

__global__ void k_count_status(const int64_t* counters, int64_t* status) {
    status[FF_ST_NCHAIN] = counters[C_N];
    status[FF_ST_NVIS] = counters[C_NVIS];
    status[FF_ST_COUNT] = counters[C_COUNT];
    status[FF_ST_ERROR] = 0;
}

template <typename F>
int dispatch_dtype(int dtype, F&& f) {
    switch (dtype) {
        case FF_BF16: return f(std::integral_constant<int, FF_BF16>());
        case FF_F16: return f(std::integral_constant<int, FF_F16>());
        case FF_F32: return f(std::integral_constant<int, FF_F32>());
    }
    return fail(FF_E_BADARG, "unsupported dtype %d", dtype);
}

int pack_aux(const ff_aux* aux, int n_aux, AuxPack* p) {
    if (n_aux < 0 || n_aux > FF_MAX_AUX) return fail(FF_E_BADARG, "n_aux %d outside [0, %d]", n_aux, FF_MAX_AUX);
    if (n_aux && !aux) return fail(FF_E_BADARG, "aux is null");
    p->n = n_aux;
    for (int i = 0; i < n_aux; ++i) {
        if (!aux[i].src || !aux[i].dst || aux[i].planes < 1 || aux[i].row_bytes < 1)
            return fail(FF_E_BADARG, "aux[%d] malformed", i);
        p->a[i] = aux[i];
    }
    return FF_OK;
}

int launch_similarity(const Ws& w, int bank, const void* hidden, int dtype, int64_t S, int64_t H, double thr,
                      cudaStream_t st) {
    const bool vec = vec_ok(hidden, hidden, H, dtype);
    const int grid = (int)((S + 7) / 8);
    if (grid == 0) return FF_OK;
    return dispatch_dtype(dtype, [&](auto dt) {
        constexpr int DT = decltype(dt)::value;
        if (vec)
            FF_LAUNCH("k_similarity", (k_similarity<DT, true>), grid, 256, 0, st, hidden, (int)H, (int)S, w.order[bank], w.chain[bank],
                      w.counters[bank], (float)thr, w.sim, w.flag, w.counters[bank]);
        else
            FF_LAUNCH("k_similarity", (k_similarity<DT, false>), grid, 256, 0, st, hidden, (int)H, (int)S, w.order[bank], w.chain[bank],
                      w.counters[bank], (float)thr, w.sim, w.flag, w.counters[bank]);
        return (int)FF_OK;
    });
}

int launch_merge_compact(const Ws& w, int bank, const void* hidden, void* out, int dtype, int64_t S, int64_t H,
                         int use_flags, cudaStream_t st) {
    const bool vec = vec_ok(hidden, out, H, dtype);
    const int grid = (int)((S + 7) / 8);
    if (grid == 0) return FF_OK;
    return dispatch_dtype(dtype, [&](auto dt) {
        constexpr int DT = decltype(dt)::value;
        if (vec)
            k_merge_compact<DT, true><<<grid, 256, 0, st>>>(hidden, out, (int)S, (int)H, w.rank[bank], w.order[bank], w.flag,
                                                            w.dst[bank], w.counters[bank], use_flags);
        else
            k_merge_compact<DT, false><<<grid, 256, 0, st>>>(hidden, out, (int)S, (int)H, w.rank[bank], w.order[bank], w.flag,
                                                             w.dst[bank], w.counters[bank], use_flags);
        FF_LAUNCH_CHECK("k_merge_compact");
        return (int)FF_OK;
    });
}



// blocks of k_merge_gather: one warp per row, plus one block of aux rows behind every four of those
static int gather_grid(int64_t S, int n_aux) {
    const int64_t main_blocks = (S + GATHER_WARPS - 1) / GATHER_WARPS;
    return (int)(n_aux ? 5 * ((main_blocks + 3) / 4) : main_blocks);
}

// merged rows + compaction of hidden and aux, two-pass path: the gather kernel when rows are 16-byte multiples,
// the generic kernels otherwise
int launch_merge_gather(const ff_ctx* ctx, const Ws& w, int bank, const void* hidden, void* out, int dtype, int64_t S,
                        int64_t H, const AuxPack& ap, cudaStream_t st) {
    if (S == 0) return FF_OK;
    const int64_t nvec = H * (dtype == FF_F32 ? 4 : 2) / 16;
    if (vec_ok(hidden, out, H, dtype)) {
        return dispatch_dtype(dtype, [&](auto dt) {
            constexpr int DT = decltype(dt)::value;
            FF_LAUNCH("k_merge_gather", k_merge_gather<DT>, gather_grid(S, ap.n), GATHER_WARPS * 32, 0, st,
                      hidden, out, (int)nvec, (int)S, w.srcidx, w.rec, w.order[bank], w.flag, w.counters[bank], w.counters[bank ^ 1], ap);
            return (int)FF_OK;
        });
    }
    if (int rc = launch_merge_compact(w, bank, hidden, out, dtype, S, H, 1, st)) return rc;
    if (ap.n) {
        k_aux_compact<<<(int)((S + 7) / 8), 256, 0, st>>>(ap, (int)S, w.dst[bank]);
        FF_LAUNCH_CHECK("k_aux_compact");
    }
    return FF_OK;
}

// the sequence length the links describe becomes known to the host only after it has synchronised and read
// S_keep from the status block; by then h_status holds it.
int links_ready(ff_ctx* ctx, int64_t S, bool need_order) {
    if (ctx->links_S == -2) ctx->links_S = ctx->h_status[FF_ST_SEQ_KEEP];
    if (ctx->links_S != S)
        return fail(FF_E_BADARG, "chain links in the workspace describe S=%lld, not %lld: call ff_build_links", (long long)ctx->links_S, (long long)S);
    if (need_order && !ctx->have_order)
        return fail(FF_E_BADARG, "by-patch order not in the workspace: call ff_build_links");
    return FF_OK;
}

int check_shape(int64_t S, int64_t H, int dtype) {
    if (dtype < 0 || dtype > 2) return fail(FF_E_BADARG, "unsupported dtype %d", dtype);
    if (S < 0 || H < 1) return fail(FF_E_BADARG, "bad shape S=%lld H=%lld", (long long)S, (long long)H);
    if (S > (1ll << 30) || H > (1ll << 24)) return fail(FF_E_UNSUPPORTED, "S=%lld H=%lld too large", (long long)S, (long long)H);
    return FF_OK;
}

// The frame-pipelined kernel (ff_frame.cuh) serves the first merge call of a prefill when the rows are 16-byte
// multiples, no chain head can pass the threshold, and a ring of at least four frames of the CTA's chains fits shared
// memory.  Whether the layout is the uniform video it is built for is checked on the device (status ERROR = 3 if not).
struct FramePlan {
    int R, n_stages, smem, grid, threads;
};

// the part of the decision that depends on the shape alone (ff_build_links_for asks before the tensors are known)
bool frame_shape(const ff_ctx* ctx, int64_t n_ids, int64_t S, int64_t row_bytes, bool force, FramePlan* fp) {
    static const int off = getenv("FF_NO_FRAME") ? atoi(getenv("FF_NO_FRAME")) : 0;
    if (off || n_ids < 1 || S < 1 || S >= (1ll << 30)) return false;
    if (row_bytes < 16 || row_bytes % 16 != 0) return false;
    const int64_t P = n_ids;
    const int64_t R = (P + ctx->sm_count - 1) / ctx->sm_count;
    if (R > FR_MAXR) return false;
    const int64_t stage = R * row_bytes;
    // (FR_ACC_INPLACE: run sums live in the stage slots themselves, else one stage-sized block of accumulator rows)
    int64_t n_stages = ((int64_t)ctx->max_smem - FR_META - (FR_ACC_INPLACE ? 0 : stage)) / stage;
    static const int max_stages = getenv("FF_FRAME_STAGES") ? atoi(getenv("FF_FRAME_STAGES")) : FR_MAXSTAGES;
    if (n_stages > max_stages) n_stages = max_stages;
    if (n_stages > FR_MAXSTAGES) n_stages = FR_MAXSTAGES;
    if (n_stages < 4) return false;
    const int64_t grid = (P + R - 1) / R;
    // Where it pays (profiles/r02_sweep.jsonl): the pipeline advances one FRAME per ~1.45 us whatever a frame of a CTA
    // holds, so a CTA has to move enough bytes per frame (>= 20 KB: 4 rows of 7 KB; 2 rows — 210 tokens per frame — lose
    // to the multi-kernel path), nearly every SM has to have chains, and the ring has to cover the ~9 us a stage lives
    // (>= 5 frames: five chains of 8-KB rows — C4 — leave five since the run sums moved into the stage slots, and win by
    // 5 %; with the four they had before they lost by 50 %).  `force` (flags bit 2) takes the kernel wherever it CAN run.
    if (!force && (stage < 20 * 1024 || n_stages < 5 || grid * 10 < (int64_t)ctx->sm_count * 9)) return false;
    fp->R = (int)R;
    fp->n_stages = (int)n_stages;
    fp->smem = (int)(FR_META + (n_stages + (FR_ACC_INPLACE ? 0 : 1)) * stage);
    fp->grid = (int)grid;
    fp->threads = 32 * (2 + (R <= 4 ? FR_NPW : 1) + 3 * (int)R);   // producer, prefix warps, (finisher slot), then S / G / aux per chain
    return true;
}

bool frame_plan(const ff_ctx* ctx, const void* hidden, const void* out, int dtype, int64_t S, int64_t H, double thr, bool force, FramePlan* fp) {
    if (!ctx->fresh_links || (((uintptr_t)hidden | (uintptr_t)out) & 15) != 0 || !(thr > -2.0)) return false;
    return frame_shape(ctx, ctx->n_ids, S, H * (dtype == FF_F32 ? 4 : 2), force, fp);
}

int launch_frame(ff_ctx* ctx, const Ws& w, int bank, const FramePlan& fp, const void* hidden, void* out, int dtype, int64_t S,
                 int64_t H, double thr, double bound, const AuxPack& ap, cudaStream_t st) {
    const int nb = bank ^ 1;
    FrameArgs a;
    a.hidden = (const char*)hidden;
    a.out = (char*)out;
    a.S = (int)S;
    a.P = (int)ctx->n_ids;
    a.R = fp.R;
    a.row_bytes = (int)(H * (dtype == FF_F32 ? 4 : 2));
    a.nvec = a.row_bytes / 16;
    a.n_stages = fp.n_stages;
    a.gbar = reinterpret_cast<unsigned*>(w.desc[bank]);
    a.words = w.desc[bank] + 1;
    a.sim = w.sim;
    a.flag = w.flag;
    a.dst = w.dst[bank];
    a.keptdst = reinterpret_cast<int*>(w.rec);
    a.len_next = w.len[1];
    a.order_next = w.order[nb];
    a.chain_next = w.chain[nb];
    a.rank_next = w.rank[nb];
    a.counters = w.counters[bank];
    a.counters_next = w.counters[nb];
    a.status = ctx->d_status;
    a.thr = (float)thr;
    a.bound = bound;
    a.seq = ++ctx->seq;
    a.trace = nullptr;
    a.trace_frames = 0;
    if (ctx->frame_trace) {
        const int64_t per_frame = (int64_t)fp.grid * FR_TRACE_K * 8;
        a.trace_frames = (int)(ctx->frame_trace_bytes / per_frame < 4096 ? ctx->frame_trace_bytes / per_frame : 4096);
        a.trace = a.trace_frames > 0 ? ctx->frame_trace : nullptr;
        ctx->frame_trace = nullptr;                        // one launch
    }
    a.auxf.n = 0;
    for (int q = 0; q < ap.n && a.auxf.n >= 0; ++q) {
        const ff_aux& x = ap.a[q];
        const uintptr_t al = (uintptr_t)x.src | (uintptr_t)x.dst | (uintptr_t)x.src_plane_stride | (uintptr_t)x.dst_plane_stride | (uintptr_t)x.row_bytes;
        const int piece = (al & 15) == 0 ? 16 : ((al & 7) == 0 ? 8 : 0);
        if (piece == 0 || x.row_bytes > 32 * piece || a.auxf.n + x.planes > 8) { a.auxf.n = -1; break; }
        for (int64_t pl = 0; pl < x.planes; ++pl) {
            const int e = a.auxf.n++;
            a.auxf.row_bytes[e] = (int)x.row_bytes;
            a.auxf.piece[e] = piece;
            a.auxf.src[e] = (const char*)x.src + pl * x.src_plane_stride;
            a.auxf.dst[e] = (char*)x.dst + pl * x.dst_plane_stride;
        }
    }
    // (ff_ctx_timing: the start event goes in when the arguments are ready — right in front of the call's first stream operation)
    if (ctx->ev_start) FF_CUDA(cudaEventRecord(ctx->ev_start, st));
    if (!ctx->words_clean[bank]) FF_CUDA(cudaMemsetAsync(w.desc[bank], 0, (size_t)(S / 32 + 2) * 8, st));
    ctx->words_clean[bank] = 0;
    return dispatch_dtype(dtype, [&](auto dt) {
        constexpr int DT = decltype(dt)::value;
        // three builds: up to four chains per CTA (16 warps), exactly five (C4: 18 warps — 112 registers, what the roles need
        // without spilling; the build for up to FR_MAXR chains has 27 warps, 72 registers and spills) and up to FR_MAXR
        auto go = [&](auto mr) {
            constexpr int MR = decltype(mr)::value;
            constexpr int NPW = MR > 4 ? 1 : FR_NPW;       // prefix warps (FramePlan::threads); two changed nothing in the five-chain build
            constexpr int slot = DT * 3 + (MR > 5 ? 2 : MR > 4 ? 1 : 0);
            if (ctx->frame_smem[slot] < fp.smem) {
                FF_CUDA(cudaFuncSetAttribute((k_frame_merge<DT, MR, NPW>), cudaFuncAttributeMaxDynamicSharedMemorySize, fp.smem));
                ctx->frame_smem[slot] = fp.smem;
            }
            FF_LAUNCH("k_frame_merge", (k_frame_merge<DT, MR, NPW>), fp.grid, fp.threads, fp.smem, st, a, ap);
            return (int)FF_OK;
        };
        if (fp.R <= 4) return go(std::integral_constant<int, 4>());
        if (fp.R == 5) return go(std::integral_constant<int, 5>());
        return go(std::integral_constant<int, FR_MAXR>());
    });
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int ff_abi_version(void) { return FF_ABI_VERSION; }

const char* ff_last_error(void) { return g_err; }

int64_t ff_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int ff_ctx_create(int device, ff_ctx** out) {
    if (!out) return fail(FF_E_BADARG, "out is null");
    *out = nullptr;
    DeviceGuard guard_(device);
    if (!guard_.ok) return fail(FF_E_CUDA, "cudaSetDevice(%d) failed", device);
    ff_ctx* c = new (std::nothrow) ff_ctx();
    if (!c) return fail(FF_E_BADARG, "out of host memory");
    c->device = device;
    c->parity = 0;
    c->last_parity = 0;
    c->links_S = -1;
    c->cap = 0;
    c->n_ids = 0;
    c->have_order = 0;
    c->fresh_links = 0;
    c->seq = 0;
    c->words_clean[0] = c->words_clean[1] = 0;
    for (int i = 0; i < 9; ++i) c->frame_smem[i] = 0;
    c->links_lite = c->links_lite_last = 0;
    c->frame_trace = nullptr;
    c->frame_trace_bytes = 0;
    cudaError_t e = cudaHostAlloc((void**)&c->h_status, FF_ST_SLOTS * 8, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) { delete c; return fail(FF_E_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e)); }
    memset(c->h_status, 0, FF_ST_SLOTS * 8);
    e = cudaHostGetDevicePointer((void**)&c->d_status, c->h_status, 0);
    if (e != cudaSuccess) { cudaFreeHost(c->h_status); delete c; return fail(FF_E_CUDA, "cudaHostGetDevicePointer: %s", cudaGetErrorString(e)); }
    e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) c->sm_count = 148;
    e = cudaDeviceGetAttribute(&c->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) c->max_smem = 48 * 1024;
    if (cudaDeviceGetAttribute(&c->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device) != cudaSuccess) c->smem_per_sm = c->max_smem;
    if (cudaDeviceGetAttribute(&c->smem_reserved, cudaDevAttrReservedSharedMemoryPerBlock, device) != cudaSuccess) c->smem_reserved = 1024;
    *out = c;
    return FF_OK;
}

int ff_ctx_destroy(ff_ctx* ctx) {
    if (!ctx) return FF_OK;
    cudaFreeHost(ctx->h_status);
    delete ctx;
    return FF_OK;
}

const int64_t* ff_ctx_status(const ff_ctx* ctx) { return ctx ? ctx->h_status : nullptr; }

int ff_ctx_timing(ff_ctx* ctx, void* ev_start, void* ev_stop) {
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    if ((ev_start == nullptr) != (ev_stop == nullptr)) return fail(FF_E_BADARG, "pass two events or two NULLs");
    ctx->ev_start = (cudaEvent_t)ev_start;
    ctx->ev_stop = (cudaEvent_t)ev_stop;
    return FF_OK;
}

int ff_stream_sync(ff_ctx* ctx, void* stream) {
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    FF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return FF_OK;
}

int ff_status_wait(ff_ctx* ctx, void* stream) {
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    cudaStream_t st = (cudaStream_t)stream;
    const volatile int64_t* seq = &ctx->h_status[FF_ST_SEQ];
    // the deciding kernel of a call runs a few tens of microseconds after its launch: poll the mapped word; look at the
    // stream now and then (a failed launch never writes it), and give up polling after a generous while
    for (long spins = 0; *seq != ctx->seq; ++spins) {
        if ((spins & 1023) == 1023) {
            const cudaError_t e = cudaStreamQuery(st);
            if (e == cudaSuccess) break;                   // everything has run: the word is there, or this call decided nothing
            if (e != cudaErrorNotReady) return fail(FF_E_CUDA, "stream: %s", cudaGetErrorString(e));
            if (spins > (1l << 26)) {
                FF_CUDA(cudaStreamSynchronize(st));
                break;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return FF_OK;
}

int64_t ff_workspace_bytes(int64_t seq_capacity, int64_t n_ids) {
    if (seq_capacity < 0 || n_ids < 0) return -1;
    return (int64_t)carve(nullptr, seq_capacity, n_ids).bytes;
}

int ff_build_links(ff_ctx* ctx, void* ws, int64_t ws_bytes, const int64_t* patch_type, int64_t S, int64_t n_ids,
                   void* stream) {
    return ff_build_links_for(ctx, ws, ws_bytes, patch_type, S, n_ids, 0, 0, stream);
}

int ff_build_links_for(ff_ctx* ctx, void* ws, int64_t ws_bytes, const int64_t* patch_type, int64_t S, int64_t n_ids,
                       int64_t next_row_bytes, int next_flags, void* stream) {
    Ws w;
    if (S < 0 || n_ids < 0 || S > (1ll << 30) || n_ids > (1ll << 24)) return fail(FF_E_BADARG, "bad S / n_ids");
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    ctx->cap = S;
    ctx->n_ids = n_ids;
    ctx->links_S = -1;
    if (int rc = check_ws(ctx, ws, ws_bytes, S, &w)) return rc;
    if (S > 0 && !patch_type) return fail(FF_E_BADARG, "patch_type is null");
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    const int n_chunks = (int)((S + LINK_CHUNK - 1) / LINK_CHUNK);
    const int n_b = (int)n_ids + 1;                        // chain buckets + the bucket of rows outside the chains
    FF_CUDA(cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st));
    ctx->words_clean[0] = ctx->words_clean[1] = 1;         // inside the region cleared above
    ctx->bar_base = 0;
    ctx->bar_dirty = 0;
    ctx->count_clean[0] = ctx->count_clean[1] = 1;
    ctx->links_lite = ctx->links_lite_last = 0;
    FramePlan fp;
    if (next_row_bytes > 0 && !(next_flags & 3) && S > 0 && frame_shape(ctx, n_ids, S, next_row_bytes, (next_flags & 4) != 0, &fp)) {
        // the frame-pipelined kernel will serve the next call: it needs the counters and the layout check, not the order
        k_links_uniform<<<n_chunks, LINK_CHUNK, 0, st>>>(patch_type, (int)S, (int)n_ids, w.counters[0], ctx->d_status);
        FF_LAUNCH_CHECK("k_links_uniform");
        ctx->parity = 0;
        ctx->last_parity = 0;
        ctx->links_S = S;
        ctx->have_order = 0;
        ctx->fresh_links = 1;
        ctx->links_lite = 1;
        return FF_OK;
    }
    if (S > 0) {
        k_links_hist<<<n_chunks, LINK_CHUNK, 0, st>>>(patch_type, (int)S, (int)n_ids, w.hist, w.counters[0]);
        FF_LAUNCH_CHECK("k_links_hist");
        FF_LAUNCH("k_links_colscan", k_links_colscan, (n_b + 7) / 8, 256, 0, st, w.hist, n_chunks, n_b, w.len[0], w.base,
                  w.counters[0], ctx->d_status);
        FF_LAUNCH("k_links_scatter", k_links_scatter, n_chunks, LINK_CHUNK, 0, st, patch_type, (int)S, (int)n_ids, w.hist,
                  w.base, w.order[0], w.chain[0], w.rank[0]);
        FF_LAUNCH("k_links_seq", k_links_seq, (int)((S + 255) / 256), 256, 0, st, w.rank[0], w.order[0], w.chain[0], w.counters[0],
                  (int)S, (int)(n_ids > 0 ? n_ids : 1), (unsigned long long*)&w.counters[0][C_FIRSTINV]);
    } else {
        k_links_status<<<1, 1, 0, st>>>(w.counters[0], ctx->d_status);
        FF_LAUNCH_CHECK("k_links_status");
    }
    ctx->parity = 0;
    ctx->last_parity = 0;
    ctx->links_S = S;
    ctx->have_order = 1;
    ctx->fresh_links = 1;
    return FF_OK;
}

int ff_similarity(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* hidden, int dtype, int64_t S, int64_t H, double thr,
                  void* sim_out, int64_t* order_out, void* stream) {
    Ws w;
    if (int rc = check_shape(S, H, dtype)) return rc;
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    if (int rc = links_ready(ctx, S, true)) return rc;
    if (int rc = check_ws(ctx, ws, ws_bytes, S, &w)) return rc;
    if (S > 0 && !hidden) return fail(FF_E_BADARG, "hidden is null");
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    const int bank = ctx->parity;
    if (!ctx->count_clean[bank]) FF_CUDA(cudaMemsetAsync(&w.counters[bank][C_COUNT], 0, 8, st));
    ctx->count_clean[bank] = 0;
    if (int rc = launch_similarity(w, bank, hidden, dtype, S, H, thr, st)) return rc;
    if (S > 0 && (sim_out || order_out)) {
        k_store_sim<<<(int)((S + 255) / 256), 256, 0, st>>>(w.sim, w.counters[bank], dtype, sim_out, w.order[bank], order_out);
        FF_LAUNCH_CHECK("k_store_sim");
    }
    k_count_status<<<1, 1, 0, st>>>(w.counters[bank], ctx->d_status);
    FF_LAUNCH_CHECK("k_count_status");
    ctx->last_parity = bank;
    return FF_OK;
}

int ff_merge_apply(ff_ctx* ctx, void* ws, int64_t ws_bytes, void* hidden, int dtype, int64_t S, int64_t H,
                   const int64_t* order, int64_t N, const int64_t* merge_index, int64_t M, uint8_t* keep_mask_out,
                   void* stream) {
    Ws w;
    if (int rc = check_shape(S, H, dtype)) return rc;
    if (N < 0 || M < 0 || N > S) return fail(FF_E_BADARG, "bad N / M");
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    if (ctx->cap < S) { ctx->cap = S; ctx->n_ids = 0; }
    ctx->links_S = -1;                                     // scratch use of the order arrays
    ctx->have_order = 0;
    if (int rc = check_ws(ctx, ws, ws_bytes, S, &w)) return rc;
    if (!keep_mask_out && S > 0) return fail(FF_E_BADARG, "keep_mask_out is null");
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    if (S > 0) FF_CUDA(cudaMemsetAsync(keep_mask_out, 1, (size_t)S, st));
    if (M == 0 || N == 0) return FF_OK;                    // main.py:264-266
    if (!order || !merge_index || !hidden) return fail(FF_E_BADARG, "null order / merge_index / hidden");
    FF_CUDA(cudaMemsetAsync(w.flag, 0, (size_t)N, st));
    const int64_t n = N > M ? N : M;
    k_flags_from_index<<<(int)((n + 255) / 256), 256, 0, st>>>(merge_index, (int)M, (int)N, order, w.order[0], w.flag, keep_mask_out);
    FF_LAUNCH_CHECK("k_flags_from_index");
    const bool vec = vec_ok(hidden, hidden, H, dtype);
    const int grid = (int)((N + 7) / 8);
    return dispatch_dtype(dtype, [&](auto dt) {
        constexpr int DT = decltype(dt)::value;
        if (vec) k_merge_inplace<DT, true><<<grid, 256, 0, st>>>(hidden, (int)H, (int)N, w.order[0], w.flag);
        else k_merge_inplace<DT, false><<<grid, 256, 0, st>>>(hidden, (int)H, (int)N, w.order[0], w.flag);
        FF_LAUNCH_CHECK("k_merge_inplace");
        return (int)FF_OK;
    });
}

int ff_merge_layer(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* hidden, void* hidden_out, int dtype, int64_t S,
                   int64_t H, double thr, double bound, const ff_aux* aux, int n_aux, int flags, void* stream) {
    Ws w;
    AuxPack ap;
    if (int rc = check_shape(S, H, dtype)) return rc;
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    if (int rc = links_ready(ctx, S, false)) return rc;
    if (int rc = check_ws(ctx, ws, ws_bytes, S, &w)) return rc;
    if (int rc = pack_aux(aux, n_aux, &ap)) return rc;
    if (S > 0 && (!hidden || !hidden_out)) return fail(FF_E_BADARG, "hidden / hidden_out is null");
    if (hidden == hidden_out) return fail(FF_E_BADARG, "hidden_out must not alias hidden");
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    const int bank = ctx->parity, nb = bank ^ 1;

    if (flags & 1) return fail(FF_E_BADARG, "flags bit 0 (the read-once kernel of ABI 2 .. 5) is reserved and must be 0");
    FramePlan fp;
    const bool take_frame = !(flags & 2) && frame_plan(ctx, hidden, hidden_out, dtype, S, H, thr, (flags & 4) != 0, &fp);
    if (ctx->links_lite && !take_frame)
        return fail(FF_E_BADARG, "ff_build_links_for left the links for the frame-pipelined kernel only, and this call cannot take it: call ff_build_links");
    if (!take_frame && !ctx->have_order) return fail(FF_E_BADARG, "by-patch order not in the workspace: call ff_build_links");

    if (take_frame) {
        if (int rc = launch_frame(ctx, w, bank, fp, hidden, hidden_out, dtype, S, H, thr, bound, ap, st)) return rc;
        if (ctx->ev_stop) FF_CUDA(cudaEventRecord(ctx->ev_stop, st));
        ctx->count_clean[bank] = 0;
        ctx->count_clean[nb] = 1;
        ctx->last_parity = bank;
        ctx->parity = nb;
        ctx->links_S = -2;
        ctx->have_order = 1;                               // the kernel leaves the compact by-patch arrays of the next call
        ctx->fresh_links = 0;
        ctx->links_lite_last = ctx->links_lite;
        ctx->links_lite = 0;
        return FF_OK;
    }

    if (!ctx->count_clean[bank]) FF_CUDA(cudaMemsetAsync(&w.counters[bank][C_COUNT], 0, 8, st));
    ctx->count_clean[bank] = 0;
    ctx->count_clean[nb] = 1;                              // the deciding kernel zeroes the next bank's count
    if (ctx->ev_start) FF_CUDA(cudaEventRecord(ctx->ev_start, st));
    if (int rc = launch_similarity(w, bank, hidden, dtype, S, H, thr, st)) return rc;
    DecideArgs a;
    a.counters = w.counters[bank];
    a.status = ctx->d_status;
    a.bound = bound;
    a.S = (int)S;
    a.sim = w.sim;
    a.flag = w.flag;
    a.order = w.order[bank];
    a.chain = w.chain[bank];
    a.rank = w.rank[bank];
    a.dst = w.dst[bank];
    a.srcidx = w.srcidx;
    a.rec = w.rec;
    a.order_next = w.order[nb];
    a.chain_next = w.chain[nb];
    a.rank_next = w.rank[nb];
    a.counters_next = w.counters[nb];
    a.force_branch = -1;
    a.seq = ++ctx->seq;
    if (S >= 2 * SEL_THREADS) {
        // a grid of co-resident blocks for the threshold branch; its block 0 handles the top-k branch alone
        ScanArgs sa;
        sa.d = a;
        sa.part = w.part;
        sa.barrier = w.barrier;
        int G = ctx->sm_count / 2;
        if (G > 160) G = 160;
        if (G < 1) G = 1;
        if (ctx->bar_dirty) {
            FF_CUDA(cudaMemsetAsync(w.barrier, 0, 4, st));
            ctx->bar_base = 0;
            ctx->bar_dirty = 0;
        }
        sa.bar_base = ctx->bar_base;
        ctx->bar_base += 2u * (unsigned)G;                 // every block adds 2, on either branch
        FF_LAUNCH("k_keep_scan", k_keep_scan, G, SEL_THREADS, 0, st, sa);
    } else {
        FF_LAUNCH("k_decide_scan", k_decide_scan, 1, SEL_THREADS, 0, st, a);
    }
    if (int rc = launch_merge_gather(ctx, w, bank, hidden, hidden_out, dtype, S, H, ap, st)) return rc;
    if (ctx->ev_stop) FF_CUDA(cudaEventRecord(ctx->ev_stop, st));
    ctx->last_parity = bank;
    ctx->parity = nb;
    ctx->links_S = -2;                                     // = S_keep, known once the host has synchronised
    ctx->have_order = 1;
    ctx->fresh_links = 0;
    ctx->links_lite_last = 0;
    return FF_OK;
}

int ff_importance(ff_ctx* ctx, const void* q, const void* k, int dtype, int64_t Hq, int64_t Hk, int64_t S, int64_t D,
                  int64_t num, int64_t q_hs, int64_t q_ss, int64_t k_hs, int64_t k_ss, int is_causal, double scale,
                  void* probs_out, void* scratch, int64_t scratch_bytes, void* stream) {
    if (!ctx || !q || !k || !probs_out || !scratch) return fail(FF_E_BADARG, "null argument");
    if (dtype < 0 || dtype > 2) return fail(FF_E_BADARG, "unsupported dtype %d", dtype);
    if (Hq < 1 || Hk < 1 || Hq % Hk || S < 1 || D < 1 || num < 1 || num > S) return fail(FF_E_BADARG, "bad attention shape");
    if (scratch_bytes < Hq * num * S * 4) return fail(FF_E_WORKSPACE, "scratch needs %lld bytes", (long long)(Hq * num * S * 4));
    const int64_t group = Hq / Hk;
    const size_t smem = ((size_t)group * num * D + IMP_LANES * IMP_PAD) * 4;
    if (smem > 96 * 1024) return fail(FF_E_UNSUPPORTED, "group*num*head_dim = %lld floats do not fit shared memory", (long long)(group * num * D));
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    const int64_t eb = dtype == FF_F32 ? 4 : 2;
    // the fast path: K rows of 16-byte vectors, four lanes per row with a whole number of vectors each
    const bool vec = (D * eb) % (16 * IMP_LANES) == 0 && ((uintptr_t)k & 15) == 0 && (k_hs * eb) % 16 == 0 && (k_ss * eb) % 16 == 0;
    // tensor-core path: 16-bit elements, head_dim a multiple of 32, rows the 16-byte loads can take
    const bool mma = vec && dtype != FF_F32 && D % 32 == 0 && D <= IMP_MMA_MAX_D && ((uintptr_t)q & 1) == 0;
    dim3 grid(mma ? (unsigned)((S + IMP_MMA_KEYS - 1) / IMP_MMA_KEYS)
                  : (vec ? (unsigned)((S + IMP_KEYS - 1) / IMP_KEYS) : (unsigned)((S + IMP_THREADS - 1) / IMP_THREADS)), (unsigned)Hk);
    int rc = dispatch_dtype(dtype, [&](auto dt) {
        constexpr int DT = decltype(dt)::value;
        if (mma) {
            constexpr int DM = DT == FF_F32 ? FF_BF16 : DT;  // (never taken for f32; keeps the instantiation set small)
            const size_t sm = (size_t)((group * num + 7) / 8 * 8) * (D + IMP_MMA_PAD) * 2;
            auto go = [&](auto nj) {
                constexpr int NJ = decltype(nj)::value;
                if (sm > 48 * 1024) FF_CUDA(cudaFuncSetAttribute(k_importance_logits_mma<DM, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                FF_LAUNCH("k_importance_logits_mma", (k_importance_logits_mma<DM, NJ>), grid, IMP_THREADS, sm, st, q, k, (int)Hq, (int)Hk,
                          (int)S, (int)num, q_hs, q_ss, k_hs, k_ss, is_causal, (float)scale, (float*)scratch);
                return (int)FF_OK;
            };
            int r2;
            switch (D / 32) {
                case 1: r2 = go(std::integral_constant<int, 1>()); break;
                case 2: r2 = go(std::integral_constant<int, 2>()); break;
                case 3: r2 = go(std::integral_constant<int, 3>()); break;
                case 4: r2 = go(std::integral_constant<int, 4>()); break;
                case 6: r2 = go(std::integral_constant<int, 6>()); break;
                case 8: r2 = go(std::integral_constant<int, 8>()); break;
                default: r2 = -1;
            }
            if (r2 > 0) return r2;
            if (r2 == 0) goto softmax;
        }
        if (vec) {
            if (smem > 48 * 1024) FF_CUDA(cudaFuncSetAttribute(k_importance_logits<DT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FF_LAUNCH("k_importance_logits", (k_importance_logits<DT, true>), dim3(vec ? (unsigned)((S + IMP_KEYS - 1) / IMP_KEYS) : 1, (unsigned)Hk), IMP_THREADS, smem, st, q, k, (int)Hq, (int)Hk,
                      (int)S, (int)D, (int)num, q_hs, q_ss, k_hs, k_ss, is_causal, (float)scale, (float*)scratch);
        } else {
            if (smem > 48 * 1024) FF_CUDA(cudaFuncSetAttribute(k_importance_logits<DT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FF_LAUNCH("k_importance_logits", (k_importance_logits<DT, false>), dim3((unsigned)((S + IMP_THREADS - 1) / IMP_THREADS), (unsigned)Hk), IMP_THREADS, smem, st, q, k, (int)Hq, (int)Hk,
                      (int)S, (int)D, (int)num, q_hs, q_ss, k_hs, k_ss, is_causal, (float)scale, (float*)scratch);
        }
    softmax:
        FF_LAUNCH("k_softmax_rows", k_softmax_rows<DT>, (int)(Hq * num) * SOFTMAX_CLUSTER, 1024, 0, st, (const float*)scratch, (int)S, probs_out);
        return (int)FF_OK;
    });
    return rc;
}

int ff_prune_layer(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* attn, int64_t n_rows, const void* hidden,
                   void* hidden_out, int dtype, int64_t S, int64_t H, int64_t start, int64_t length, int64_t k,
                   const ff_aux* aux, int n_aux, void* importance_out, void* stream) {
    Ws w;
    AuxPack ap;
    if (int rc = check_shape(S, H, dtype)) return rc;
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    if (ctx->cap < S) { ctx->cap = S; ctx->n_ids = 0; ctx->links_S = -1; ctx->have_order = 0; }
    if (int rc = check_ws(ctx, ws, ws_bytes, S, &w)) return rc;
    if (int rc = pack_aux(aux, n_aux, &ap)) return rc;
    if (!attn || !hidden || !hidden_out || n_rows < 1 || S < 1) return fail(FF_E_BADARG, "null / empty argument");
    if (start < 0 || length < 0 || start + length > S) return fail(FF_E_BADARG, "vision span [%lld, %lld) outside the sequence", (long long)start, (long long)(start + length));
    if (k < 0 || k > length) return fail(FF_E_BADARG, "k=%lld outside [0, %lld] (torch.topk would raise)", (long long)k, (long long)length);
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    const int bank = ctx->parity;
    const bool grid_select = S >= 2 * SEL_THREADS;
    if (ctx->ev_start) FF_CUDA(cudaEventRecord(ctx->ev_start, st));
    if (grid_select) {
        FF_CUDA(cudaMemsetAsync(w.barrier, 0, 256 + 4 * 256 * 4, st));     // barrier word + histograms (adjacent)
    }
    int rc = dispatch_dtype(dtype, [&](auto dt) {
        constexpr int DT = decltype(dt)::value;
        FF_LAUNCH("k_row_mean", k_row_mean<DT>, (int)((S + 255) / 256), 256, 0, st, attn, (int)n_rows, (int)S, w.sim);
        return (int)FF_OK;
    });
    if (rc) return rc;
    PruneArgs a;
    a.counters = w.counters[bank];
    a.status = ctx->d_status;
    a.imp = w.sim;
    a.sel = w.flag;
    a.dst = w.dst[bank];
    a.srcidx = w.srcidx;
    a.S = (int)S;
    a.start = (int)start;
    a.length = (int)length;
    a.k = k;
    a.seq = ++ctx->seq;
    if (grid_select) {
        PruneGridArgs ga;
        ga.p = a;
        ga.hist = w.sel_hist;
        ga.part = w.part;
        ga.barrier = w.barrier;
        ga.n_passes = dtype == FF_BF16 ? 2 : (dtype == FF_F16 ? 3 : 4);
        int G = ctx->sm_count / 2;
        if (G > 160) G = 160;
        if (G < 1) G = 1;
        FF_LAUNCH("k_prune_select", k_prune_select, G, SEL_THREADS, 0, st, ga);
        ctx->bar_dirty = 1;
    } else {
        FF_LAUNCH("k_prune_scan", k_prune_scan, 1, SEL_THREADS, 0, st, a);
    }
    if (vec_ok(hidden, hidden_out, H, dtype)) {
        const int64_t nvec = H * (dtype == FF_F32 ? 4 : 2) / 16;
        int rcg = dispatch_dtype(dtype, [&](auto dt) {
            constexpr int DT = decltype(dt)::value;
            FF_LAUNCH("k_merge_gather", k_merge_gather<DT>, gather_grid(S, ap.n), GATHER_WARPS * 32, 0, st,
                      hidden, hidden_out, (int)nvec, (int)S, w.srcidx, nullptr, w.order[bank], w.flag, w.counters[bank], nullptr, ap);
            return (int)FF_OK;
        });
        if (rcg) return rcg;
    } else {
        if (int rc2 = launch_merge_compact(w, bank, hidden, hidden_out, dtype, S, H, 0, st)) return rc2;
        if (ap.n) {
            k_aux_compact<<<(int)((S + 7) / 8), 256, 0, st>>>(ap, (int)S, w.dst[bank]);
            FF_LAUNCH_CHECK("k_aux_compact");
        }
    }
    if (ctx->ev_stop) FF_CUDA(cudaEventRecord(ctx->ev_stop, st));
    ctx->last_parity = bank;
    if (importance_out) {
        // counters[C_N] is not S here: store with an explicit count
        k_store_vals<<<(int)((S + 255) / 256), 256, 0, st>>>(w.sim, (int)S, dtype, importance_out);
        FF_LAUNCH_CHECK("k_store_vals");
    }
    return FF_OK;
}

int ff_compact_mask(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* mask, void* mask_out, int64_t S, int64_t S_keep,
                    int64_t elem_bytes, void* stream) {
    Ws w;
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    if (int rc = check_ws(ctx, ws, ws_bytes, S, &w)) return rc;
    if (S_keep < 0 || S_keep > S) return fail(FF_E_BADARG, "bad S_keep");
    if (S_keep == 0) return FF_OK;
    if (!mask || !mask_out) return fail(FF_E_BADARG, "null mask");
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    k_srcidx_from_dst<<<(int)((S + 255) / 256), 256, 0, st>>>(w.dst[ctx->last_parity], (int)S, w.srcidx);
    FF_LAUNCH_CHECK("k_srcidx_from_dst");
    switch (elem_bytes) {
        case 1: k_compact_mask<uint8_t><<<(int)S_keep, 256, 0, st>>>((const uint8_t*)mask, (uint8_t*)mask_out, (int)S, (int)S_keep, w.srcidx); break;
        case 2: k_compact_mask<uint16_t><<<(int)S_keep, 256, 0, st>>>((const uint16_t*)mask, (uint16_t*)mask_out, (int)S, (int)S_keep, w.srcidx); break;
        case 4: k_compact_mask<uint32_t><<<(int)S_keep, 256, 0, st>>>((const uint32_t*)mask, (uint32_t*)mask_out, (int)S, (int)S_keep, w.srcidx); break;
        case 8: k_compact_mask<uint64_t><<<(int)S_keep, 256, 0, st>>>((const uint64_t*)mask, (uint64_t*)mask_out, (int)S, (int)S_keep, w.srcidx); break;
        default: return fail(FF_E_BADARG, "elem_bytes %lld", (long long)elem_bytes);
    }
    FF_LAUNCH_CHECK("k_compact_mask");
    return FF_OK;
}

int ff_debug_frame_trace(ff_ctx* ctx, void* device_buf, int64_t bytes) {
    if (!ctx) return fail(FF_E_BADARG, "null ctx");
    if (!FR_TRACE && device_buf)
        return fail(FF_E_UNSUPPORTED, "this build carries no tracing code: build a variant with -DFR_TRACE=1 (tools/build_variant.sh) and select it with FF_LIB_PATH");
    ctx->frame_trace = (long long*)device_buf;
    ctx->frame_trace_bytes = device_buf ? bytes : 0;
    return FF_OK;
}

int ff_debug_read(ff_ctx* ctx, void* ws, int64_t ws_bytes, int what, void* dst_device, int64_t n, int dtype, void* stream) {
    Ws w;
    if (!ctx || !dst_device) return fail(FF_E_BADARG, "null argument");
    if (n < 0) return fail(FF_E_BADARG, "bad n");
    if (int rc = check_ws(ctx, ws, ws_bytes, n, &w)) return rc;
    if (n == 0) return FF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FF_DEVICE(ctx);
    const int bank = ctx->last_parity;
    const int grid = (int)((n + 255) / 256);
    switch (what) {
        case 0: k_keep_from_dst<<<grid, 256, 0, st>>>(w.dst[bank], (int)n, (uint8_t*)dst_device); break;
        case 1: k_copy_u8<<<grid, 256, 0, st>>>(w.flag, (int)n, (uint8_t*)dst_device); break;
        case 2: k_store_vals<<<grid, 256, 0, st>>>(w.sim, (int)n, dtype, dst_device); break;
        case 3:
            if (ctx->links_lite_last) return fail(FF_E_BADARG, "the by-patch order was not built (ff_build_links_for): call ff_build_links");
            k_store_order<<<grid, 256, 0, st>>>(w.order[bank], (int)n, (int64_t*)dst_device); break;
        default: return fail(FF_E_BADARG, "what=%d", what);
    }
    FF_LAUNCH_CHECK("debug_read");
    return FF_OK;
}

}  // extern "C"
