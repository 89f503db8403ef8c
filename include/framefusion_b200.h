/*
 * framefusion_b200 — C ABI of the B200-native FrameFusion token-reduction path.
 *
 * One shared library (libframefusion_b200.so, built from framefusion_b200/csrc/ for sm_100a).
 * Plain pointers and sizes only: no torch / C++ types cross this boundary.  All device work is
 * enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*); nothing here
 * synchronises the stream — the caller does that once, then reads the status block.
 *
 * The reference (thu-nics/FrameFusion) is pure Python/PyTorch and has no FFI; each entry point below
 * replaces a span of ATen ops in /root/reference/framefusion/main.py (cited per function) and is what a
 * reference-side binding (ctypes) calls instead of those ops.
 *
 * Conventions
 *   - B = 1 (the reference asserts it, main.py:203).  hidden is row-major contiguous [S, H].
 *   - dtype: FF_BF16 / FF_F16 / FF_F32 — the dtype of hidden_states ("T" below).
 *   - every function returns 0 on success, a negative FF_E* code otherwise; ff_last_error() gives the
 *     text for the calling thread.
 *   - `ws` is a caller-allocated device workspace of at least ff_workspace_bytes(capacity, n_ids) bytes,
 *     256-byte aligned.  It carries the chain links (by-patch order) from one call to the next, so one
 *     workspace belongs to one FrameFusion object / one request at a time.
 *   - ff_ctx owns a pinned, device-mapped status block; kernels write results there (sizes, counts, the
 *     branch taken) so that a reducing call costs exactly one host<->device synchronisation.
 */
#ifndef FRAMEFUSION_B200_H
#define FRAMEFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FF_ABI_VERSION 6

enum ff_dtype { FF_BF16 = 0, FF_F16 = 1, FF_F32 = 2 };

enum ff_error {
    FF_OK = 0,
    FF_E_BADARG = -1,      /* null pointer, negative size, unsupported dtype, misaligned workspace */
    FF_E_WORKSPACE = -2,   /* workspace too small for this call */
    FF_E_CUDA = -3,        /* a CUDA runtime call or kernel launch failed */
    FF_E_UNSUPPORTED = -4  /* shape outside what the kernels handle (see ff_last_error) */
};

/* status block layout (int64 slots, host-visible after the stream has been synchronised) */
enum ff_status_slot {
    FF_ST_SEQ_KEEP = 0,    /* S_keep: rows in the compacted outputs */
    FF_ST_COUNT = 1,       /* #{j : sim[j] >= T(similarity_lower_bound)}            (main.py:113) */
    FF_ST_NVIS = 2,        /* #{i : patch_type[i] != -1}                            (main.py:112) */
    FF_ST_NCHAIN = 3,      /* N: tokens whose patch id is in [0, n_ids)             (main.py:208-214) */
    FF_ST_BRANCH = 4,      /* 0 = threshold branch, 1 = top-k branch                (main.py:116-127) */
    FF_ST_TOPK = 5,        /* k used by the branch that ran (merge top-k or prune top-k) */
    FF_ST_ERROR = 6,       /* device-side error: 0 ok, 1 n_vis == 0, 2 k > N, 3 the single-launch kernel (FUSED = 2) could
                              not serve the call — it speculated on the threshold branch and the count says top-k, or the
                              sequence is not the uniform video the frame-pipelined kernel is built for: hidden is
                              untouched, rebuild the links and call again with flags = 2 */
    FF_ST_NMERGED = 7,     /* tokens merged away by this call */
    FF_ST_FUSED = 8,       /* 0 multi-kernel path, 2 the frame-pipelined kernel (first merge call of a prefill on a uniform
                              video); 1 was the read-once kernel of ABI 2 .. 5, removed in ABI 6 */
    FF_ST_SEQ = 10,        /* number of the reducing call whose results the block holds: written LAST, after a system-wide
                            * fence, by the kernel that decides the call (ff_status_wait) */
    FF_ST_INTERNAL = 9,    /* 1 if a wait inside the single-launch kernel gave up (the results of the call are invalid) */
    FF_ST_SLOTS = 16
};

typedef struct ff_ctx ff_ctx;

/* aux tensor compacted along the sequence axis together with hidden (cos, sin, patch_type, position ids):
 * `planes` slabs of [S, row_bytes] at src (slab stride src_plane_stride bytes) -> slabs of
 * [S_keep, row_bytes] at dst (slab stride dst_plane_stride bytes).        (main.py:132, 142-178) */
typedef struct ff_aux {
    const void* src;
    void* dst;
    int64_t planes;
    int64_t src_plane_stride;
    int64_t dst_plane_stride;
    int64_t row_bytes;
} ff_aux;

#define FF_MAX_AUX 6

int ff_abi_version(void);
const char* ff_last_error(void);
/* number of kernels this library has launched since it was loaded (process-wide; bench.py's gpu_launches) */
int64_t ff_launch_count(void);

/* ---- context -------------------------------------------------------------------------------- */
int ff_ctx_create(int device, ff_ctx** out);
int ff_ctx_destroy(ff_ctx* ctx);
/* host pointer to FF_ST_SLOTS int64 values */
const int64_t* ff_ctx_status(const ff_ctx* ctx);
/* waits for the work enqueued on `stream`: the one synchronisation a reducing call needs before the status block
 * (S_keep, the branch taken) can be read */
int ff_stream_sync(ff_ctx* ctx, void* stream);
/* Returns as soon as the status block holds the results of the last ff_merge_layer / ff_prune_layer call enqueued on
 * `stream` — S_keep, the count, the branch — which the deciding kernel writes BEFORE the gather of that call runs: the host
 * learns the output length (the reference learns it from `.item()` / a boolean-mask select, main.py:112-138) while the rows
 * are still being moved, and its work for the next call overlaps them.  Everything enqueued later on the same stream is
 * ordered behind the call as usual; a host read of the outputs needs its own synchronisation (ff_stream_sync, or any
 * stream-ordered copy).  Falls back to waiting for the stream if the status does not show up. */
int ff_status_wait(ff_ctx* ctx, void* stream);
/* Profiling aid: two caller-owned cudaEvent_t (timing enabled), or NULLs to switch it off.  While set, ff_merge_layer and
 * ff_prune_layer record `ev_start` on the stream right before their first kernel launch and `ev_stop` right after their
 * last one, so the elapsed time between them is the GPU time of the call's launches and nothing of the host's path to
 * the first launch (bench.py's roofline.kernel_us). */
int ff_ctx_timing(ff_ctx* ctx, void* ev_start, void* ev_stop);

int64_t ff_workspace_bytes(int64_t seq_capacity, int64_t n_ids);

/* ---- chain links: replaces the eq-broadcast + nonzero "sort by patch" of main.py:208-214 ------
 * patch_type [S] int64 on the device; n_ids = ceil(patch_num).  Leaves in `ws` the by-patch order, the
 * patch id of every by-patch position and the inverse map; status: NCHAIN, NVIS. */
int ff_build_links(ff_ctx* ctx, void* ws, int64_t ws_bytes, const int64_t* patch_type, int64_t seq_len,
                   int64_t n_ids, void* stream);

/* The same, told what comes next: the caller's next call on this context is ff_merge_layer with rows of
 * `next_row_bytes` bytes and `next_flags`.  If the library will serve that call with the frame-pipelined kernel (see
 * ff_merge_layer), the by-patch order is not needed — the kernel walks frames, not sorted positions — and this call only
 * counts the tokens and checks the layout (one small kernel instead of the four of the counting sort: ~25 us at
 * 36 898 tokens).  ff_merge_layer then refuses anything but that kernel (FF_E_BADARG: call ff_build_links), ff_similarity
 * and ff_debug_read(what = 3) likewise; if the kernel reports ERROR = 3 the caller rebuilds the links with ff_build_links
 * and redoes the call with flags = 2, as always.  next_row_bytes = 0: exactly ff_build_links. */
int ff_build_links_for(ff_ctx* ctx, void* ws, int64_t ws_bytes, const int64_t* patch_type, int64_t seq_len,
                       int64_t n_ids, int64_t next_row_bytes, int next_flags, void* stream);

/* ---- similarity: replaces main.py:216-238 + cosine_similarity (main.py:345-349) ----------------
 * Needs links for this sequence in `ws`.  Writes sim_out [N] in T (-2 at chain heads) and, if not null,
 * order_out [N] int64.  thr = the similarity lower bound already rounded to T (as a double); the count of
 * sim >= thr goes to status COUNT and the flags stay in `ws`. */
int ff_similarity(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* hidden, int dtype, int64_t seq_len,
                  int64_t hidden_size, double thr, void* sim_out, int64_t* order_out, void* stream);

/* ---- merge (in place) + keep mask: replaces merge_tokens_and_get_mask, main.py:243-319 ----------
 * order [N] int64, merge_index [M] int64 ascending by-patch positions (both on the device).  hidden is
 * modified in place at anchor rows exactly as the reference does; keep_mask_out [S] bytes (0/1). */
int ff_merge_apply(ff_ctx* ctx, void* ws, int64_t ws_bytes, void* hidden, int dtype, int64_t seq_len,
                   int64_t hidden_size, const int64_t* order, int64_t n_chain, const int64_t* merge_index,
                   int64_t n_merge, uint8_t* keep_mask_out, void* stream);

/* ---- one merge-stage call of FrameFusion.forward, fused: main.py:104-138 -----------------------
 * similarity -> count -> branch (r = count / n_vis < bound ? threshold : top-k with k = int(bound*n_vis))
 * -> runs -> merged rows -> compaction of hidden and of every aux tensor, and the links for the next call.
 * hidden [S,H] is read only; hidden_out must hold S rows (only S_keep are written).  patch_type is aux[0]
 * by convention of the host wrapper but the library does not care.  status: SEQ_KEEP, COUNT, NVIS, NCHAIN,
 * BRANCH, TOPK, ERROR, NMERGED, FUSED, INTERNAL.  `flags`: bit 0 is reserved and must be 0 (the read-once kernel of ABI 2 .. 5:
 * FF_E_BADARG);
 * bit 1 = do NOT use the frame-pipelined kernel, bit 2 = use it wherever it can run.  Without bit 1 the first merge call
 * after ff_build_links runs as ONE launch in which every row travels HBM -> shared memory -> HBM once
 * (csrc/ff_frame.cuh) when the shape is one it is faster on (at least 20 KB of rows per frame and SM, chains on at
 * least nine SMs in ten, a ring of five frames in shared memory: 300 .. 592 patches of 7-KB rows, 729 of 8-KB rows on a B200; bit 2 drops
 * these three conditions); the kernel checks on the device that the sequence is a uniform video (one span of chain rows,
 * patch ids 0 .. n_ids-1 repeating) and that the threshold branch applies, and reports ERROR = 3 / FUSED = 2 otherwise. */
int ff_merge_layer(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* hidden, void* hidden_out, int dtype,
                   int64_t seq_len, int64_t hidden_size, double thr, double bound, const ff_aux* aux,
                   int n_aux, int flags, void* stream);

/* ---- importance: replaces utils.scaled_dot_product_attention, utils.py:27-57 -------------------
 * q [Hq, S, D] / k [Hk, S, D] with element strides (head, seq; last dim contiguous); only the last `num`
 * queries are used.  probs_out [Hq, num, S] in T.  Hq % Hk == 0 (GQA aware: K is read once).
 * scratch: device buffer of at least Hq*num*S*4 bytes (float32 logits). */
int ff_importance(ff_ctx* ctx, const void* q, const void* k, int dtype, int64_t n_q_heads, int64_t n_kv_heads,
                  int64_t seq_len, int64_t head_dim, int64_t num, int64_t q_head_stride, int64_t q_seq_stride,
                  int64_t k_head_stride, int64_t k_seq_stride, int is_causal, double scale, void* probs_out,
                  void* scratch, int64_t scratch_bytes, void* stream);

/* ---- one prune-stage call: main.py:61-101 ------------------------------------------------------
 * attn [n_rows, S] in T (n_rows = heads*num): importance = T(mean over rows); keeps [0,start), the top-k of
 * [start, start+length) (ties: lowest index first) and [start+length, S); compacts hidden and aux.
 * importance_out [S] in T may be null.  status: SEQ_KEEP, TOPK. */
int ff_prune_layer(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* attn, int64_t n_rows,
                   const void* hidden, void* hidden_out, int dtype, int64_t seq_len, int64_t hidden_size,
                   int64_t start, int64_t length, int64_t k, const ff_aux* aux, int n_aux,
                   void* importance_out, void* stream);

/* ---- 4-D mask compaction mask[keep][:, keep]: main.py:100, 138 ---------------------------------
 * Uses the destination map the last merge/prune call left in `ws`.  mask [S, S] -> mask_out [S_keep, S_keep]. */
int ff_compact_mask(ff_ctx* ctx, void* ws, int64_t ws_bytes, const void* mask, void* mask_out, int64_t seq_len,
                    int64_t seq_keep, int64_t elem_bytes, void* stream);

/* ---- introspection of the last merge call (tests, static API): copies device arrays out of `ws` ------
 * what: 0 = keep mask by sequence position (uint8 [S]), 1 = merge flags by by-patch position (uint8 [N]),
 *       2 = sim (T [N] by by-patch position), 3 = order (int64 [N]; needs ff_build_links, not the short form of
 *       ff_build_links_for) */
int ff_debug_read(ff_ctx* ctx, void* ws, int64_t ws_bytes, int what, void* dst_device, int64_t n, int dtype,
                  void* stream);

/* ---- measurement aid (libraries built with -DFR_TRACE=1 only; the shipped build returns FF_E_UNSUPPORTED for a non-null
 * buffer: its kernel carries no tracing code): the next frame-pipelined launch writes %globaltimer stamps into device_buf, laid out
 * [CTA][frame][8] int64 (0 load requested, 1 similarity done, 2 destinations known, 3 rows out, 4 aux rows out);
 * null switches it off.  The buffer must stay alive until that launch has run. */
int ff_debug_frame_trace(ff_ctx* ctx, void* device_buf, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif
