"""GPU: the frame-pipelined merge kernel (csrc/ff_frame.cuh) — that it IS the kernel serving the first merge call of a
uniform video, that it equals the multi-kernel path bit for bit (hidden_states, rotary rows, patch_type, keep mask,
similarities) over the shapes it takes (dtypes, chains per CTA, ring depths, short videos, long text spans), that its
outputs feed the calls behind it, that layouts and branches it cannot serve are redone on the multi-kernel path, and
that repeated launches are identical (its warps hand over through mbarriers: a race would show here)."""
import numpy as np
import pytest
import torch

from _harness import set_mode, t2f
from oracle import ff_oracle as orc
from framefusion_b200 import _lib, synth
from framefusion_b200.main import FrameFusion

pytestmark = pytest.mark.gpu


def first_call(wl, mode, cost=0.3, slb=0.6, rlb=0.1, debug=True):
    ff = FrameFusion(cost, slb, rlb)
    set_mode(ff, mode)
    ff.debug_trace = debug          # (with it the by-patch order is built in full; without, ff_build_links_for only checks the layout)
    ff.prepare(*wl.prepare_args())
    h, pos, _ = ff(wl.hidden, [wl.cos, wl.sin], None)
    torch.cuda.synchronize()
    st = ff._state(wl.hidden.device)
    return ff, h, pos, int(st.status[_lib.ST_FUSED])


def same(a, b):
    return a.shape == b.shape and torch.equal(a, b)


@pytest.mark.parametrize("frames,patches,hidden,dtype", [
    (16, 576, 3584, torch.bfloat16),        # C2's row and chain layout, four chains per CTA
    (12, 729, 4096, torch.bfloat16),        # C4's: five chains per CTA (the build for up to eight), a ring of four frames
    (9, 196, 1024, torch.float32),          # C1's: float32 rows
    (20, 210, 3584, torch.float16),         # 210 tokens per frame (two chains per CTA, 105 CTAs), f16
    (33, 49, 896, torch.bfloat16),          # fewer chains than SMs: one chain per CTA
    (1, 144, 512, torch.bfloat16),          # one frame: nothing to compare, everything kept
    (2, 144, 512, torch.bfloat16),
    (40, 300, 1000 * 8 // 8, torch.bfloat16),   # rows of 2000 bytes (a multiple of 16, not of 128)
], ids=["c2rows", "c4rows", "f32", "f16_210", "p49", "one_frame", "two_frames", "rows2000"])
def test_frame_kernel_equals_the_multi_kernel_path(frames, patches, hidden, dtype):
    wl = synth.to_device(synth.make_workload(frames, patches, hidden, dtype, seed=5, r_lo=0.0, r_hi=1.0), "cuda")
    ff_m, h_m, pos_m, k_m = first_call(wl, "multi")
    ff_f, h_f, pos_f, k_f = first_call(wl, "frame")
    assert k_m == 0 and k_f == 2, "the frame-pipelined kernel did not serve the call"
    assert same(h_f, h_m) and same(pos_f[0], pos_m[0]) and same(pos_f[1], pos_m[1]) and same(ff_f.patch_type, ff_m.patch_type)
    tf, tm = ff_f.last_trace, ff_m.last_trace
    assert np.array_equal(tf["keep_mask"], tm["keep_mask"]) and np.array_equal(tf["order"], tm["order"])
    assert np.array_equal(tf["merge_index"], tm["merge_index"])
    # both kernels sum the row products in float32 in their own order: identical except on a rounding boundary of T
    if dtype == torch.float32:
        assert np.allclose(tf["sim_values"], tm["sim_values"], rtol=1e-5, atol=1e-6)
    else:
        assert (tf["sim_values"] != tm["sim_values"]).mean() < 2e-3
    assert ff_f.sparsity_list == ff_m.sparsity_list and ff_f.finish_merging == ff_m.finish_merging


def test_frame_kernel_feeds_the_calls_behind_it():
    """merge (frame kernel) -> merge (multi-kernel path on the arrays the frame kernel left) -> ... -> prune, against the
    oracle call by call."""
    wl = synth.make_workload(24, 300, 1024, torch.bfloat16, seed=11, r_lo=0.0, r_hi=1.0)
    ff = FrameFusion(0.3, 0.6, 0.1)
    ff.use_frame = "force"
    ff.prepare(*synth.to_device(wl, "cuda").prepare_args())
    o = orc.OracleFrameFusion(0.3, 0.6, 0.1, "bf16")
    o.prepare(wl.patch_type.numpy(), wl.patch_num, *wl.prepare_args()[2:])
    h, pos = wl.hidden.cuda(), [wl.cos.cuda(), wl.sin.cuda()]
    kernels = []
    for c in range(6):
        if ff.finish_merging and ff.finish_pruning:
            break
        if c > 0:
            h = synth.apply_drift(h, 0.35, 11, c)
        attn = synth.make_attention_row(h.shape[1], n_heads=28, num=1, dtype=torch.bfloat16, seed=c).cuda() if ff.finish_merging else None
        h_in, p_in = t2f(h[0]), [t2f(pos[0][0]), t2f(pos[1][0])]
        merging = not ff.finish_merging
        h, pos, _ = ff(h, pos, None, attn)
        torch.cuda.synchronize()
        if merging:
            kernels.append(int(ff._state(h.device).status[_lib.ST_FUSED]))
        want_h, want_p, _ = o.forward(h_in, p_in, None, None if attn is None else t2f(attn[0]))
        assert h.shape[1] == want_h.shape[0] and np.array_equal(t2f(h[0]), want_h), f"call {c}"
        assert np.array_equal(t2f(pos[0][0]), want_p[0]) and np.array_equal(ff.patch_type[0].cpu().numpy(), o.patch_type)
    assert kernels[0] == 2 and len(kernels) >= 2 and all(k == 0 for k in kernels[1:]), kernels
    assert ff.finish_pruning and ff.sparsity_list == o.sparsity_list


def run_against_oracle(wl, cost=0.3, slb=0.6):
    ff = FrameFusion(cost, slb, 0.1)
    ff.use_frame = "force"
    ff.prepare(*synth.to_device(wl, "cuda").prepare_args())
    o = orc.OracleFrameFusion(cost, slb, 0.1, "bf16")
    o.prepare(wl.patch_type.numpy(), wl.patch_num, *wl.prepare_args()[2:])
    h, pos, _ = ff(wl.hidden.cuda(), [wl.cos.cuda(), wl.sin.cuda()], None)
    torch.cuda.synchronize()
    want_h, want_p, _ = o.forward(t2f(wl.hidden[0]), [t2f(wl.cos[0]), t2f(wl.sin[0])], None)
    assert h.shape[1] == want_h.shape[0] and np.array_equal(t2f(h[0]), want_h)
    assert np.array_equal(t2f(pos[0][0]), want_p[0]) and np.array_equal(ff.patch_type[0].cpu().numpy(), o.patch_type)
    assert ff.sparsity_list == o.sparsity_list and (ff.finish_merging, ff.finish_pruning) == (o.finish_merging, o.finish_pruning)
    return ff, o, int(ff._state(h.device).status[_lib.ST_FUSED])


def test_top_k_branch_is_redone_on_the_multi_kernel_path():
    """Highly similar frames: count / n_vis >= the bound, the reference takes the top-k branch (main.py:116-127).  The
    frame kernel speculated on the threshold branch, reports it, and the call is redone."""
    wl = synth.make_workload(24, 128, 512, torch.bfloat16, seed=9, r_lo=0.8, r_hi=1.0)
    ff, o, kernel = run_against_oracle(wl)
    assert o.last["branch"] == "topk" and kernel == 0


@pytest.mark.parametrize("layout", ["two_spans", "shuffled_ids", "ragged_tail", "id_out_of_range"])
def test_layouts_that_are_not_a_uniform_video_are_redone(layout):
    wl = synth.make_workload(12, 96, 512, torch.bfloat16, seed=4, r_lo=0.0, r_hi=1.0, n_pre=5, n_post=7)
    pt = wl.patch_type.clone()
    first = 5
    if layout == "two_spans":
        pt[0, first + 96 * 6] = -1                          # a text token inside the video
    elif layout == "shuffled_ids":
        pt[0, first + 3], pt[0, first + 4] = pt[0, first + 4].clone(), pt[0, first + 3].clone()
    elif layout == "ragged_tail":
        pt[0, first + 96 * 11 + 40:first + 96 * 12] = -1    # the last frame is cut short
    else:
        pt[0, first + 96 * 2 + 7] = 96                      # an id outside [0, patch_num): a vision token outside the chains
    wl.patch_type = pt
    ff, o, kernel = run_against_oracle(wl)
    assert kernel == 0


def test_threshold_at_the_sentinel_never_takes_the_frame_kernel():
    wl = synth.make_workload(10, 64, 512, torch.bfloat16, seed=2, r_lo=0.0, r_hi=1.0)
    ff, o, kernel = run_against_oracle(wl, cost=0.99, slb=-2.0)
    assert kernel == 0


def test_long_text_spans_around_the_video():
    wl = synth.to_device(synth.make_workload(20, 144, 1024, torch.bfloat16, seed=6, r_lo=0.0, r_hi=1.0, n_pre=700, n_post=900), "cuda")
    ff_m, h_m, pos_m, k_m = first_call(wl, "multi")
    ff_f, h_f, pos_f, k_f = first_call(wl, "frame")
    assert k_f == 2 and same(h_f, h_m) and same(pos_f[0], pos_m[0]) and same(ff_f.patch_type, ff_m.patch_type)


def test_position_tensors_of_other_shapes_travel_with_the_rows():
    """M-RoPE style [3, 1, S, D] rotary tensors and int64 position ids as aux tensors of the frame kernel."""
    wl = synth.to_device(synth.make_workload(14, 160, 768, torch.bfloat16, seed=8, r_lo=0.0, r_hi=1.0), "cuda")
    S = wl.seq_len
    cos3 = torch.stack([wl.cos, wl.cos * 0.5, wl.cos * 0.25])           # [3, 1, S, D]
    sin3 = torch.stack([wl.sin, wl.sin * 0.5, wl.sin * 0.25])
    outs = {}
    for mode in ("multi", "frame"):
        ff = FrameFusion(0.3, 0.6, 0.1)
        set_mode(ff, mode)
        ff.prepare(*wl.prepare_args())
        h, pos, _ = ff(wl.hidden, [cos3, sin3], None)
        torch.cuda.synchronize()
        outs[mode] = (h, pos, int(ff._state(h.device).status[_lib.ST_FUSED]))
    assert outs["frame"][2] == 2 and outs["multi"][2] == 0
    assert same(outs["frame"][0], outs["multi"][0])
    assert same(outs["frame"][1][0], outs["multi"][1][0]) and same(outs["frame"][1][1], outs["multi"][1][1])
    assert outs["frame"][1][0].shape[0] == 3 and outs["frame"][1][0].shape[2] == outs["frame"][0].shape[1] < S


def test_repeated_launches_are_identical():
    wl = synth.to_device(synth.make_workload(48, 576, 3584, torch.bfloat16, seed=0), "cuda")
    _ff, h0, pos0, k0 = first_call(wl, "multi")
    for it in range(40):
        ff, h, pos, k = first_call(wl, "frame")
        assert k == 2
        assert same(h, h0) and same(pos[0], pos0[0]) and same(pos[1], pos0[1]), f"launch {it} differs"


def test_the_library_picks_the_frame_kernel_where_it_is_the_faster_one():
    """Default settings: 576 patches of 7-KB rows (C2 / C3) and 729 patches of 8-KB rows (C4: five chains per SM, a ring of
    five frames) take the frame kernel; 210 tokens per frame and float32 C1 rows stay on the multi-kernel path
    (profiles/r02_sweep.jsonl)."""
    for frames, patches, hidden, dtype, want in ((8, 576, 3584, torch.bfloat16, 2), (8, 210, 3584, torch.bfloat16, 0),
                                                 (6, 729, 4096, torch.bfloat16, 2), (8, 196, 1024, torch.float32, 0)):
        wl = synth.to_device(synth.make_workload(frames, patches, hidden, dtype, seed=1), "cuda")
        ff = FrameFusion(0.3, 0.6, 0.1)
        ff.prepare(*wl.prepare_args())
        ff(wl.hidden, [wl.cos, wl.sin], None)
        torch.cuda.synchronize()
        assert int(ff._state(wl.hidden.device).status[_lib.ST_FUSED]) == want, (frames, patches, hidden)


def test_links_left_for_the_frame_kernel_only():
    """Without debug_trace the host announces the merge call to ff_build_links_for, which then skips the counting sort: same
    outputs, four launches fewer, and the library refuses what it has no links for."""
    lib = _lib.load()
    wl = synth.to_device(synth.make_workload(10, 576, 1024, torch.bfloat16, seed=12, r_lo=0.0, r_hi=1.0), "cuda")
    _ffm, h_m, pos_m, k_m = first_call(wl, "multi")
    n0 = lib.ff_launch_count()
    ff_full, h_full, pos_full, k_full = first_call(wl, "frame", debug=True)
    n1 = lib.ff_launch_count()
    ff_lite, h_lite, pos_lite, k_lite = first_call(wl, "frame", debug=False)
    n2 = lib.ff_launch_count()
    assert k_full == 2 and k_lite == 2
    assert same(h_lite, h_m) and same(pos_lite[0], pos_m[0]) and same(pos_lite[1], pos_m[1]) and same(ff_lite.patch_type, ff_full.patch_type)
    # full: 4 link kernels + the merge kernel + 4 debug reads; lite: 1 + 1
    assert n2 - n1 == 2, (n1 - n0, n2 - n1)
    # a second merge call runs the multi-kernel path on the arrays the frame kernel left
    h2, pos2, _ = ff_lite(h_lite, pos_lite, None)
    h2m, pos2m, _ = ff_full(h_full, pos_full, None)
    torch.cuda.synchronize()
    assert same(h2, h2m) and same(pos2[0], pos2m[0])
    # straight through the C ABI: after ff_build_links_for nothing but the frame kernel may follow
    st = ff_lite._state(wl.hidden.device)
    wp, wb = st.ws_ptr()
    pt = wl.patch_type.reshape(-1).to(torch.int64).contiguous()
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.ff_build_links_for(st.ctx, wp, wb, pt.data_ptr(), wl.seq_len, 576, 2048, 4, stream) == 0     # (4: wherever it can run)
    out = torch.empty_like(wl.hidden)
    rc = lib.ff_merge_layer(st.ctx, wp, wb, wl.hidden.data_ptr(), out.data_ptr(), 0, wl.seq_len, 1024, 0.6, 0.7, None, 0, 2, stream)
    assert rc == -1 and b"ff_build_links" in lib.ff_last_error()
    torch.cuda.synchronize()
