"""GPU run of the EDLines kernels (csrc/edlines.cu) against the oracle (oracle/oracle_edlines.cpp), through the C ABI: segment lists,
chain statistics, both blur generations, and the descriptors of detect_descrip_lines with use_LSD = false -- all bit for bit."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("n,w,h,seed,kw", [(5, 640, 480, 3, {}), (2, 641, 479, 4, dict(texture=1.0, noise_sigma=5.0)), (2, 1242, 375, 5, {}), (3, 97, 61, 6, {})])
def test_segments_match_the_oracle(ctx, n, w, h, seed, kw):
    from cube_slam_wu_b200 import synth
    frames = synth.make_lsd_frames(n, w, h, seed=seed, **kw)
    for filt, thr in ((True, 15.0), (False, 0.0)):
        lines, st = ctx.edlines_detect_batch(frames, line_length_thres=thr, filter=filt)
        assert st.n_frames_failed == 0
        n_ref = 0
        for f in range(n):
            ref, _ = O.edlines_detect(frames[f], filter=filt, length_thres=thr)
            got = lines[f]
            assert got.shape == ref.shape, "%dx%d frame %d: %d vs %d segments" % (w, h, f, len(got), len(ref))
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "%dx%d frame %d: segments differ" % (w, h, f)
            n_ref += len(ref)
        assert st.n_lines == n_ref
    chains = [O.edlines_chains(fr) for fr in frames]
    assert st.n_chains == sum(len(c) for c in chains) and st.n_chain_px == sum(sum(len(x) for x in c) for c in chains)


def test_blur_generation_3(ctx):
    """csb_set_blur_generation(3): the 8-bit Gaussian taps of OpenCV <= 3.4.0 in front of LBD / EDLines (gradient images bit-exact, segments identical)"""
    from cube_slam_wu_b200 import synth
    frames = synth.make_lsd_frames(2, 333, 251, seed=31, texture=1.0, noise_sigma=4.0)
    frames[0, 20:80, 30:120] = 255          # a saturated patch: the generation-3 taps sum to 257
    ctx.set_blur_generation(3); O.lbd_set_blur_generation(3)
    try:
        ctx.lbd_upload(frames, [np.zeros((0, 4), np.float32)] * 2); ctx.lbd_run()
        for f in range(2):
            dx, dy = ctx.lbd_debug_gradients(f, (251, 333))
            _, rdx, rdy = O.lbd_gradients(frames[f])
            assert np.array_equal(dx, rdx) and np.array_equal(dy, rdy), "generation-3 gradients differ"
        lines, st = ctx.edlines_detect_batch(frames)
        for f in range(2):
            ref, _ = O.edlines_detect(frames[f])
            assert np.array_equal(lines[f].view(np.uint32), ref.view(np.uint32)), "generation-3 segments differ"
    finally:
        ctx.set_blur_generation(4); O.lbd_set_blur_generation(4)


def test_detect_descrip_lines_edlines_branch(ctx):
    """detect_descrip_lines with use_LSD = false: descriptors of the key lines, from the detector's own fields (csb_edlines_describe)"""
    from cube_slam_wu_b200 import synth
    frames = synth.make_lsd_frames(3, 640, 480, seed=9)
    out = ctx.edlines_detect_describe_batch(frames, want_float=True)
    n_desc = 0
    for f in range(3):
        ref, extra = O.edlines_detect(frames[f])
        assert np.array_equal(out["lines"][f].view(np.uint32), ref.view(np.uint32))
        r72, r32 = O.lbd_describe_keylines(frames[f], ref, extra[:, 0], extra[:, 1])
        assert np.array_equal(out["desc"][f], r32), "frame %d: binary descriptors differ" % f
        g72 = out["desc_float"][f]
        assert np.array_equal(np.isnan(g72), np.isnan(r72)) and np.array_equal(np.nan_to_num(g72).view(np.uint32), np.nan_to_num(r72).view(np.uint32))
        n_desc += len(ref)
    assert n_desc > 50


def test_reference_image_407(ctx, csb):
    d = np.load(os.path.join(GOLD, "lsd_407.npz"))
    lines, st = ctx.edlines_detect_batch(d["gray"][None])
    ref, _ = O.edlines_detect(d["gray"])
    assert np.array_equal(lines[0].view(np.uint32), ref.view(np.uint32)) and len(ref) > 100
    # the host-side mirror of class line_lbd_detect, EDLines branch (what object_slam selects, main_obj.cpp:503-505)
    det = csb.line_lbd_detect(ctx)
    det.use_LSD = False
    det.line_length_thres = 15.0
    got = det.detect_filter_lines(d["gray"])
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
