"""GPU parity of the LSD line detector (SURVEY.md 8 "next" row f-2): csb_lsd_detect_batch vs the CPU oracle (oracle/oracle_lsd.cpp,
reference seed order), through the C ABI.

Bar: the streaming stages (blurred + down-scaled image, gradient norm, level-line angle) are bit-exact FP64; the segment lists have the
same length and order, endpoints within 1e-4 px (they are expected to be bit-identical: everything that decides region membership is
specified arithmetic; only the NFA's log/exp/pow come from different libms and they feed comparisons only)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, frames, **kw):
    lines, st = ctx.lsd_detect_batch(frames, **kw)
    kw.pop("unit_link_deg", None)
    mode = 1 if kw.get("filter", True) else 0
    thr = kw.get("line_length_thres", 15.0)
    worst = 0.0
    total = 0
    for f in range(len(frames)):
        ref = oracle.lsd_detect(frames[f], seed_order=0, libm_trig=0, refine=2, mode=mode, length_thres=thr)
        got = lines[f]
        assert got.shape == ref.shape, "frame %d: %d segments on the GPU, %d in the oracle" % (f, len(got), len(ref))
        if len(ref):
            worst = max(worst, float(np.abs(got - ref).max()))
        total += len(ref)
    assert worst <= 1e-4, "segment endpoints differ by %g px" % worst
    assert st.n_lines == total
    return lines, st, worst


def test_maps_bit_exact(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    for (w, h) in ((640, 480), (1242, 375), (333, 251)):
        frames = synth.make_lsd_frames(2, w, h, seed=5)
        ctx.lsd_upload(frames)
        ctx.lsd_run()
        lines, st = ctx.lsd_download()
        assert (st.scaled_width, st.scaled_height) == (int(np.rint(w * 0.8)), int(np.rint(h * 0.8)))
        for f in range(2):
            sc, mg, an = ctx.lsd_debug_maps(f, (st.scaled_height, st.scaled_width))
            rsc, rmg, ran = oracle.lsd_maps(frames[f])
            assert np.array_equal(sc, rsc), "scaled image differs, max %g" % np.abs(sc - rsc).max()
            assert np.array_equal(mg[:-1, :-1], rmg[:-1, :-1]), "gradient norm differs"
            assert np.array_equal(an, ran), "level-line angles differ in %d px" % int((an != ran).sum())


def test_segments_vga(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    frames = synth.make_lsd_frames(4, 640, 480, seed=11)
    lines, st, worst = _check(ctx, oracle, frames)
    assert st.n_lines > 100 and st.n_regions > st.n_lines
    print("lsd vga: %d segments, %d regions, %d region px, max diff %g" % (st.n_lines, st.n_regions, st.n_region_px, worst))


def test_segments_kitti_raw_and_textured(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    frames = synth.make_lsd_frames(2, 1242, 375, seed=12, texture=1.0, noise_sigma=5.0)
    _check(ctx, oracle, frames, filter=False)  # every segment LineSegmentDetectorImpl::detect returns
    _check(ctx, oracle, frames, line_length_thres=50.0)  # the class default (line_lbd_allclass.cpp:126)


def test_edge_cases(ctx, csb, oracle):
    flat = np.full((1, 64, 80), 128, np.uint8)
    lines, st = ctx.lsd_detect_batch(flat)
    assert len(lines[0]) == 0 and st.n_regions == 0
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (2, 97, 131)).astype(np.uint8)  # odd sizes, every pixel a seed candidate
    _check(ctx, oracle, noise, filter=False)
    # a segment hugging the border is dropped by the 10 px border filter but present in the raw list
    img = np.full((1, 120, 200), 40, np.uint8)
    img[0, :, :4] = 220
    raw, _ = ctx.lsd_detect_batch(img, filter=False)
    flt, _ = ctx.lsd_detect_batch(img, filter=True)
    assert len(raw[0]) >= 1 and len(flt[0]) == 0
    _check(ctx, oracle, img, filter=False)
    # capacity: CSB_ERR_CAPACITY when a frame yields more segments than max_lines
    frames = np.ascontiguousarray(noise[:1])
    n_ref = len(oracle.lsd_detect(frames[0], mode=0))
    if n_ref > 1:
        with pytest.raises(csb.CsbError) as e:
            ctx.lsd_detect_batch(frames, filter=False, max_lines=1)
        assert e.value.code == csb.CSB_ERR_CAPACITY


def test_class_mirror_and_rerun(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    frames = synth.make_lsd_frames(3, 640, 480, seed=21)
    det = csb.line_lbd_detect(ctx)
    det.use_LSD = True
    det.line_length_thres = 15
    a = det.detect_filter_lines(frames)
    b = det.detect_filter_lines(frames[1])
    assert np.array_equal(a[1], b)
    ref = oracle.lsd_detect(frames[1])
    assert b.shape == ref.shape and np.abs(b - ref).max() <= 1e-4
    ctx.lsd_upload(frames)
    ctx.lsd_run(); l1, _ = ctx.lsd_download()
    ctx.lsd_run(); l2, _ = ctx.lsd_download()  # the used map is rebuilt on every run
    assert all(np.array_equal(x, y) for x, y in zip(l1, l2)) and all(np.array_equal(x, y) for x, y in zip(l1, a))


@pytest.mark.parametrize("link_deg", [3, 8, 20])
def test_unit_partition_does_not_change_the_result(ctx, csb, oracle, link_deg):
    """The work partition (units = pixels linked by similar level-line angles) is validated at run time: units whose regions find a pixel
    of another unit aligned are merged and redone.  A tiny link angle fragments every edge and forces many merge rounds (up to the
    whole-frame fallback); the segments must not change."""
    from cube_slam_wu_b200 import synth
    frames = np.concatenate([synth.make_lsd_frames(2, 640, 480, seed=31), synth.make_lsd_frames(2, 640, 480, seed=32, texture=1.0, noise_sigma=6.0)])
    lines, st, worst = _check(ctx, oracle, frames, unit_link_deg=link_deg)
    assert st.n_merge_rounds > 0 and st.n_unit_conflicts > 0
    print("link %d deg: %d merge rounds, %d conflicts, max diff %g" % (link_deg, st.n_merge_rounds, st.n_unit_conflicts, worst))


def test_reference_golden_vector_on_gpu(ctx, csb, oracle):
    """The reference's own LSD output for its bundled demo image (tests/golden/lsd_demo.npz, see test_lsd_oracle.py): the GPU path
    reproduces the oracle bit for bit and hence the file (271 segments, same order, 6 printed digits but for one segment at 7e-4 px)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lsd_demo.npz"))
    frames = np.ascontiguousarray(d["gray"][None])
    lines, st, worst = _check(ctx, oracle, frames)
    ref = d["ref_lines"]
    assert lines[0].shape == ref.shape
    assert np.abs(lines[0].astype(np.float64) - ref).max() < 2e-3


def test_second_reference_golden_vector_on_gpu(ctx, csb, oracle):
    """line_lbd/data/saved_edges.txt (295 segments of line_lbd/data/407.jpg, 640 x 479; tests/golden/lsd_407.npz) through the GPU path, and the
    descriptors of those segments against the descriptor oracle (detect_descrip_lines on a real image)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lsd_407.npz"))
    frames = np.ascontiguousarray(d["gray"][None])
    lines, st, worst = _check(ctx, oracle, frames)
    ref = d["ref_lines"]
    assert lines[0].shape == ref.shape
    assert np.abs(lines[0].astype(np.float64) - ref).max() < 2e-3
    det = csb.line_lbd_detect(ctx)
    det.line_length_thres = 15.0
    l2, desc = det.detect_descrip_lines(frames[0])
    assert np.array_equal(l2, lines[0])
    _, r32, _ = oracle.lbd_describe(frames[0], l2)
    assert np.array_equal(desc, r32)


def test_full_hd_frame_and_tiny_frames(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    big = synth.make_lsd_frames(1, 1920, 1080, seed=51, n_polys=30, n_lines=60)
    lines, st, worst = _check(ctx, oracle, big)   # used bitmap = 166 KB of shared memory
    assert st.n_lines > 200
    for (w, h) in ((8, 8), (9, 33), (40, 8)):
        tiny = np.random.default_rng(w * h).integers(0, 256, (2, h, w)).astype(np.uint8)
        _check(ctx, oracle, tiny, filter=False)
    with pytest.raises(csb.CsbError):
        ctx.lsd_detect_batch(np.zeros((1, 4, 4), np.uint8))   # below the 8 x 8 minimum
