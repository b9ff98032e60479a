"""GPU parity of the LBD line descriptor (SURVEY.md 8 "next" row f-2, BASELINE config #3): csb_lbd_* vs the CPU oracle
(oracle/oracle_lbd.cpp), through the C ABI.

Bar: bit-exact.  The int16 gradient images are integer work; the 72-float descriptors are float arithmetic in the reference's order with
no contraction on either side and specified atan2 / sin / cos, so they are compared bit for bit (NaN = NaN: a flat support region divides
0 by 0 in the reference too); the 32-byte binary descriptors must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32


def _lines(rng, w, h, n):
    p = rng.uniform([0, 0, 0, 0], [w - 1, h - 1, w - 1, h - 1], (n, 4)).astype(f32)
    if n >= 4:
        p[0] = [-5, 10, w + 20, 12.5]       # clamped at both ends
        p[1] = [30.5, 40.5, 30.5, 40.5]     # zero length
        p[2] = [3, 3, 3, h - 2]             # vertical; the support region leaves the frame
        p[3] = [w - 2, 20, 10, 20]          # horizontal, right to left
    return p


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32) if a.dtype == f32 else a, b.view(np.uint32) if b.dtype == f32 else b)


def _same_f(a, b):
    """bit-identical floats, any NaN equal to any NaN (x86 and CUDA produce different NaN payloads for sqrt(-x) / 0 * inf)"""
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a).view(np.uint32), np.nan_to_num(b).view(np.uint32))


def test_gradients_bit_exact(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    for (w, h) in ((640, 480), (200, 150), (132, 33), (333, 251), (64, 16), (9, 8)):  # width % 4 == 0: the packed kernel
        frames = synth.make_lsd_frames(2, w, h, seed=21, texture=1.0, noise_sigma=4.0)
        ctx.lbd_upload(frames, [np.zeros((0, 4), f32)] * 2)
        ctx.lbd_run()
        for f in range(2):
            dx, dy = ctx.lbd_debug_gradients(f, (h, w))
            _, rdx, rdy = oracle.lbd_gradients(frames[f])
            assert np.array_equal(dx, rdx) and np.array_equal(dy, rdy), "%dx%d frame %d" % (w, h, f)
        out = ctx.lbd_download()
        assert out["stats"].n_lines == 0 and all(len(d) == 0 for d in out["desc"])


def test_descriptors_bit_exact_ragged_batch(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    w, h = 640, 480
    frames = synth.make_lsd_frames(5, w, h, seed=22, texture=1.0, noise_sigma=4.0)
    rng = np.random.default_rng(22)
    counts = [60, 0, 7, 133, 1]
    lines = [_lines(rng, w, h, n) for n in counts]
    d32, d72, st = ctx.lbd_describe_batch(frames, lines, want_float=True)
    assert st.n_lines == sum(counts) and st.n_kernel_launches == 3
    samples = 0
    for f in range(5):
        r72, r32, kl = oracle.lbd_describe(frames[f], lines[f])
        assert _same_f(d72[f], r72), "frame %d: float descriptors differ (max %g)" % (f, np.nanmax(np.abs(d72[f] - r72)) if len(r72) else 0)
        assert np.array_equal(d32[f], r32), "frame %d: binary descriptors differ" % f
        samples += int(63 * kl[:, 1].sum())
    assert st.n_samples == samples
    # binary only (no float buffer), resident variant, key-line fields
    ctx.lbd_upload(frames, lines)
    ctx.lbd_run(timed=True)
    out = ctx.lbd_download(keylines=True)
    for f in range(5):
        _, r32, kl = oracle.lbd_describe(frames[f], lines[f])
        assert np.array_equal(out["desc"][f], r32)
        assert _same(np.ascontiguousarray(out["keylines"][f][:, :3]), kl)
    assert out["stats"].gpu_ms_grad > 0 and out["stats"].gpu_ms_describe > 0


def test_descriptors_small_and_kitti_frames(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    for (w, h, seed) in ((1242, 375, 23), (97, 61, 24)):
        frames = synth.make_lsd_frames(2, w, h, seed=seed, texture=1.0, noise_sigma=3.0)
        rng = np.random.default_rng(seed)
        lines = [_lines(rng, w, h, 40), _lines(rng, w, h, 3)]
        d32, d72, st = ctx.lbd_describe_batch(frames, lines, want_float=True)
        for f in range(2):
            r72, r32, _ = oracle.lbd_describe(frames[f], lines[f])
            assert _same_f(d72[f], r72) and np.array_equal(d32[f], r32)


def test_detect_descrip_lines_chain(ctx, csb, oracle):
    """line_lbd_detect::detect_descrip_lines: LSD segments stay on the device and are described in place."""
    from cube_slam_wu_b200 import synth
    frames = synth.make_lsd_frames(3, 640, 480, seed=25)
    det = csb.line_lbd_detect(ctx)
    det.line_length_thres = 15.0
    lines, desc = det.detect_descrip_lines(frames)
    total = 0
    for f in range(3):
        ref_lines = oracle.lsd_detect(frames[f], length_thres=15.0)
        assert lines[f].shape == ref_lines.shape and (len(ref_lines) == 0 or np.abs(lines[f] - ref_lines).max() <= 1e-4)
        _, r32, _ = oracle.lbd_describe(frames[f], lines[f])
        assert np.array_equal(desc[f], r32), "frame %d: %d of %d descriptors differ" % (f, int((desc[f] != r32).any(axis=1).sum()), len(r32))
        total += len(r32)
    assert total > 50
    one_l, one_d = det.detect_descrip_lines(frames[1])
    assert np.array_equal(one_l, lines[1]) and np.array_equal(one_d, desc[1])
    # float descriptors of the chained run, and the capacity error of the download
    ctx.lsd_upload(frames, 15.0, True, 4096)
    ctx.lsd_run()
    ctx.lbd_run_on_lsd(want_float=True)
    out = ctx.lbd_download(want_float=True)
    for f in range(3):
        r72, _, _ = oracle.lbd_describe(frames[f], lines[f])
        assert _same_f(out["desc_float"][f], r72)
    import ctypes as C
    cnt = np.zeros(3, np.int32)
    rc = csb.lib().csb_lbd_download(ctx._h, None, None, None, cnt.ctypes.data_as(C.c_void_p), C.c_int64(total - 1), None)
    assert rc == csb.CSB_ERR_CAPACITY and int(cnt.sum()) == total


def test_call_order(ctx, csb):
    import ctypes as C
    c2 = csb.Context(0)
    try:
        assert csb.lib().csb_lbd_run(c2._h, 0) == csb.CSB_ERR_STATE
        assert csb.lib().csb_lbd_run_on_lsd(c2._h, 0, 0) == csb.CSB_ERR_STATE
        assert csb.lib().csb_lbd_download(c2._h, None, None, None, None, C.c_int64(0), None) == csb.CSB_ERR_STATE
    finally:
        c2.close()
