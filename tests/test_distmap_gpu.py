"""GPU parity of the Canny + distance-transform stage (SURVEY.md 8 "next" row f-1): csb_detect_batch_gray vs python cv2 4.13.

Parity target = what the reference's C++ computes (box_proposal_detail.cpp:320-327) with OpenCV's open-source algorithms:
Sobel on the parent image (a cv::Mat ROI view sees its real neighbours), cv::Canny(dx, dy, 80, 200) on the ROI, fixed-point 3x3
chamfer distance transform (cv2 with IPP disabled).  Bar: Canny map and float32 distance map bit-exact; the scored path fed by the
GPU maps then matches the oracle fed by the cv2 maps exactly as in test_proposal_gpu.py."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["256", "1024"])
def distmap_cta(request, monkeypatch):
    """k_distmap has two CTA sizes: 256 threads for a batch (four ROIs per SM) and 1024 when a call holds at most one ROI per SM (a single
    frame; latency).  launch_distmaps picks by task count; every case here runs through both (CSB_DISTMAP_CTA forces one)."""
    monkeypatch.setenv("CSB_DISTMAP_CTA", request.param)
    return request.param


def _gray(batch):
    return np.ascontiguousarray(np.concatenate([im.ravel() for im in batch["images"]]).astype(np.uint8))


def _run_gray(ctx, csb, batch, params):
    from cube_slam_wu_b200 import synth
    frames = csb.make_frames(batch["K"], batch["T"], batch["img_w"], batch["img_h"], batch["box_ranges"], batch["line_ranges"])
    boxes = np.ascontiguousarray(batch["boxes"], np.float64).reshape(-1, 5)
    lines = np.ascontiguousarray(batch["lines"], np.float64).reshape(-1, 4)
    tasks, n_tasks, n_map = csb.detect_plan(frames, boxes, params)
    cub, ncub, st = ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, _gray(batch), params)
    worst = 0.0
    for i in range(n_tasks):
        t = tasks[i]
        dm, ed = ctx.debug_map(i, t, edges=True)
        ref_dm, ref_ed = synth.dist_map_for_roi_reference(batch["images"][t.frame_id], t.roi_left, t.roi_top, t.roi_width, t.roi_height, return_edges=True)
        assert np.array_equal((ed == 2), ref_ed > 0), "task %d: Canny map differs in %d px" % (i, int(((ed == 2) != (ref_ed > 0)).sum()))
        assert np.array_equal(dm, ref_dm), "task %d: distance map differs, max %g" % (i, float(np.abs(dm - ref_dm).max()))
        worst = max(worst, float(dm.max()))
    return cub, ncub, st, tasks, n_tasks


def test_maps_and_cuboids_kitti(ctx, csb):
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(3, seed=77)
    p = csb.DetectParams.default()
    cub, ncub, st, tasks, n_tasks = _run_gray(ctx, csb, batch, p)
    assert st.gpu_ms_distmap > 0
    # scored path on top of the GPU maps == oracle on top of the cv2 maps
    imgs = batch["images"]
    batch["map_fn"] = lambda f, l, t, w, h: synth.dist_map_for_roi_reference(imgs[f], l, t, w, h)
    ora = H.run_oracle(batch, p, leak=0)
    H.compare_with_oracle(ctx, csb, batch, p, cub, ncub, ora)


def test_noise_frames_and_image_borders(ctx, csb):
    """Dense edges (uniform noise) and ROIs clamped at the image border (BORDER_REPLICATE there, real neighbours elsewhere)."""
    from cube_slam_wu_b200 import synth
    rng = np.random.default_rng(5)
    batch = synth.make_kitti_batch(2, seed=78)
    W, Hh = batch["img_w"], batch["img_h"]
    batch["images"] = [rng.integers(0, 256, (Hh, W)).astype(np.uint8), (rng.integers(0, 2, (Hh, W)) * 255).astype(np.uint8)]
    b0 = batch["box_ranges"][0][0]
    batch["boxes"][b0] = [0, 0, 150, 120, 0.5]
    batch["boxes"][b0 + 1] = [W - 201, Hh - 141, 200, 140, 0.5]
    batch["boxes"][b0 + 2] = [3, Hh - 90, 130, 85, 0.5]
    _run_gray(ctx, csb, batch, csb.DetectParams.default())


def test_wide_and_tiny_rois(ctx, csb):
    """ROI widths that take every column-per-thread class of the sweeps (<= 256, <= 512, <= 768, wider), the shared-memory row pass of ROIs
    wider than 416 columns, and ROIs of a few pixels; textured frames so that every stage has work."""
    from cube_slam_wu_b200 import synth
    rng = np.random.default_rng(11)
    batch = synth.make_kitti_batch(2, boxes_per_frame=8, seed=80)
    W, Hh = batch["img_w"], batch["img_h"]
    yy, xx = np.mgrid[0:Hh, 0:W]
    tex = (127 + 90 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + rng.normal(0, 12, (Hh, W))).clip(0, 255).astype(np.uint8)
    batch["images"] = [tex, np.ascontiguousarray(tex[::-1, ::-1])]
    b0 = batch["box_ranges"][0][0]
    for k, bx in enumerate([[20, 20, 1190, 320, 0.5], [30, 40, 700, 200, 0.5], [100, 50, 450, 250, 0.5], [600, 200, 6, 5, 0.5], [5, 5, 3, 40, 0.5],
                            [300, 100, 257, 9, 0.5], [640, 60, 513, 120, 0.5], [200, 150, 769, 100, 0.5]]):
        batch["boxes"][b0 + k] = bx
    _run_gray(ctx, csb, batch, csb.DetectParams.default(whether_sample_bbox_height=1))


def test_blank_and_single_edge_rois(ctx, csb):
    """No edge pixel at all in a ROI (distance saturates at DIST_MAX) and a single vertical step edge."""
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(2, seed=79)
    W, Hh = batch["img_w"], batch["img_h"]
    blank = np.full((Hh, W), 90, np.uint8)
    step = np.full((Hh, W), 20, np.uint8); step[:, W // 2:] = 220
    batch["images"] = [blank, step]
    _run_gray(ctx, csb, batch, csb.DetectParams.default(whether_sample_bbox_height=1))


def test_pinned_frames_are_fetched_by_roi_segments(ctx, csb):
    """csb_detect_upload_gray with a PINNED gray buffer (CSB_OPT_GRAY_GATHER, default on): a kernel fetches only the 32-byte segments the
    ROIs (+ Sobel halo) touch from the caller's memory.  Same maps, same cuboids as the whole-frame copy; fewer bytes over PCIe."""
    import torch
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(5, boxes_per_frame=3, seed=91)
    p = csb.DetectParams.default()
    frames = csb.make_frames(batch["K"], batch["T"], batch["img_w"], batch["img_h"], batch["box_ranges"], batch["line_ranges"])
    boxes = np.ascontiguousarray(batch["boxes"], np.float64).reshape(-1, 5)
    lines = np.ascontiguousarray(batch["lines"], np.float64).reshape(-1, 4)
    tasks, n_tasks, n_map = csb.detect_plan(frames, boxes, p)
    g = _gray(batch)
    # poison the device frame buffer first: whatever the gather does not fetch must not matter
    ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, np.full_like(g, 255), p)
    tg = torch.from_numpy(g).pin_memory()
    cub1, ncub1, st1 = ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, tg.numpy(), p)
    for i in range(n_tasks):
        t = tasks[i]
        dm, ed = ctx.debug_map(i, t, edges=True)
        ref_dm, ref_ed = synth.dist_map_for_roi_reference(batch["images"][t.frame_id], t.roi_left, t.roi_top, t.roi_width, t.roi_height, return_edges=True)
        assert np.array_equal((ed == 2), ref_ed > 0) and np.array_equal(dm, ref_dm), "task %d" % i
    ctx.set_option(csb.CSB_OPT_GRAY_GATHER, 0)
    try:
        cub0, ncub0, st0 = ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, tg.numpy(), p)
    finally:
        ctx.set_option(csb.CSB_OPT_GRAY_GATHER, 1)
    assert np.array_equal(ncub0, ncub1) and bytes(cub0) == bytes(cub1)
    assert st0.h2d_bytes - st1.h2d_bytes > 0.3 * g.size and st1.h2d_bytes > 0.1 * g.size, (st0.h2d_bytes, st1.h2d_bytes, g.size)
