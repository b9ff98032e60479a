"""The per-frame pipeline end to end on the GPU (BASELINE config #5's frame stage): gray frames -> csb_lsd_detect_batch (the line table)
-> csb_detect_batch_gray (Canny + distance transform + cuboid proposals) -> best cuboid per 2D box, against the same chain of oracles
(oracle LSD -> cv2-style distance maps -> oracle detect_cuboid).  Because each GPU stage reproduces its oracle's output exactly
(segments bit-identical, maps bit-identical), the chained results obey the same bar as the single stages: identical index lists,
floats within 1e-9."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_lsd_lines_feed_the_proposal_stage(ctx, csb, oracle):
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(3, boxes_per_frame=4, seed=123)
    p = csb.DetectParams.default()
    gray = np.ascontiguousarray(np.stack(batch["images"]).astype(np.uint8))
    # stage 1: line table of every frame (what line_lbd_detect::detect_filter_lines hands to detect_cuboid, main_obj.cpp:596-599)
    lines_gpu, _ = ctx.lsd_detect_batch(gray, line_length_thres=15.0)
    lines_ref = [oracle.lsd_detect(gray[f]) for f in range(len(gray))]
    for a, b in zip(lines_gpu, lines_ref):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert sum(len(a) for a in lines_gpu) > 30
    ranges, off = [], 0
    for a in lines_gpu:
        ranges.append((off, off + len(a)))
        off += len(a)
    batch = dict(batch)
    batch["lines"] = np.concatenate(lines_gpu).astype(np.float64)  # float -> double like the caller (main_obj.cpp:598)
    batch["line_ranges"] = ranges
    # stage 2 + 3: Canny / distance transform / proposals from the gray frames, with the detected lines
    frames = csb.make_frames(batch["K"], batch["T"], batch["img_w"], batch["img_h"], batch["box_ranges"], batch["line_ranges"])
    boxes = np.ascontiguousarray(batch["boxes"], np.float64).reshape(-1, 5)
    tasks, n_tasks, n_map = csb.detect_plan(frames, boxes, p)
    cub, ncub, st = ctx.detect_batch_gray(frames, boxes, batch["lines"], tasks, n_tasks, gray.ravel(), p)
    imgs = batch["images"]
    batch["map_fn"] = lambda f, l, t, w, h: synth.dist_map_for_roi_reference(imgs[f], l, t, w, h)
    ora = H.run_oracle(batch, p, leak=0)
    s = H.compare_with_oracle(ctx, csb, batch, p, cub, ncub, ora)
    assert st.n_scored == s["n_scored"]
    print("pipeline: %d segments -> %d scored proposals, max float diff %g" % (off, st.n_scored, s["max_float_diff"]))


def test_config5_frames_to_graph(ctx, csb, oracle):
    """BASELINE config #5 at a small size (cube_slam_wu_b200.pipeline.run_config5): frames -> proposal kernels -> observation records -> host
    graph assembly (main_obj.cpp:738-803) with noisy camera estimates -> csb_ba_set_graph + linearisation, against the oracle's linearisation
    of the same graph: residuals, chi2, right-hand sides and every Hessian block."""
    from cube_slam_wu_b200 import pipeline
    out = pipeline.run_config5(n_frames_total=192, depth=2, ctx=ctx, keep=True)
    g, lin, lin_a, rec = out["_graph"], out["_lin"], out["_lin_analytic"], out["_records"]
    assert out["frames"] == 192 and out["graph"]["cameras"] == 192 and out["graph"]["edges_odometry"] == 191
    n_valid = int((rec[..., 2] == 1).sum())
    assert out["graph"]["edges_cuboid"] == n_valid and n_valid > 192      # most of the 8 boxes of a frame yield a cuboid
    # every box of the 64 distinct frames is a landmark, observed once per pass over the frames
    assert out["graph"]["landmarks"] == n_valid // 3 and g["cubes10"].shape == (512, 10)
    assert np.array_equal(rec[0, 0, :, 2:], rec[0, 1, :, 2:]) and np.array_equal(rec[0, 0, :, 2:], rec[0, 2, :, 2:])
    E = oracle.ba_edges(ec=g["ec"], ep=None, eo=g["eo"])
    ref = oracle.ba_linearize(g["cams7"], g["cam_fixed"], g["cubes10"], g["cube_fixed"], E)
    # not degenerate: the cameras of the three passes sit at different (perturbed) poses, the re-observations disagree
    assert ref["chi2"] > 1.0 and np.abs(ref["b_cam"]).max() > 1e-2 and np.abs(ref["b_cube"]).max() > 1e-2
    assert abs(out["chi2"] - ref["chi2"]) <= 1e-9 * ref["chi2"]
    rel = lambda a, b: float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))
    for k in ("ec_err", "eo_err"):
        assert rel(lin[k], ref[k]) <= 1e-9, k
    # (a) the reference's definition, delta = 1e-9 central differences, on both sides.  The camera and odometry blocks are well conditioned
    # (1e-4, the north-star bar).  The cuboid-side blocks divide the round-off of a residual that multiplies 60-100 m translations by 1e-9:
    # perturbing the cuboid estimates by ONE ulp moves the oracle's own H_cube / ec_Hij by 6e-4 / 1.4e-3 of the block scale, so two correct
    # implementations cannot agree better than that; (b) below is the sharp check of those blocks.
    for k, tol in (("H_cam", 1e-4), ("eo_Hij", 1e-4), ("b_cam", 1e-4), ("H_cube", 2e-2), ("ec_Hij", 2e-2), ("b_cube", 2e-2)):
        assert rel(lin[k], ref[k]) <= tol, "%s differs by %g" % (k, rel(lin[k], ref[k]))
    # (b) closed-form Jacobians on the device against central differences with delta = 1e-5 in the oracle (round-off 1e-11 instead of 1e-7,
    # truncation O(delta^2)): the camera / cuboid blocks and both right-hand sides to 1e-6 of the block scale (north star: 1e-4).  The
    # odometry blocks to 5e-5: SE3Quat::log switches to omega = deltaR / 2, V^-1 = I - W/2 + W^2/12 below 4.5 mrad (se3quat.h:238-247); the
    # derivative of that branch differs from the exact logarithm's by O(theta^2) <= 2e-5, which the central differences see and the
    # closed form (derivative of the exact logarithm) does not -- these residuals (pose noise of 5 mrad) sit right at the switch.
    oracle.ba_set_delta(1e-5)
    try:
        acc = oracle.ba_linearize(g["cams7"], g["cam_fixed"], g["cubes10"], g["cube_fixed"], E)
    finally:
        oracle.ba_set_delta(1e-9)
    for k, tol in (("H_cam", 1e-6), ("b_cam", 1e-6), ("H_cube", 1e-6), ("b_cube", 1e-6), ("ec_Hij", 1e-6), ("eo_Hij", 5e-5)):
        assert rel(lin_a[k], acc[k]) <= tol, "analytic %s differs by %g" % (k, rel(lin_a[k], acc[k]))
    assert out["edges_per_s"]["numeric"] > 0 and out["allgather_bytes_per_rank"] == 3 * 512 * 128
