"""Shared helpers of the parity tests: run the oracle frame by frame, run the CUDA path on the whole batch, compare."""
import os

import numpy as np

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north-star tolerance for pose / score floats; the index lists must match exactly
TOL_NORTH_STAR = 1e-4
# what the implementation actually achieves against the no-leak oracle (only atan2 differs by <= 2 ulp)
TOL_TIGHT = 1e-9


def demo_batch():
    d = np.load(os.path.join(GOLDEN, "demo_case.npz"))
    img_w, img_h = int(d["img_w"]), int(d["img_h"])
    return dict(K=d["K"][None], T=d["T"][None], boxes=d["boxes"], lines=d["lines"], box_ranges=[(0, 1)], line_ranges=[(0, len(d["lines"]))],
                images=None, img_w=img_w, img_h=img_h, demo_map=d["dist_map"], demo_roi=d["roi"])


def oracle_params(p, leak, libm=0):
    return O.default_params(consider_config_1=p.consider_config_1, consider_config_2=p.consider_config_2,
                            whether_sample_cam_roll_pitch=p.whether_sample_cam_roll_pitch, whether_sample_bbox_height=p.whether_sample_bbox_height,
                            max_cuboid_num=p.max_cuboid_num, leak_cam_state=leak, libm_atan2=libm, nominal_skew_ratio=p.nominal_skew_ratio, max_cut_skew=p.max_cut_skew)


def maps_for(batch, f, rois, offsets, total):
    """float32 buffer with one distance map per ROI (left, top, w, h) at the given float offsets."""
    from cube_slam_wu_b200 import synth
    buf = np.zeros(int(total) + 16, np.float32)
    for (l, t, w, h), o in zip(rois, offsets):
        if batch.get("demo_map") is not None:
            assert tuple(batch["demo_roi"]) == (l, t, w, h)
            dm = batch["demo_map"]
        elif batch.get("map_fn") is not None:
            dm = batch["map_fn"](f, l, t, w, h)
        else:
            dm = synth.dist_map_for_roi(batch["images"][f], l, t, w, h)
        buf[o:o + w * h] = dm.ravel()
    return buf


def run_oracle(batch, params, leak=0, libm=0):
    """One oracle detect_cuboid() per frame.  Returns list of FrameResult."""
    res = []
    sample_h = bool(params.whether_sample_bbox_height)
    for f in range(len(batch["K"])):
        b0, b1 = batch["box_ranges"][f]
        l0, l1 = batch["line_ranges"][f]
        boxes = batch["boxes"][b0:b1]
        tasks = O.plan(boxes, batch["img_w"], batch["img_h"], sample_h)
        rois = [(t.left, t.top, t.width, t.height) for t in tasks]
        offs = [t.map_offset for t in tasks]
        total = sum(w * h for (_, _, w, h) in rois)
        maps = maps_for(batch, f, rois, offs, total)
        res.append(O.detect_frame(batch["K"][f], batch["T"][f], batch["img_w"], batch["img_h"], boxes, batch["lines"][l0:l1], maps, oracle_params(params, leak, libm)))
    return res


def gpu_inputs(csb, batch, params):
    frames = csb.make_frames(batch["K"], batch["T"], batch["img_w"], batch["img_h"], batch["box_ranges"], batch["line_ranges"])
    boxes = np.ascontiguousarray(batch["boxes"], np.float64).reshape(-1, 5)
    lines = np.ascontiguousarray(batch["lines"], np.float64).reshape(-1, 4)
    tasks, n_tasks, n_map = csb.detect_plan(frames, boxes, params)
    buf = np.zeros(int(n_map) + 16, np.float32)
    for i in range(n_tasks):
        t = tasks[i]
        m = maps_for(batch, t.frame_id, [(t.roi_left, t.roi_top, t.roi_width, t.roi_height)], [0], t.roi_width * t.roi_height)
        buf[t.map_offset:t.map_offset + t.roi_width * t.roi_height] = m[:t.roi_width * t.roi_height]
    return frames, boxes, lines, tasks, n_tasks, buf, n_map


def compare_with_oracle(ctx, csb, batch, params, cub, ncub, oracle_res, tol=TOL_TIGHT, exact_dist=True, check_tasks=True):
    """Asserts index lists identical and floats within tol.  Returns summary dict."""
    kmax = params.max_cuboid_num
    t_global = 0
    n_scored = 0
    max_f = 0.0
    for f, R in enumerate(oracle_res):
        b0, b1 = batch["box_ranges"][f]
        task_base = t_global
        for ti, ot in enumerate(R.tasks):
            if check_tasks:
                g = ctx.debug_task(t_global, 0)
                assert g["n_valid"] == ot["n_valid"], "frame %d task %d: n_valid %d != oracle %d" % (f, ti, g["n_valid"], ot["n_valid"])
                assert g["n_merged"] == len(ot["merged"]), "frame %d task %d: merged lines %d != %d" % (f, ti, g["n_merged"], len(ot["merged"]))
                assert np.array_equal(g["merged"], ot["merged"]), "merged line table differs"
                assert np.array_equal(g["hyp_id"], ot["hyp_id"]), "frame %d task %d: valid hypothesis ids differ" % (f, ti)
                if ot["n_valid"]:
                    if exact_dist:
                        assert np.array_equal(g["corners"], ot["corners"]), "corners differ: max %g" % np.abs(g["corners"] - ot["corners"]).max()
                        assert np.array_equal(g["dist"], ot["rows"][:, 4]), "distance errors differ: max %g" % np.abs(g["dist"] - ot["rows"][:, 4]).max()
                    else:
                        assert np.abs(g["corners"] - ot["corners"]).max() <= tol
                        assert np.abs(g["dist"] - ot["rows"][:, 4]).max() <= tol
                    da = np.abs(g["angle"] - ot["rows"][:, 5]).max()
                    assert da <= tol, "angle errors differ by %g" % da
                    max_f = max(max_f, da)
                assert np.array_equal(g["keep"], ot["keep"]), "frame %d task %d: kept proposal indices differ (gpu %d, oracle %d)" % (f, ti, len(g["keep"]), len(ot["keep"]))
                if len(ot["keep"]):
                    dn = np.abs(g["norm_score"] - ot["norm_score"])
                    dn = dn[np.isfinite(dn)]
                    if len(dn):
                        assert dn.max() <= tol, "normalized scores differ by %g" % dn.max()
                        max_f = max(max_f, dn.max())
            n_scored += ot["n_valid"]
            t_global += 1
        for b in range(b1 - b0):
            ob = R.boxes[b]
            gb = b0 + b
            assert ncub[gb] == len(ob["sorted"]), "frame %d box %d: %d cuboids vs oracle %d" % (f, b, ncub[gb], len(ob["sorted"]))
            for r in range(ncub[gb]):
                gc = cub[gb * kmax + r]
                oc = ob["raw"][ob["sorted"][r]]
                assert gc.rank_index == ob["sorted"][r], "frame %d box %d: ranking index %d != oracle %d" % (f, b, gc.rank_index, ob["sorted"][r])
                assert gc.raw_cube_ind == oc.raw_cube_ind and gc.task_id == task_base + oc.task_id
                assert list(gc.box_corners_2d) == list(oc.box_corners_2d)
                assert list(gc.box_config_type) == list(oc.box_config_type)
                for name in ("pos", "scale", "box_corners_3d_world", "rect_detect_2d"):
                    d = np.abs(np.array(getattr(gc, name)) - np.array(getattr(oc, name))).max()
                    assert d <= tol, "%s differs by %g" % (name, d)
                    max_f = max(max_f, d)
                for name in ("rotY", "edge_distance_error", "edge_angle_error", "normalized_error", "skew_ratio", "down_expand_height", "camera_roll_delta", "camera_pitch_delta"):
                    a, o = getattr(gc, name), getattr(oc, name)
                    if np.isfinite(o) or np.isfinite(a):
                        assert abs(a - o) <= tol, "%s: %r vs %r" % (name, a, o)
                        max_f = max(max_f, abs(a - o))
    return dict(n_scored=n_scored, max_float_diff=max_f, n_tasks=t_global)
