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
