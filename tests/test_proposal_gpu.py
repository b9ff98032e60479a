"""GPU parity tests of the proposal half: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): proposal ranking indices bit-exact, pose / score floats within 1e-4.  Against the
oracle built with leak_cam_state=0 (boxes independent, the product's documented semantics) the implementation is
held to a much tighter bar: valid-hypothesis lists, corners, distance errors, kept-index lists and ranking indices
bit-exact; everything that passes through atan2 within 1e-9.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _run(ctx, csb, batch, params, leak=0, tol=H.TOL_TIGHT, exact=True, check_tasks=True):
    ora = H.run_oracle(batch, params, leak=leak)
    frames, boxes, lines, tasks, n_tasks, maps, n_map = H.gpu_inputs(csb, batch, params)
    cub, ncub, st = ctx.detect_batch(frames, boxes, lines, tasks, n_tasks, maps, n_map, params)
    s = H.compare_with_oracle(ctx, csb, batch, params, cub, ncub, ora, tol=tol, exact_dist=exact, check_tasks=check_tasks)
    assert st.n_scored == sum(r.n_scored for r in ora)
    assert st.n_enumerated == sum(r.n_enum for r in ora)
    return s, st, ora


def test_demo_no_sampling(ctx, csb):
    """Config #1: bundled single-image demo, flags of detect_3d_cuboid/src/main.cpp:62-68 -> 320 enumerated, 111 valid."""
    p = csb.DetectParams.default(whether_sample_cam_roll_pitch=0)
    s, st, ora = _run(ctx, csb, H.demo_batch(), p)
    assert st.n_enumerated == 320 and st.n_scored == 111


def test_demo_roll_pitch_sampling(ctx, csb):
    p = csb.DetectParams.default()
    s, st, ora = _run(ctx, csb, H.demo_batch(), p)
    assert st.n_enumerated == 6400 and st.n_scored == 1799
    assert st.n_tasks_smem_map == 0  # 351x241 map does not fit shared memory: exercises the global-gather path


@pytest.mark.parametrize("seed", [20260925, 7, 123])
def test_kitti_frames(ctx, csb, seed):
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(3, seed=seed)
    s, st, _ = _run(ctx, csb, batch, csb.DetectParams.default())
    assert s["n_scored"] > 1000
    assert st.n_tasks_smem_map > 0


def test_kitti_leaky_reference_semantics(ctx, csb):
    """Against the literal reference behaviour (cam_pose leaks across boxes): indices identical, floats within 1e-4."""
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(2, seed=99)
    _run(ctx, csb, batch, csb.DetectParams.default(), leak=1, tol=H.TOL_NORTH_STAR, exact=False)


def test_height_sampling_topk_skew(ctx, csb):
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(2, seed=5)
    p = csb.DetectParams.default(whether_sample_bbox_height=1, max_cuboid_num=5, nominal_skew_ratio=2.0)
    s, st, _ = _run(ctx, csb, batch, p)
    assert s["n_tasks"] > 16  # several height samples per box


@pytest.mark.parametrize("c1,c2", [(1, 0), (0, 1)])
def test_single_configuration(ctx, csb, c1, c2):
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(1, seed=11)
    _run(ctx, csb, batch, csb.DetectParams.default(consider_config_1=c1, consider_config_2=c2))


def test_no_lines_angle_saturates(ctx, csb):
    """No line segments: every VP lacks support, all angle errors tie at pi -> the angle filter is dropped and the kept list
    keeps std::partial_sort's order (object_3d_util.cpp:783-786)."""
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(2, seed=3)
    batch["lines"] = np.zeros((0, 4))
    batch["line_ranges"] = [(0, 0)] * 2
    _run(ctx, csb, batch, csb.DetectParams.default())


def test_sparse_lines(ctx, csb):
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(2, seed=4, lines_per_box=3, bg_lines=2)
    _run(ctx, csb, batch, csb.DetectParams.default())


@pytest.mark.parametrize("with_lines", [True, False])
def test_constant_distance_map_massive_ties(ctx, csb, with_lines):
    """All distance errors collide per configuration: membership of the kept 2/3 is decided by libstdc++'s heap order,
    which the GPU path must reproduce exactly."""
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(2, seed=8)
    batch["map_fn"] = lambda f, l, t, w, h: np.full((h, w), 2.5, np.float32)
    if not with_lines:
        batch["lines"] = np.zeros((0, 4)); batch["line_ranges"] = [(0, 0)] * 2
    _run(ctx, csb, batch, csb.DetectParams.default(max_cuboid_num=3))


def test_quantised_distance_map_ties(ctx, csb):
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(2, seed=21)
    imgs = batch["images"]
    batch["map_fn"] = lambda f, l, t, w, h: np.floor(synth.dist_map_for_roi(imgs[f], l, t, w, h) / 4.0).astype(np.float32)
    _run(ctx, csb, batch, csb.DetectParams.default())


def test_edge_boxes(ctx, csb):
    """Tiny boxes (no top samples -> empty ObjectSet), boxes touching the image border (ROI clamped, samples on the ROI bound),
    frames without boxes / lines."""
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(3, seed=13)
    W, Hh = batch["img_w"], batch["img_h"]
    bx = batch["boxes"]
    b0, b1 = batch["box_ranges"][0]
    bx[b0] = [100, 100, 8, 40, 0.5]              # width < 10: top_sample_resolution < 1 -> break
    bx[b0 + 1] = [0, 0, 150, 120, 0.5]           # top-left corner of the image
    bx[b0 + 2] = [W - 201, Hh - 141, 200, 140, 0.5]  # bottom-right: right_x_raw == img_width-1
    bx[b0 + 3] = [300, 5, 25, 30, 0.5]           # small box: few top samples
    # frame 1: no boxes; frame 2: no lines
    lo, hi = batch["box_ranges"][1]
    batch["boxes"] = np.concatenate([bx[:lo], bx[hi:]])
    n_removed = hi - lo
    batch["box_ranges"] = [batch["box_ranges"][0], (lo, lo), (batch["box_ranges"][2][0] - n_removed, batch["box_ranges"][2][1] - n_removed)]
    l2 = batch["line_ranges"][2]
    batch["line_ranges"][2] = (l2[0], l2[0])
    s, st, ora = _run(ctx, csb, batch, csb.DetectParams.default())
    assert len(ora[0].boxes[0]["sorted"]) == 0


def test_config2_batch_64_frames(ctx, csb):
    """BASELINE config #2 at full size: 64 KITTI-shaped frames x 8 boxes.  Box-level parity for all 512 boxes, plus
    size-independent properties: resident re-run is idempotent, results do not depend on batch composition."""
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(64)
    p = csb.DetectParams.default()
    ora = H.run_oracle(batch, p, leak=0)
    frames, boxes, lines, tasks, n_tasks, maps, n_map = H.gpu_inputs(csb, batch, p)
    ctx.detect_upload(frames, boxes, lines, tasks, n_tasks, maps, n_map, p)
    ctx.detect_run()
    cub, ncub, st = ctx.detect_download()
    H.compare_with_oracle(ctx, csb, batch, p, cub, ncub, ora, check_tasks=False)
    assert st.n_scored == sum(r.n_scored for r in ora)
    first = bytes(cub)
    ctx.detect_run(timed=True)
    cub2, ncub2, st2 = ctx.detect_download()
    assert bytes(cub2) == first and np.array_equal(ncub, ncub2)  # idempotent
    assert st2.gpu_ms_score > 0
    # a sub-batch (frames 10..13) gives byte-identical cuboids for its boxes
    sub = dict(batch)
    sub["K"], sub["T"] = batch["K"][10:14], batch["T"][10:14]
    b0 = batch["box_ranges"][10][0]; l0 = batch["line_ranges"][10][0]
    sub["box_ranges"] = [(a - b0, b - b0) for a, b in batch["box_ranges"][10:14]]
    sub["line_ranges"] = [(a - l0, b - l0) for a, b in batch["line_ranges"][10:14]]
    sub["boxes"] = batch["boxes"][b0:batch["box_ranges"][13][1]]
    sub["lines"] = batch["lines"][l0:batch["line_ranges"][13][1]]
    sub["images"] = batch["images"][10:14]
    f2, bx2, ln2, t2, nt2, m2, nm2 = H.gpu_inputs(csb, sub, p)
    cub3, ncub3, _ = ctx.detect_batch(f2, bx2, ln2, t2, nt2, m2, nm2, p)
    for i in range(len(bx2)):
        a, b = cub[b0 + i], cub3[i]
        assert ncub[b0 + i] == ncub3[i]
        if ncub3[i]:
            assert list(a.pos) == list(b.pos) and a.rank_index == b.rank_index and a.normalized_error == b.normalized_error


def test_config2_against_the_literal_reference_variant(ctx, csb):
    """BASELINE config #2 (512 boxes) against the oracle in LITERAL mode -- libm atan2 and the cam_pose state leak, what the reference binary
    does (and what bench.py's reference arm runs).  The product's specified atan2 differs from glibc's by <= 1 ulp, which can flip
    structural near-ties of the 2/3 selection (yaw and yaw + 90 degrees describe the same cuboid): count the boxes whose
    best proposal differs, and hold the others to the north star's 1e-4."""
    from cube_slam_wu_b200 import synth
    batch = synth.make_kitti_batch(64)
    p = csb.DetectParams.default()
    lit = H.run_oracle(batch, p, leak=1, libm=1)
    frames, boxes, lines, tasks, n_tasks, maps, n_map = H.gpu_inputs(csb, batch, p)
    cub, ncub, st = ctx.detect_batch(frames, boxes, lines, tasks, n_tasks, maps, n_map, p)
    n_box = n_diff = 0
    worst = 0.0
    for f, R in enumerate(lit):
        b0, b1 = batch["box_ranges"][f]
        for b in range(b1 - b0):
            ob, gb = R.boxes[b], b0 + b
            n_box += 1
            assert ncub[gb] == len(ob["sorted"])
            if not ncub[gb]:
                continue
            gc, oc = cub[gb], ob["raw"][ob["sorted"][0]]
            if gc.rank_index != ob["sorted"][0]:
                n_diff += 1
                continue
            assert list(gc.box_corners_2d) == list(oc.box_corners_2d)
            for name in ("pos", "scale", "box_corners_3d_world"):
                worst = max(worst, float(np.abs(np.array(getattr(gc, name)) - np.array(getattr(oc, name))).max()))
            for name in ("rotY", "edge_distance_error", "edge_angle_error", "normalized_error", "skew_ratio"):
                worst = max(worst, abs(getattr(gc, name) - getattr(oc, name)))
    print("literal-reference variant: %d of %d boxes with a different top-1 ranking index; the others agree to %.2e" % (n_diff, n_box, worst))
    assert n_box == 512 and worst <= H.TOL_NORTH_STAR
    assert n_diff == 0, "%d of 512 boxes rank differently against the libm / state-leak variant" % n_diff   # measured: none on this workload


def test_api_errors(ctx, csb):
    import ctypes as C
    L = csb.lib()
    assert L.csb_detect_run(None, 0) == csb.CSB_ERR_INVALID
    c2 = csb.Context(0)
    with pytest.raises(csb.CsbError):
        c2.detect_run()  # before upload
    c2.close()


def test_batched_atan2(ctx, csb, oracle):
    """det_atan2_x6 (six edge angles of a proposal evaluated as independent straight-line chains) against the scalar det_atan2 of the
    oracle: identical bits, including groups that contain operands of the special paths (zeros, huge ratios, infinities)."""
    L = oracle.lib()
    rng = np.random.default_rng(7)
    n = 6 * 20000
    y = np.concatenate([rng.normal(0, 50, n // 2), rng.uniform(-1e-3, 1e-3, n // 4), rng.normal(0, 1e5, n // 4)])
    x = np.concatenate([rng.normal(0, 50, n // 2), rng.normal(0, 1, n // 4), rng.uniform(-1, 1, n // 4)])
    # interval boundaries of the argument reduction and special operands
    for i, q in enumerate([0.4375, 0.6875, 1.1875, 2.4375, 2.0 ** -29, 2.0 ** 66, 1.0, 0.0, np.inf, 1e-320]):
        y[6 * i] = q; x[6 * i] = 1.0
        y[6 * i + 1] = np.nextafter(q, 0) if np.isfinite(q) else q; x[6 * i + 1] = -1.0
        y[6 * i + 2] = -q; x[6 * i + 2] = np.nextafter(1.0, 2)
    y[6 * 30] = 1.0; x[6 * 30] = 0.0
    y[6 * 31] = 1e200; x[6 * 31] = -1e-200
    out, n_fallback = ctx.debug_atan2(y, x)
    ref = np.array([L.orc_det_atan2(float(a), float(b)) for a, b in zip(y, x)])
    same = (out == ref) | (np.isnan(out) & np.isnan(ref))
    assert same.all(), "%d of %d differ, e.g. %r" % (int((~same).sum()), n, (y[~same][0], x[~same][0], out[~same][0], ref[~same][0]))
    assert np.array_equal(np.signbit(out), np.signbit(ref))
    assert 5 <= n_fallback < 200   # the special operands take the scalar route, the bulk does not
