"""CPU tests of the product's host side: the shared library loads without a GPU, exports every symbol the header declares,
its host planner agrees with the oracle's restatement of box_proposal_detail.cpp:143-256, and computing entry points fail
loudly (no CPU fallback) when CUDA is unavailable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_all_declared_symbols(csb):
    L = csb.lib()
    hdr = open(os.path.join(ROOT, "include", "cubeslam_b200.h")).read()
    names = sorted(set(re.findall(r"\b(csb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    assert b"sm_100a" in L.csb_version()


def test_struct_layouts(csb):
    assert C.sizeof(csb.Task) == 48 and C.sizeof(csb.Frame) == 9 * 8 + 16 * 8 + 6 * 4
    assert C.sizeof(csb.DetectParams) == 40
    assert C.sizeof(csb.Cuboid) == 8 * (3 + 3 + 1 + 2 + 24 + 4 + 7) + 4 * 20


@pytest.mark.parametrize("sample_h", [0, 1])
def test_plan_matches_oracle(csb, sample_h):
    rng = np.random.default_rng(3)
    W, Hh = 1242, 375
    boxes = []
    for _ in range(200):
        w = int(rng.uniform(5, 400)); h = int(rng.uniform(10, 300))
        w = min(w, W - 2); h = min(h, Hh - 2)
        x = int(rng.uniform(0, W - w - 1)); y = int(rng.uniform(0, Hh - h - 1))
        boxes.append([x + rng.uniform(0, 0.9), y + rng.uniform(0, 0.9), w + rng.uniform(0, 0.9), h, 0.5])
    boxes = np.array(boxes)
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.22], [0, 0, 1.0]])
    T = np.eye(4); T[:3, :3] = [[1, 0, 0], [0, 0, 1], [0, -1, 0]]; T[2, 3] = 1.6
    frames = csb.make_frames([K], [T], W, Hh, [(0, len(boxes))], [(0, 0)])
    p = csb.DetectParams.default(whether_sample_bbox_height=sample_h)
    tasks, n, n_map = csb.detect_plan(frames, boxes, p)
    ot = O.plan(boxes, W, Hh, bool(sample_h))
    assert n == len(ot)
    for i in range(n):
        a, b = tasks[i], ot[i]
        assert (a.box_id, a.hs_id, a.down_expand, a.roi_left, a.roi_top, a.roi_width, a.roi_height, a.n_top) == \
               (b.box_id, b.hs_id, b.down_expand, b.left, b.top, b.width, b.height, b.n_top)
        assert a.map_offset % 4 == 0
    assert any(t.n_top for t in ot)


def test_plan_rejects_bad_ranges(csb):
    K = np.eye(3); T = np.eye(4)
    frames = csb.make_frames([K], [T], 100, 100, [(0, 5)], [(0, 0)])
    with pytest.raises(csb.CsbError):
        csb.detect_plan(frames, np.zeros((2, 5)), csb.DetectParams.default())


def test_plan_empty_ragged_and_misordered_batches(csb):
    """Ragged input on the host side of csb_detect_plan / csb_detect_upload (csrc/host_plan.cpp): an empty batch plans to nothing, frames
    without boxes between frames with boxes keep the frame ids of the others, an ROI is clipped to the image (box_proposal_detail.cpp:262-270
    expands by 10 px and clamps), and frame box ranges that overlap or go backwards are refused (k_rank / k_observe rely on a box's tasks
    being contiguous)."""
    K = np.eye(3); T = np.eye(4)
    p = csb.DetectParams.default()
    tasks, n, n_map = csb.detect_plan(csb.make_frames([], [], 100, 100, [], []), np.zeros((0, 5)), p)
    assert (n, n_map) == (0, 0)
    boxes = np.array([[10, 10, 100, 100, .5], [600, 400, 39, 79, .5]])
    frames = csb.make_frames([K, K, K], [T, T, T], 640, 480, [(0, 1), (1, 1), (1, 2)], [(0, 0)] * 3)
    tasks, n, n_map = csb.detect_plan(frames, boxes, p)
    assert [(t.frame_id, t.box_id, t.roi_left, t.roi_top, t.roi_width, t.roi_height) for t in tasks[:n]] == \
        [(0, 0, 0, 0, 120, 120), (2, 1, 590, 390, 49, 89)]
    assert tasks[1].map_offset >= 120 * 120 and tasks[1].map_offset % 4 == 0 and n_map >= tasks[1].map_offset + 49 * 89
    ot = O.plan(boxes, 640, 480, False)
    assert [(b.left, b.top, b.width, b.height) for b in ot] == [(0, 0, 120, 120), (590, 390, 49, 89)]
    for ranges in ([(0, 2), (1, 2)], [(1, 2), (0, 1)]):
        with pytest.raises(csb.CsbError):
            csb.detect_plan(csb.make_frames([K, K], [T, T], 640, 480, ranges, [(0, 0)] * 2), boxes, p)


def test_no_cpu_fallback(csb):
    """Without a CUDA device csb_create fails and nothing computes.  (On a GPU box this test only checks the error paths.)"""
    import torch
    if torch.cuda.is_available():
        L = csb.lib()
        assert L.csb_detect_run(None, 0) == csb.CSB_ERR_INVALID
        return
    with pytest.raises(csb.CsbError) as e:
        csb.Context(0)
    assert e.value.code == csb.CSB_ERR_CUDA


def test_null_context_is_rejected_everywhere(csb):
    """Entry points added for the line detector and the Jacobian mode: a NULL context is an argument error, not a crash, with or without a GPU."""
    L = csb.lib()
    p = csb.LsdParams(15.0, 1, 16, 0)
    img = np.zeros((1, 16, 16), np.uint8)
    out = np.zeros((1, 16, 4), np.float32); n = np.zeros(1, np.int32)
    assert L.csb_lsd_detect_batch(None, img.ctypes.data_as(C.c_void_p), 1, 16, 16, C.byref(p), out.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), None) == csb.CSB_ERR_INVALID
    assert L.csb_lsd_upload(None, img.ctypes.data_as(C.c_void_p), 1, 16, 16, C.byref(p)) == csb.CSB_ERR_INVALID
    assert L.csb_lsd_run(None, 0) == csb.CSB_ERR_STATE
    assert L.csb_lsd_download(None, None, None, None) == csb.CSB_ERR_STATE
    assert L.csb_ba_set_jacobian_mode(None, 1) == csb.CSB_ERR_INVALID
    assert C.sizeof(csb.LsdParams) == 16 and C.sizeof(csb.LsdStats) == 5 * 8 + 4 * 4 + 2 * 4 + 7 * 8


def test_null_context_lbd(csb):
    """The descriptor entry points: NULL context / call-order errors are reported, not crashes, with or without a GPU."""
    L = csb.lib()
    img = np.zeros((1, 16, 16), np.uint8); ln = np.zeros((1, 4), np.float32); off = np.array([0, 1], np.int32); d = np.zeros((1, 32), np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.csb_lbd_describe_batch(None, vp(img), 1, 16, 16, vp(ln), vp(off), vp(d), None, None) == csb.CSB_ERR_INVALID
    assert L.csb_lbd_upload(None, vp(img), 1, 16, 16, vp(ln), vp(off), 0) == csb.CSB_ERR_INVALID
    assert L.csb_lbd_run(None, 0) == csb.CSB_ERR_STATE
    assert L.csb_lbd_run_on_lsd(None, 0, 0) == csb.CSB_ERR_INVALID
    assert L.csb_lbd_download(None, None, None, None, None, C.c_int64(0), None) == csb.CSB_ERR_STATE
    assert L.csb_lbd_debug_gradients(None, 0, None, None) == csb.CSB_ERR_STATE
    assert C.sizeof(csb.LbdStats) == 4 * 8 + 2 * 4 + 2 * 4
    p = csb.LsdParams(15.0, 1, 16, 0)
    assert L.csb_edlines_detect_batch(None, vp(img), 1, 16, 16, C.byref(p), None, None, None) == csb.CSB_ERR_INVALID
    assert L.csb_edlines_upload(None, vp(img), 1, 16, 16, C.byref(p)) == csb.CSB_ERR_INVALID
    assert L.csb_edlines_run(None, 0) == csb.CSB_ERR_STATE
    assert L.csb_edlines_download(None, None, None, None) == csb.CSB_ERR_STATE
    assert L.csb_edlines_describe(None, 0) == csb.CSB_ERR_STATE
    assert L.csb_set_blur_generation(None, 3) == csb.CSB_ERR_INVALID
    assert L.csb_edlines_download_descriptors(None, None, None, None, C.c_int64(0)) == csb.CSB_ERR_STATE
    assert C.sizeof(csb.EdlinesStats) == 6 * 8 + 2 * 4 + 4 * 4


def test_line_lbd_mirror_refuses_unported_modes(csb):
    class _Ctx:  # no device needed: the checks happen before any call into the library
        pass
    with pytest.raises(csb.CsbError):
        csb.line_lbd_detect(_Ctx(), numoctaves=2)
    d = csb.line_lbd_detect(_Ctx())
    assert d.use_LSD is True and d.line_length_thres == 50.0  # the reference's default threshold (line_lbd_allclass.cpp:126)
    # use_LSD = False (EDLines, what object_slam selects) routes to csb_edlines_*: exercised on the GPU by tests/test_edlines_gpu.py
    calls = []
    ctx = _Ctx()
    ctx.edlines_detect_batch = lambda g, thr, filt, cap: (calls.append(("edlines", thr, filt)) or ([np.zeros((0, 4), np.float32)], None))
    ctx.lsd_detect_batch = lambda g, thr, filt, cap: (calls.append(("lsd", thr, filt)) or ([np.zeros((0, 4), np.float32)], None))
    d = csb.line_lbd_detect(ctx)
    d.line_length_thres = 15.0
    d.detect_filter_lines(np.zeros((8, 8), np.uint8))
    d.use_LSD = False
    d.detect_filter_lines(np.zeros((8, 8), np.uint8))
    assert calls == [("lsd", 15.0, True), ("edlines", 15.0, True)]


def test_ba_online_entry_points_without_a_context(csb):
    """csb_ba_add_frame / csb_ba_optimize: argument errors, not crashes, with or without a GPU; the mirrored structs have the header's layout."""
    L = csb.lib()
    f = csb.BAFrame()
    idx = C.c_int32(-1)
    assert L.csb_ba_add_frame(None, C.byref(f), C.byref(idx)) == csb.CSB_ERR_INVALID
    assert L.csb_ba_add_frame(None, None, None) == csb.CSB_ERR_INVALID
    assert L.csb_ba_optimize(None, 5, None, None, None) == csb.CSB_ERR_INVALID
    assert L.csb_ba_set_graph(None, None) == csb.CSB_ERR_INVALID
    # csb_ba_frame: pointer, 2 x int32, 2 pointers, int32 (+ pad), 3 pointers, int32 (+ pad), 3 pointers
    assert C.sizeof(csb.BAFrame) == 8 + 8 + 16 + 8 + 24 + 8 + 24
    assert C.sizeof(csb.BAOptimizeStats) == 4 * 4 + 2 * 8 + 2 * 4
    assert C.sizeof(csb.DetectStats) == 5 * 8 + 2 * 4 + 6 * 4


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the reference's CPU path (the oracle port, all host threads) on the bench workload; one JSON line with the
    base contract's keys plus impl / cpu_baseline / e2e (no copies).  One step of one warm-up here; the driver runs it with its own K / W."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "cuboid_proposals_scored_per_sec" and d["unit"] == "proposals/s"
    assert d["value"] > 0 and d["steps"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_tools_and_diagnostics_compile_and_stay_off_the_oracle():
    """Every script under tools/ and tests/diag/ byte-compiles; nothing under tools/, in the package or in bench.py's GPU arm imports the
    oracle or the test helpers (the oracle is test infrastructure: only tests/, smoke() and bench.py's CPU legs may touch it)."""
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, "tools", "*.py")) + glob.glob(os.path.join(root, "tests", "diag", "*.py")):
        py_compile.compile(f, doraise=True)
    for f in glob.glob(os.path.join(root, "tools", "*.py")) + glob.glob(os.path.join(root, "cube_slam_wu_b200", "*.py")):
        src = open(f).read()
        assert "oracle_lib" not in src and "import helpers" not in src and "liboracle" not in src, f
