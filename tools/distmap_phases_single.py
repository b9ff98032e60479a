"""Development probe: stage cycles of k_distmap for ONE frame (8 ROIs) per call, both CTA sizes (needs the CSB_DM_PHASES build, see distmap_phases.py)."""
import ctypes as C
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth, pipeline

ctx = csb.Context(0)
params = csb.DetectParams.default()
batch = synth.make_kitti_batch(1, boxes_per_frame=8, seed=20260925)
frames, boxes, lines, tasks, n_tasks, _, n_map = pipeline.pack_inputs(csb, batch, params, with_maps=False)
gray = np.ascontiguousarray(batch["images"][0].ravel().astype(np.uint8))
names = ["nms", "hysteresis", "cmap+init", "sweeps", "row pass"]
for cta in ("256", "1024"):
    os.environ["CSB_DISTMAP_CTA"] = cta
    ctx.detect_upload_gray(frames, boxes, lines, tasks, n_tasks, gray, params)
    ms = []
    for _ in range(6):
        ctx.detect_run(timed=True)
        _, _, st = ctx.detect_download()
        ms.append(st.gpu_ms_distmap)
    print("CTA %s: gpu_ms_distmap" % cta, ["%.3f" % m for m in ms])
    if not hasattr(csb.lib(), "csb_debug_distmap_phases"):
        continue
    out = np.zeros((n_tasks, 8), np.int64)
    assert csb.lib().csb_debug_distmap_phases(out.ctypes.data_as(C.c_void_p), n_tasks) == 0
    W = out[:, 7] >> 32; H = out[:, 7] & 0xffffffff
    us = out[:, :5] / 1965.0
    for i in range(n_tasks):
        print("  task %d %4dx%-4d " % (i, W[i], H[i]) + "  ".join("%s %.1f" % (n, us[i, k]) for k, n in enumerate(names)) + "  rounds %d  total %.1f" % (out[i, 5], us[i].sum()))
