"""Development probe: per-task stage cycles of k_distmap on the bench workload (run on a GPU box; needs a library built with the counters:
CSB_DM_PHASES=1 python -m cube_slam_wu_b200.build --force)."""
import ctypes as C
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth, pipeline

ctx = csb.Context(0)
params = csb.DetectParams.default()
batch = synth.make_kitti_batch(64, boxes_per_frame=8, seed=20260925)
frames, boxes, lines, tasks, n_tasks, _, n_map = pipeline.pack_inputs(csb, batch, params, with_maps=False)
gray = np.ascontiguousarray(np.concatenate([im.ravel() for im in batch["images"]]).astype(np.uint8))
ctx.detect_upload_gray(frames, boxes, lines, tasks, n_tasks, gray, params)
ms = []
for _ in range(5):
    ctx.detect_run(timed=True)
    _, _, st = ctx.detect_download()
    ms.append(st.gpu_ms_distmap)
print("gpu_ms_distmap", ms)
if not hasattr(csb.lib(), "csb_debug_distmap_phases"):
    print("(library built without CSB_DM_PHASES: no per-task stage cycles)")
    sys.exit(0)
out = np.zeros((n_tasks, 8), np.int64)
rc = csb.lib().csb_debug_distmap_phases(out.ctypes.data_as(C.c_void_p), n_tasks)
assert rc == 0, rc
W = out[:, 7] >> 32; H = out[:, 7] & 0xffffffff
us = out[:, :5] / 1965.0
names = ["nms", "hysteresis", "cmap+init", "sweeps", "row pass"]
tot = us.sum(1)
print("tasks %d; ROI px: mean %.0f max %d; W max %d H max %d" % (n_tasks, (W * H).mean(), (W * H).max(), W.max(), H.max()))
print("per-task total us: mean %.1f  p50 %.1f  p90 %.1f  max %.1f" % (tot.mean(), np.median(tot), np.percentile(tot, 90), tot.max()))
for i, n in enumerate(names):
    print("  %-18s mean %7.1f us   max %7.1f us   share %5.1f %%" % (n, us[:, i].mean(), us[:, i].max(), 100 * us[:, i].sum() / tot.sum()))
print("hysteresis rounds: mean %.1f max %d" % (out[:, 5].mean(), out[:, 5].max()))
o = np.argsort(-tot)[:8]
for i in o:
    print("  task %4d  %4dx%-4d  " % (i, W[i], H[i]) + "  ".join("%s %.1f" % (n, us[i, k]) for k, n in enumerate(names)) + "  rounds %d" % out[i, 5])
