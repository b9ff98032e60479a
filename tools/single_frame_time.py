"""Development probe: per-kernel CUDA-event times of ONE frame (8 boxes) per blocking csb_detect_batch_gray() call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth, pipeline

ctx = csb.Context(0)
params = csb.DetectParams.default()
batch = synth.make_kitti_batch(16, boxes_per_frame=8, seed=20260925)
rows, wall = [], []
subs = []
for f in range(16):
    b0, b1 = batch["box_ranges"][f]; l0, l1 = batch["line_ranges"][f]
    sub = dict(K=batch["K"][f:f + 1], T=batch["T"][f:f + 1], boxes=batch["boxes"][b0:b1], lines=batch["lines"][l0:l1], box_ranges=[(0, b1 - b0)],
               line_ranges=[(0, l1 - l0)], images=[batch["images"][f]], img_w=batch["img_w"], img_h=batch["img_h"])
    fr, bx, ln, tk, nt, _, _ = pipeline.pack_inputs(csb, sub, params, with_maps=False)
    t = torch.from_numpy(np.ascontiguousarray(batch["images"][f].ravel(), np.uint8)).pin_memory()
    subs.append((fr, bx, ln, tk, nt, t, t.numpy()))
for rep in range(4):
    for (fr, bx, ln, tk, nt, t, g) in subs:
        t0 = time.perf_counter()
        _, _, st = ctx.detect_batch_gray(fr, bx, ln, tk, nt, g, params, want_stats=True)
        if rep:
            wall.append(time.perf_counter() - t0)
            rows.append((st.gpu_ms_distmap, st.gpu_ms_prep, st.gpu_ms_score, st.gpu_ms_select, st.gpu_ms_recover, st.gpu_ms_rank))
m = np.median(np.array(rows), axis=0)
print("one frame per call: wall %.3f ms median | distmap %.3f prep(wait) %.3f score %.3f select %.3f recover %.3f rank %.3f = %.3f ms on the GPU" % (
    (1e3 * np.median(wall),) + tuple(m) + (m.sum(),)))
