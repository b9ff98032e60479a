"""Development probe: per-kernel CUDA-event times of the resident proposal path on the bench workload (config #2), one step at a time."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import pipeline, synth

ctx = csb.Context(0)
params = csb.DetectParams.default()
batch = synth.make_kitti_batch(64, boxes_per_frame=8, seed=20260925)
frames, boxes, lines, tasks, n_tasks, maps, n_map = pipeline.pack_inputs(csb, batch, params)
ctx.detect_upload(frames, boxes, lines, tasks, n_tasks, maps, n_map, params)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    ctx.detect_run(timed=True)
ctx.detect_download()
rows = []
for _ in range(reps):
    ctx.detect_run(timed=True)
    _, _, st = ctx.detect_download()
    rows.append((st.gpu_ms_prep, st.gpu_ms_score, st.gpu_ms_select, st.gpu_ms_recover, st.gpu_ms_rank))
m = np.median(np.array(rows), axis=0)
print("scored %d enumerated %d | prep %.4f score %.4f select %.4f recover %.4f rank %.4f ms (median of %d, L2 warm)" % ((st.n_scored, st.n_enumerated) + tuple(m) + (reps,)))
print("k_score roofline fraction at 550 B / proposal: %.3f of 6548.5 GB/s" % (550.0 * st.n_scored / (m[1] * 1e-3) / 1e9 / 6548.5))
