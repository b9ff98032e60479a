"""Development probe: cycle share of every k_score phase on the bench workload (run on a GPU box; needs a library built with the counters:
CSB_SCORE_PHASES=1 python -m cube_slam_wu_b200.build --force)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth
from cube_slam_wu_b200 import pipeline

ctx = csb.Context(0)
params = csb.DetectParams.default()
batch = synth.make_kitti_batch(64, boxes_per_frame=8, seed=20260925)
frames, boxes, lines, tasks, n_tasks, maps, n_map = pipeline.pack_inputs(csb, batch, params)
ctx.detect_upload(frames, boxes, lines, tasks, n_tasks, maps, n_map, params)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for _ in range(3):
    ctx.detect_run(timed=True)
ctx.detect_download()
ctx.score_phases(reset=True)
ms = []
for _ in range(reps):
    ctx.detect_run(timed=True)
    _, _, st = ctx.detect_download()
    ms.append(st.gpu_ms_score)
ph = ctx.score_phases(reset=True).astype(np.float64)
names = ["fetch/wait", "lines+VPs", "VP support", "phase1 corners", "prefix", "map wait", "phase2 score", "exit"]
tot = ph[:7].sum()
print("k_score %.3f ms; cycles per CTA per run %.0f (busy) ; 148 CTAs" % (np.mean(ms), tot / reps / 148))
print("VP-support units per run: float %d, double %d, exact %d" % tuple(int(x / reps) for x in ph[8:11]))
for n, v in zip(names, ph):
    print("  %-16s %6.2f %%   %.1f us/CTA/run" % (n, 100 * v / tot, v / reps / 148 / 1965.0))
