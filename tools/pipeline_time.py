"""Times the chained per-frame pipeline on one GPU (BASELINE config #5's frame stage): gray frames -> LSD line table -> Canny + distance
transform + cuboid proposals -> best cuboid per box.   python tools/pipeline_time.py [n_frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = synth.make_kitti_batch(min(n, 16), boxes_per_frame=8, seed=5)
reps = (n + len(base["K"]) - 1) // len(base["K"])
gray16 = np.stack(base["images"]).astype(np.uint8)
gray = np.ascontiguousarray(np.concatenate([gray16] * reps)[:n])
nb = 8
K = np.concatenate([base["K"]] * reps)[:n]; T = np.concatenate([base["T"]] * reps)[:n]
boxes = np.concatenate([np.asarray(base["boxes"]).reshape(-1, 5)] * reps)[:n * nb]
ctx = csb.Context(0)
p = csb.DetectParams.default()
for it in range(3):
    t0 = time.perf_counter()
    lines, lst = ctx.lsd_detect_batch(gray)
    t1 = time.perf_counter()
    ranges, off = [], 0
    for a in lines:
        ranges.append((off, off + len(a))); off += len(a)
    L = np.concatenate(lines).astype(np.float64)
    frames = csb.make_frames(K, T, base["img_w"], base["img_h"], [(nb * f, nb * f + nb) for f in range(n)], ranges)
    tasks, n_tasks, n_map = csb.detect_plan(frames, boxes, p)
    t2 = time.perf_counter()
    cub, ncub, st = ctx.detect_batch_gray(frames, boxes, L, tasks, n_tasks, gray.ravel(), p)
    t3 = time.perf_counter()
    print("n=%d frames: lsd %.1f ms (%d segments), host glue %.1f ms, detect_gray %.1f ms (%d scored) -> %.0f frames/s end to end"
          % (n, 1e3 * (t1 - t0), off, 1e3 * (t2 - t1), 1e3 * (t3 - t2), st.n_scored, n / (t3 - t0)))
