"""Times the LSD kernels on a batch of synthetic frames: python tools/lsd_time.py [n_frames] [w] [h] [texture]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
tex = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
link = int(sys.argv[5]) if len(sys.argv) > 5 else 0
base = synth.make_lsd_frames(min(n, 16), w, h, seed=3, texture=tex)
frames = np.ascontiguousarray(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n])
ctx = csb.Context(0)
ctx.lsd_upload(frames, unit_link_deg=link)
for i in range(3):
    ctx.lsd_run(timed=True)
    lines, st = ctx.lsd_download()
    print("n=%d %dx%d tex=%g: maps %.3f ms, grow %.3f ms, %d segments, %d regions, %d region px -> %.0f frames/s" % (
        n, w, h, tex, st.gpu_ms_maps, st.gpu_ms_grow, st.n_lines, st.n_regions, st.n_region_px, n / ((st.gpu_ms_maps + st.gpu_ms_grow) * 1e-3)))
    cy = np.array(list(st.grow_cycles), np.float64)
    print("   cycles/frame: grow %.2fM rect %.2fM refine %.2fM nfa %.2fM total %.2fM; grow cycles per region px %.0f" % (*(cy / n / 1e6), cy[0] / max(st.n_region_px, 1)))
    print("   merge rounds %d, unit conflicts %d" % (st.n_merge_rounds, st.n_unit_conflicts))
t = time.time(); lines, st = ctx.lsd_detect_batch(frames); print("host-buffer call: %.1f ms" % ((time.time() - t) * 1e3))
