"""Times the LBD descriptor kernels chained behind the line detector (BASELINE config #3): python tools/lbd_time.py [n_frames] [w] [h]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
base = synth.make_lsd_frames(min(n, 32), w, h, seed=3)
frames = np.ascontiguousarray(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n])
ctx = csb.Context(0)
ctx.lsd_upload(frames)
ctx.lsd_run()
for i in range(4):
    ctx.lbd_run_on_lsd(timed=True)
    out = ctx.lbd_download()
    st = out["stats"]
    print("n=%d %dx%d: grad %.3f ms (%.0f GB/s of 5 B/px), describe %.3f ms, %d lines, %d samples (%.1f Gsamples/s)" % (
        n, w, h, st.gpu_ms_grad, n * w * h * 5 / st.gpu_ms_grad / 1e6, st.gpu_ms_describe, st.n_lines, st.n_samples, st.n_samples / st.gpu_ms_describe / 1e6))
