"""Times the EDLines kernels (and the descriptor stage on their key lines) on a batch of synthetic frames, checking the first frames against the
oracle: python tests/diag/edlines_time.py [n_frames] [w] [h]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth
import oracle_lib as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
base = synth.make_lsd_frames(min(n, 32), w, h, seed=3)
frames = np.ascontiguousarray(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n])
ctx = csb.Context(0)
for i in range(3):
    lines, st = ctx.edlines_detect_batch(frames)
    print("n=%d %dx%d: maps %.3f ms, draw %.3f ms, fit %.3f ms, %d segments, %d chains, %d chain px, %d frames failed -> %.0f frames/s" % (
        n, w, h, st.gpu_ms_maps, st.gpu_ms_draw, st.gpu_ms_fit, st.n_lines, st.n_chains, st.n_chain_px, st.n_frames_failed,
        n / ((st.gpu_ms_maps + st.gpu_ms_draw + st.gpu_ms_fit) * 1e-3)))
ok = True
for f in range(min(n, 4)):
    ref, _ = O.edlines_detect(frames[f])
    ok = ok and lines[f].shape == ref.shape and np.array_equal(lines[f].view(np.uint32), ref.view(np.uint32))
print("first frames identical to the oracle:", ok)
out = ctx.edlines_detect_describe_batch(frames[:min(n, 8)])
ref, extra = O.edlines_detect(frames[0])
_, r32 = O.lbd_describe_keylines(frames[0], ref, extra[:, 0], extra[:, 1])
print("descriptors of frame 0 identical to the oracle:", np.array_equal(out["desc"][0], r32))
