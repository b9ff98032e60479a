import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cube_slam_wu_b200 as csb
import oracle_lib as O
ctx = csb.Context(0)
rng = np.random.default_rng(1)
for name, fr in (("uniform noise", rng.integers(0, 256, (4, 480, 640)).astype(np.uint8)),
                 ("smooth noise", None), ("checker", None)):
    if name == "smooth noise":
        import cv2
        fr = np.stack([np.clip(128 + 60 * cv2.GaussianBlur(rng.normal(0, 1, (480, 640)), (0, 0), 1.5) * 3, 0, 255).astype(np.uint8) for _ in range(4)])
    if name == "checker":
        yy, xx = np.mgrid[0:480, 0:640]
        fr = np.stack([(((yy // 8 + xx // 8) % 2) * 200 + 20).astype(np.uint8)] * 4)
    t = time.time(); lines, st = ctx.lsd_detect_batch(fr, filter=False); dt = time.time() - t
    t = time.time(); ref = O.lsd_detect(fr[0], mode=0); dto = time.time() - t
    ok = lines[0].shape == ref.shape and (len(ref) == 0 or np.abs(lines[0] - ref).max() <= 1e-4)
    print("%-14s gpu %.1f ms for 4 frames (grow %.1f ms), %d segments, %d regions, merge rounds %d, conflicts %d | oracle %.1f ms/frame | match %s"
          % (name, dt * 1e3, st.gpu_ms_grow, st.n_lines, st.n_regions, st.n_merge_rounds, st.n_unit_conflicts, dto * 1e3, ok))
