"""Development probe: BA half on BASELINE config #4 (200 keyframes, 50 cuboids, ~4k edges): linearisation (numeric / analytic) and the
device LM.  `python tools/ba_time.py [reps]`; under `ncu --metrics gpu__time_duration.sum` it yields the per-kernel launch list."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = csb.Context(0)
g = synth.make_ba_graph()
ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=g["ec"], ep=g["ep"], eo=g["eo"])
ctx.ba_upload_estimates(g["cams7"], g["cubes10"])
n_edges = len(g["ec"][0]) + len(g["eo"][0])
for mode, name in ((False, "numeric"), (True, "analytic")):
    ctx.ba_set_jacobian_mode(mode)
    for _ in range(5):
        ctx.ba_run()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.ba_run()
    ctx.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print("linearise %-8s: %.4f ms  (%d edges, %.3g edges/s)" % (name, ms, n_edges, n_edges / (ms * 1e-3)))
ctx.ba_set_jacobian_mode(False)
for _ in range(3):
    ctx.ba_upload_estimates(g["cams7"], g["cubes10"])
    t0 = time.perf_counter()
    _, _, so = ctx.ba_optimize(5)
    wall = 1e3 * (time.perf_counter() - t0)
    print("optimize(5): %.3f ms gpu, %.3f ms wall, %d iterations, %d linear solves, %d kernels in %d launches, chi2 %.6g" % (so.gpu_ms, wall, so.iterations, so.trials, so.n_kernel_launches, so.n_launches, so.chi2))
ctx.close()
