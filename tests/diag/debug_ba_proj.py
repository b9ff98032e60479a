"""GPU debug helper: prints GPU vs oracle Jacobians of the EdgeSE3CuboidProj edges of the small test graph."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cube_slam_wu_b200 as csb
from cube_slam_wu_b200 import synth
import oracle_lib as O

np.set_printoptions(precision=6, linewidth=200, suppress=False)
g = synth.make_ba_graph(n_cam=12, n_cube=3, obs_per_cube=6, seed=1, with_proj=True)
ctx = csb.Context(0)
for variant in ("all", "ep_only"):
    ec, eo = (g["ec"], g["eo"]) if variant == "all" else (None, None)
    ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=ec, ep=g["ep"], eo=eo)
    gpu = ctx.ba_linearize(g["cams7"], g["cubes10"], jacobians=True)
    E = O.ba_edges(ec=ec, ep=g["ep"], eo=eo)
    ora = O.ba_linearize(g["cams7"], g["cam_fixed"], g["cubes10"], g["cube_fixed"], E)
    print("==", variant)
    for e in range(len(g["ep"][0])):
        dj = np.abs(gpu["ep_Ji"][e] - ora["ep_Ji"][e]); dk = np.abs(gpu["ep_Jj"][e] - ora["ep_Jj"][e])
        print("edge", e, "err diff", np.abs(gpu["ep_err"][e] - ora["ep_err"][e]).max(), "Ji maxdiff", dj.max(), "at", dj.argmax(), "Jj maxdiff", dk.max(), "at", dk.argmax())
    e = 0
    print("gpu Ji[0]", gpu["ep_Ji"][e].reshape(6, 4)); print("ora Ji[0]", ora["ep_Ji"][e].reshape(6, 4))
    print("gpu Jj[0]", gpu["ep_Jj"][e].reshape(9, 4)); print("ora Jj[0]", ora["ep_Jj"][e].reshape(9, 4))
