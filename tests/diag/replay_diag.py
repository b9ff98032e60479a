"""Stage-by-stage comparison of the GPU replay with the oracle replay (tests/replay.py) on the TUM sequence: the oracle drives the state; at
every frame each GPU stage gets the oracle's inputs, so the first stage that differs is named.  Diagnostic tool (needs a GPU)."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cube_slam_wu_b200 as csb
import replay


class Both:
    def __init__(self, ctx, use_lsd):
        self.o = replay.OracleBackend(use_lsd=use_lsd)
        self.g = replay.GpuBackend(ctx, csb, use_lsd=use_lsd)
        self.f = 0

    def lines(self, gray):
        a = self.o.lines(gray)
        try:
            b = self.g.lines(gray)
            same = a.shape == b.shape and np.array_equal(a, b)
            if not same:
                print("frame %d LINES differ: oracle %s gpu %s maxdiff %s" % (self.f, a.shape, b.shape, np.abs(a - b).max() if a.shape == b.shape else "n/a"))
        except Exception:
            print("frame %d LINES raised" % self.f); traceback.print_exc()
        return a

    def best_cuboid(self, gray, T0, box, lines, sample):
        a = self.o.best_cuboid(gray, T0, box, lines, sample)
        try:
            b = self.g.best_cuboid(gray, T0, box, lines, sample)
            if (a is None) != (b is None):
                print("frame %d CUBOID none-ness differs" % self.f, a, b)
            elif a is not None:
                d = max(np.abs(a["pos"] - b["pos"]).max(), abs(a["rotY"] - b["rotY"]), np.abs(a["scale"] - b["scale"]).max(), abs(a["err"] - b["err"]),
                        abs(a["droll"] - b["droll"]), abs(a["dpitch"] - b["dpitch"]))
                if d > 1e-9:
                    print("frame %d CUBOID differs by %.3g: oracle %s\n gpu %s" % (self.f, d, a, b))
        except Exception:
            print("frame %d CUBOID raised" % self.f); traceback.print_exc()
        return a

    def optimize(self, cams, fixed, cube, ec, eo):
        c2, q2 = self.o.optimize(cams, fixed, cube, ec, eo)
        try:
            g2, gq = self.g.optimize(cams, fixed, cube, ec, eo)
            dc = np.abs(np.asarray(g2)[:len(cams)] - np.asarray(c2)[:len(cams)]).max()
            dq = np.abs(gq - q2).max()
            print("frame %d OPT cams %.3g cube %.3g" % (self.f, dc, dq))
        except Exception:
            print("frame %d OPT raised" % self.f); traceback.print_exc()
        self.f += 1
        return c2, q2


if __name__ == "__main__":
    frames, boxes, truth, out_obj, out_cam = replay.load_sequence()
    ctx = csb.Context(0)
    for lsd in (1, 0):
        print("=== use_lsd=%d" % lsd)
        n = int(sys.argv[1]) if len(sys.argv) > 1 else None
        replay.run(Both(ctx, lsd), frames, boxes, truth, n_frames=n)
    ctx.close()
