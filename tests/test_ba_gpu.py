"""GPU parity tests of the BA half: csb_ba_linearize (C ABI) vs the CPU oracle's computeActiveErrors + buildSystem.

Tolerance: 1e-4 absolute on residuals / Jacobians / Hessian blocks (north star).  The Jacobians are central differences
with delta = 1e-9 (base_binary_edge.hpp:147), so a 1-ulp difference in acos/tan/sin/cos between glibc and CUDA is
amplified by 5e8; the achieved agreement is reported and asserted at 2e-5 relative to the block scale.
"""
import os

import numpy as np
import pytest

import helpers as H
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _compare(gpu, ora, tol_rel=2e-5):
    worst = 0.0
    for k in ("ec_err", "ep_err", "eo_err"):
        if gpu[k].size:
            d = np.abs(gpu[k] - ora[k]).max()
            assert d <= 1e-9 * max(1.0, np.abs(ora[k]).max()), "%s differs by %g" % (k, d)
    for k in ("ec_Ji", "ec_Jj", "ep_Ji", "ep_Jj", "eo_Ji", "eo_Jj", "H_cam", "b_cam", "H_cube", "b_cube", "ec_Hij", "ep_Hij", "eo_Hij"):
        if gpu[k].size:
            scale = max(1.0, np.abs(ora[k]).max())
            d = np.abs(gpu[k] - ora[k]).max()
            assert d <= H.TOL_NORTH_STAR * scale, "%s differs by %g (scale %g)" % (k, d, scale)
            worst = max(worst, d / scale)
    assert abs(gpu["chi2"][0] - ora["chi2"]) <= 1e-9 * max(1.0, abs(ora["chi2"]))
    assert worst <= tol_rel, "relative block error %g" % worst
    return worst


def _run(ctx, g):
    ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=g["ec"], ep=g["ep"], eo=g["eo"])
    gpu = ctx.ba_linearize(g["cams7"], g["cubes10"], jacobians=True)
    E = O.ba_edges(ec=g["ec"], ep=g["ep"], eo=g["eo"])
    ora = O.ba_linearize(g["cams7"], g["cam_fixed"], g["cubes10"], g["cube_fixed"], E)
    return gpu, ora


def test_small_graph_all_edge_types(ctx):
    from cube_slam_wu_b200 import synth
    g = synth.make_ba_graph(n_cam=12, n_cube=3, obs_per_cube=6, seed=1, with_proj=True)
    gpu, ora = _run(ctx, g)
    _compare(gpu, ora)
    # fixed camera 0: no Jacobian / Hessian for it
    assert np.all(gpu["H_cam"][0] == 0) and np.all(gpu["b_cam"][0] == 0)


def test_config4_graph(ctx):
    """BASELINE config #4: 200 keyframes, 50 cuboids, 4000 EdgeSE3Cuboid + 199 EdgeSE3Expmap."""
    from cube_slam_wu_b200 import synth
    g = synth.make_ba_graph()
    assert len(g["ec"][0]) == 4000
    gpu, ora = _run(ctx, g)
    _compare(gpu, ora)
    # resident path: same blocks without Jacobians materialised, idempotent
    ctx.ba_upload_estimates(g["cams7"], g["cubes10"])
    ctx.ba_run()
    r1 = ctx.ba_download()
    ctx.ba_run()
    r2 = ctx.ba_download()
    for k in ("H_cam", "H_cube", "b_cam", "b_cube", "ec_Hij", "eo_Hij"):
        assert np.array_equal(r1[k], r2[k]) and np.array_equal(r1[k], gpu[k])
    # symmetry of the diagonal blocks (size-independent property)
    Hc = gpu["H_cube"].reshape(-1, 9, 9)
    assert np.abs(Hc - Hc.transpose(0, 2, 1)).max() <= 1e-6 * np.abs(Hc).max()


def test_fixed_cuboid_and_yaw_flips(ctx):
    from cube_slam_wu_b200 import synth
    g = synth.make_ba_graph(n_cam=20, n_cube=4, obs_per_cube=10, seed=3)
    g["cube_fixed"][1] = 1
    g["cam_fixed"][5] = 1
    gpu, ora = _run(ctx, g)
    _compare(gpu, ora)
    assert np.all(gpu["H_cube"][1] == 0)


def test_tum_offline_fixture(ctx):
    """The reference's own offline BA inputs (object_slam/data, via tests/golden/tum_ba.npz): graph of the last frame as
    main_obj.cpp:686-800 builds it, initial estimates from the oracle's incremental run."""
    d = np.load(os.path.join(H.GOLDEN, "tum_ba.npz"))
    g = tum_graph(d)
    gpu, ora = _run(ctx, g)
    _compare(gpu, ora)


def tum_graph(d, optimise_each_frame=True):
    """Replays main_obj.cpp's offline loop with the oracle's LM and returns the final-frame graph."""
    det, pop, truth = d["det"], d["pop"], d["truth"]
    N = len(truth)
    ident = np.array([0, 0, 0, 0, 0, 0, 1.0])
    Twc0 = O.se3_mul(truth[0, 1:8], ident)
    cams, fixed = [], []
    ec = ([], [], [], [])
    eo = ([], [], [], [])
    cube = None
    row = 0
    for f in range(N):
        odom = ident.copy()
        if f == 0:
            Twc = Twc0
        else:
            prev = cams[f - 1]
            if f > 1:
                odom = O.se3_mul(prev, O.se3_inverse(cams[f - 2]))
            Twc = O.se3_inverse(O.se3_mul(odom, prev))
        has = row < len(det) and int(det[row, 0]) == f
        if has:
            m = det[row]
            cg = O.cuboid_from_minimal([m[1], m[2], m[3], 0, 0, m[4], m[5], m[6], m[7]])
            meas = O.cuboid_transform_to(cg, O.se3_mul(pop[f, 1:8], ident))
            q = (1 - m[8] + 0.5) / 2
            row += 1
        if f == 0:
            cube = O.cuboid_transform_from(meas, Twc)
        cams.append(O.se3_inverse(Twc)); fixed.append(1 if f == 0 else 0)
        if has:
            ec[0].append(f); ec[1].append(0); ec[2].append(meas); ec[3].append((np.eye(9) * (2 * q) ** 2).ravel())
        if f > 0:
            eo[0].append(f - 1); eo[1].append(f); eo[2].append(odom); eo[3].append(np.eye(6).ravel())
        if optimise_each_frame:
            E = O.ba_edges(ec=(ec[0], ec[1], np.array(ec[2]), np.array(ec[3])),
                           eo=(eo[0], eo[1], np.array(eo[2]).reshape(-1, 7), np.array(eo[3]).reshape(-1, 36)) if f > 0 else None)
            c2, q2, _, _ = O.ba_optimize(np.array(cams), fixed, cube.reshape(1, 10), [0], E, 5)
            cams = [c2[i] for i in range(f + 1)]
            cube = q2[0]
    return dict(cams7=np.array(cams), cubes10=cube.reshape(1, 10), cam_fixed=np.array(fixed, np.int32), cube_fixed=np.zeros(1, np.int32),
                ec=(np.array(ec[0], np.int32), np.array(ec[1], np.int32), np.array(ec[2]), np.array(ec[3])), ep=None,
                eo=(np.array(eo[0], np.int32), np.array(eo[1], np.int32), np.array(eo[2]).reshape(-1, 7), np.array(eo[3]).reshape(-1, 36)))


def test_analytic_jacobians_match_numeric(ctx):
    """closed-form Jacobians of EdgeSE3Cuboid / EdgeSE3Expmap (csb_ba_set_jacobian_mode) against the reference's
    definition, the delta = 1e-9 central differences (oracle, base_binary_edge.hpp:130-205).  Bar: 1e-4 relative to the block scale
    (north star); the numeric Jacobians themselves carry ~1e-7 of round-off.  Residuals and chi2 do not depend on the mode."""
    from cube_slam_wu_b200 import synth
    for g in (synth.make_ba_graph(n_cam=12, n_cube=3, obs_per_cube=6, seed=4), synth.make_ba_graph()):
        ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=g["ec"], ep=g["ep"], eo=g["eo"])
        E = O.ba_edges(ec=g["ec"], ep=g["ep"], eo=g["eo"])
        ora = O.ba_linearize(g["cams7"], g["cam_fixed"], g["cubes10"], g["cube_fixed"], E)
        try:
            ctx.ba_set_jacobian_mode(True)
            ana = ctx.ba_linearize(g["cams7"], g["cubes10"], jacobians=True)
        finally:
            ctx.ba_set_jacobian_mode(False)
        num = ctx.ba_linearize(g["cams7"], g["cubes10"], jacobians=True)
        worst = 0.0
        for k in ("ec_err", "eo_err"):
            assert np.array_equal(ana[k], num[k])
        for k in ("ec_Ji", "ec_Jj", "eo_Ji", "eo_Jj", "H_cam", "b_cam", "H_cube", "b_cube", "ec_Hij", "eo_Hij"):
            scale = max(1.0, np.abs(ora[k]).max())
            d = np.abs(ana[k] - ora[k]).max() / scale
            assert d <= H.TOL_NORTH_STAR, "%s: analytic vs the reference's numeric Jacobians differ by %g of the block scale" % (k, d)
            worst = max(worst, d)
        assert worst <= 5e-6, worst  # observed ~3e-7: the round-off of the central differences
        # fixed vertices: untouched in both modes
        assert np.all(ana["H_cam"][0] == 0) and np.all(ana["b_cam"][0] == 0)
        print("analytic vs numeric Jacobians: worst %.3g of the block scale" % worst)


def test_analytic_mode_optimizes_to_the_same_point(ctx):
    from cube_slam_wu_b200 import synth
    g = synth.make_ba_graph(n_cam=40, n_cube=8, obs_per_cube=20, seed=9)
    ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=g["ec"], ep=g["ep"], eo=g["eo"])
    res = []
    for analytic in (False, True):
        try:
            ctx.ba_set_jacobian_mode(analytic)
            ctx.ba_upload_estimates(g["cams7"], g["cubes10"])
            res.append(ctx.ba_optimize(5))
        finally:
            ctx.ba_set_jacobian_mode(False)
    (c0, q0, s0), (c1, q1, s1) = res
    assert s0.iterations == s1.iterations
    assert abs(s0.chi2 - s1.chi2) <= 1e-5 * max(1.0, s0.chi2)
    # Both runs stop by g2o's rule (chi2 gain below 1e-3 of chi2) after the same number of iterations, at the same chi2 to 1e-10; the estimates
    # then agree as far as the numeric Jacobians' own round-off (4e-6 of the block scale, test above) pins a point in the flat directions of
    # the cost: observed 3e-6 on the cameras, 1.2e-5 on the cuboids (it moves in the last digit with the summation order of the linear solve).
    # Bar: the north star's 1e-4 for poses / scales.
    assert np.abs(c0 - c1).max() < 1e-4 and np.abs(q0 - q1).max() < 1e-4
