"""csb_ba_add_frame (the online caller of object_slam/src/main_obj.cpp:738-803 adds one keyframe and its edges per frame): the graph grown
frame by frame must be the very device state csb_ba_set_graph builds from the concatenated arrays -- same edge order, so bit-identical
blocks -- with the device-resident estimates preserved across appends."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _by_camera(g):
    """ec edges grouped by camera (ascending), eo edges by their later camera: the order frame-by-frame insertion produces."""
    ec_cam, ec_cube, ec_meas, ec_info = [np.asarray(a) for a in g["ec"]]
    o = np.argsort(ec_cam, kind="stable")
    ec = (ec_cam[o], ec_cube[o], ec_meas[o], ec_info[o])
    eo_i, eo_j, eo_meas, eo_info = [np.asarray(a) for a in g["eo"]]
    assert np.all(eo_i < eo_j)
    o = np.argsort(eo_j, kind="stable")
    return ec, (eo_i[o], eo_j[o], eo_meas[o], eo_info[o])


def test_frame_by_frame_equals_set_graph(ctx, csb):
    from cube_slam_wu_b200 import synth
    g = synth.make_ba_graph(n_cam=24, n_cube=6, obs_per_cube=10, seed=5)
    ec, eo = _by_camera(g)
    cams7, cubes10 = np.asarray(g["cams7"]), np.asarray(g["cubes10"])
    ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=ec, eo=eo)
    ref = ctx.ba_linearize(cams7, cubes10, jacobians=False)

    # the same graph grown camera by camera; a cuboid enters with the first frame that observes it, which renumbers the cuboids
    first_seen = {}
    for cam, cube in zip(ec[0], ec[1]):
        first_seen.setdefault(int(cube), len(first_seen))
    for j in range(len(cubes10)):
        first_seen.setdefault(j, len(first_seen))  # never observed: appended with the last frame
    ctx.ba_set_graph([], [], ec=None, eo=None)
    n_known = 0
    for cam in range(len(cams7)):
        m = ec[0] == cam
        cubes_here = [first_seen[int(c)] for c in ec[1][m]]
        want = max(cubes_here) + 1 if cubes_here else n_known
        if cam == len(cams7) - 1:
            want = len(cubes10)
        new_ids = [j for j in range(len(cubes10)) if n_known <= first_seen[j] < want]
        new_ids.sort(key=lambda j: first_seen[j])
        mo = eo[1] == cam
        idx = ctx.ba_add_frame(cams7[cam], cam_fixed=g["cam_fixed"][cam], new_cubes10=cubes10[new_ids] if new_ids else None,
                               new_cube_fixed=[g["cube_fixed"][j] for j in new_ids] if new_ids else None,
                               ec=(np.array(cubes_here, np.int32), ec[2][m], ec[3][m]), eo=(eo[0][mo], eo[2][mo], eo[3][mo]))
        assert idx == cam
        n_known = max(n_known, want)
    ctx.ba_run()
    got = ctx.ba_download(jacobians=False)
    perm = np.array([j for j in sorted(range(len(cubes10)), key=lambda j: first_seen[j])])  # new index -> old index
    assert np.array_equal(got["H_cam"], ref["H_cam"]) and np.array_equal(got["b_cam"], ref["b_cam"])
    assert np.array_equal(got["H_cube"], ref["H_cube"][perm]) and np.array_equal(got["b_cube"], ref["b_cube"][perm])
    assert np.array_equal(got["ec_Hij"], ref["ec_Hij"]) and np.array_equal(got["eo_Hij"], ref["eo_Hij"])
    assert got["chi2"][0] == ref["chi2"][0]


def test_estimates_survive_an_append_and_optimize_runs(ctx, csb):
    """optimize, append a frame, optimize again: the optimised estimates stay on the device (the online mode's warm start)."""
    from cube_slam_wu_b200 import synth
    g = synth.make_ba_graph(n_cam=13, n_cube=3, obs_per_cube=8, seed=9)
    ec, eo = _by_camera(g)
    cams7, cubes10 = np.asarray(g["cams7"]), np.asarray(g["cubes10"])
    last = len(cams7) - 1
    m, mo = ec[0] < last, eo[1] < last
    ctx.ba_set_graph(g["cam_fixed"][:last], g["cube_fixed"], ec=tuple(a[m] for a in ec), eo=tuple(a[mo] for a in eo))
    ctx.ba_upload_estimates(cams7[:last], cubes10)
    c1, q1, st1 = ctx.ba_optimize(5)
    ctx.ba_add_frame(cams7[last], cam_fixed=False, ec=(ec[1][~m], ec[2][~m], ec[3][~m]), eo=(eo[0][~mo], eo[2][~mo], eo[3][~mo]))
    c2, q2, st2 = ctx.ba_optimize(0)  # no iteration: just read the estimates back
    assert np.array_equal(c2[:last], c1) and np.array_equal(q2, q1) and np.array_equal(c2[last], cams7[last])
    c3, q3, st3 = ctx.ba_optimize(5)
    assert st3.iterations >= 1 and np.isfinite(st3.chi2)
    # same result as loading the whole graph with those warm-start estimates
    ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=ec, eo=eo)
    ctx.ba_upload_estimates(c2, q2)
    c4, q4, st4 = ctx.ba_optimize(5)
    assert st4.iterations == st3.iterations and np.array_equal(c4, c3) and np.array_equal(q4, q3)


def test_download_of_jacobians_that_were_not_computed_is_refused(ctx, csb):
    from cube_slam_wu_b200 import synth
    g = synth.make_ba_graph(n_cam=12, n_cube=3, obs_per_cube=6, seed=3)
    ctx.ba_set_graph(g["cam_fixed"], g["cube_fixed"], ec=g["ec"], eo=g["eo"])
    ctx.ba_upload_estimates(g["cams7"], g["cubes10"])
    ctx.ba_run()
    with pytest.raises(Exception):
        ctx.ba_download(jacobians=True)
    assert "chi2" in ctx.ba_download(jacobians=False)
