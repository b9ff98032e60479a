"""CPU tests of the host-side graph assembly (cube_slam_wu_b200/graph.py; BASELINE config #5: observation records gathered over the ranks
-> camera-object graph, the recipe of object_slam/src/main_obj.cpp:738-803), against the oracle's SE3 / cuboid functions and its
linearisation."""
import numpy as np

import oracle_lib as O
from cube_slam_wu_b200 import graph, synth


def _records(rng, cams_wc, n_lm, drop=0.2):
    truth = np.zeros((n_lm, 10))
    for l in range(n_lm):
        yaw = rng.uniform(-np.pi, np.pi)
        truth[l] = np.r_[rng.uniform(-5, 5, 2), 0.5, 0, 0, np.sin(yaw / 2), np.cos(yaw / 2), rng.uniform(0.3, 1.5, 3)]
    rec = []
    for f, Twc in enumerate(cams_wc):
        for b in range(n_lm):
            r = np.zeros(16)
            r[0], r[1] = f, b
            if rng.uniform() > drop:
                noisy = truth[b].copy()
                noisy[:3] += rng.normal(0, 0.05, 3)
                r[2] = 1
                r[3] = rng.uniform(0.25, 0.75)
                r[4:14] = O.cuboid_transform_to(noisy, Twc)
                r[14] = 1.5 - 2 * r[3]
            rec.append(r)
    return np.array(rec), truth


def test_se3_helpers_match_oracle():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = np.r_[rng.normal(size=3), rng.normal(size=4)]; a[3:] /= np.linalg.norm(a[3:])
        b = np.r_[rng.normal(size=3), rng.normal(size=4)]; b[3:] /= np.linalg.norm(b[3:])
        assert np.allclose(graph.se3_mul(a, b), O.se3_mul(a, b), atol=1e-14)
        assert np.allclose(graph.se3_inv(a), O.se3_inverse(a), atol=1e-14)
    R = synth.euler_zyx_to_rot(0.3, -0.2, 2.5)
    q = graph.rot_to_quat(np.asarray(R))
    assert abs(np.linalg.norm(q) - 1) < 1e-14
    v = rng.normal(size=3)
    assert np.allclose(graph._qrot(q, v), np.asarray(R) @ v, atol=1e-14)


def test_assembled_graph_follows_the_reference_recipe():
    rng = np.random.default_rng(1)
    g0 = synth.make_ba_graph(n_cam=12, n_cube=3, obs_per_cube=6, seed=3)
    cams_wc = np.array([O.se3_inverse(c) for c in g0["cams7"]])
    rec, truth = _records(rng, cams_wc, 4)
    shuffled = rec[rng.permutation(len(rec))]           # gathered records arrive rank by rank, not frame by frame
    g = graph.assemble_graph(shuffled, cams_wc, 4)
    assert g["cam_fixed"].tolist() == [1] + [0] * 11 and g["cube_fixed"].tolist() == [0] * 4
    n_valid = int((rec[:, 2] == 1).sum())
    assert len(g["ec"][0]) == n_valid and len(g["eo"][0]) == 11
    assert np.all(np.diff(g["ec"][0]) >= 0)              # frame after frame
    for l in range(4):                                   # landmark = first valid observation moved to the world frame (main_obj.cpp:745-751)
        first = next(r for r in rec if r[1] == l and r[2] == 1)
        assert np.allclose(g["cubes10"][l], O.cuboid_transform_from(first[4:14], cams_wc[int(first[0])]), atol=1e-12)
    q = rec[rec[:, 2] == 1][:, 3]
    assert np.allclose(g["ec"][3][:, 0], (2 * q) ** 2) and np.allclose(g["ec"][3][:, 10], (2 * q) ** 2) and np.all(g["ec"][3][:, 1] == 0)
    E = O.ba_edges(ec=g["ec"], ep=None, eo=g["eo"])
    lin = O.ba_linearize(g["cams7"], g["cam_fixed"], g["cubes10"], g["cube_fixed"], E)
    assert np.abs(lin["eo_err"]).max() < 1e-9           # odometry measurements are consistent with the poses
    err = np.linalg.norm(lin["ec_err"], axis=1)
    assert err.max() < 1.0 and err.min() < 1e-9         # 5 cm noise; the initialising observation has zero residual


def test_unseen_landmarks_and_single_frame():
    cams_wc = np.array([[0, 0, 1.6, 0, 0, 0, 1.0]])
    rec = np.zeros((3, 16)); rec[:, 1] = [0, 1, 2]; rec[1, 2] = 1; rec[1, 3] = 0.5; rec[1, 4:14] = [0, 0, 5, 0, 0, 0, 1, 1, 1, 1]
    g = graph.assemble_graph(rec, cams_wc, 3)
    assert g["eo"] is None and len(g["ec"][0]) == 1 and g["cube_fixed"].tolist() == [1, 0, 1]
    assert np.allclose(g["cubes10"][1], [0, 0, 6.6, 0, 0, 0, 1, 1, 1, 1])
