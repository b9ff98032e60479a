"""cube_slam_wu_b200.node -- the host side of the reference's object_slam node in online mode (object_slam/src/main_obj.cpp:479-841) -- on the
CPU: the class driven through a stand-in context that computes with the oracles (tests/test_reference_replay.py::_FakeCtx), so what is
checked here is the node's own bookkeeping (pose prediction, the measurement recipe, the graph growth, the reference's txt formats) against
the reference's committed output files.  The same class with the real Context on the B200: tests/test_node_gpu.py."""
import os

import numpy as np
import pytest

import replay
from test_reference_replay import _FakeCtx
from cube_slam_wu_b200 import node


@pytest.fixture(scope="module")
def seq():
    return replay.load_sequence()


def test_pose_helpers_match_the_oracle():
    O = replay.O
    rng = np.random.default_rng(5)
    for _ in range(20):
        v = np.concatenate([rng.normal(size=3), rng.normal(size=4)])
        p = node.se3_from_vector7(v)
        assert np.allclose(p, O.se3_mul(v, np.array([0, 0, 0, 0, 0, 0, 1.0])), atol=1e-15)
        m9 = np.concatenate([rng.normal(size=3), rng.uniform(-1.5, 1.5, 3), rng.uniform(0.1, 2, 3)])
        c = node.cuboid_from_minimal(m9)
        assert np.allclose(c, O.cuboid_from_minimal(m9), atol=1e-15)
        assert np.allclose(node.cuboid_transform_to(c, p), O.cuboid_transform_to(c, p), atol=1e-14)
        assert np.allclose(node.cuboid_transform_from(c, p), O.cuboid_transform_from(c, p), atol=1e-14)
        assert np.allclose(node.cuboid_to_minimal(c), m9, atol=1e-12)


def test_node_reproduces_the_reference_output_files(seq, csb, tmp_path):
    """All 58 frames: the landmark after every frame and the camera track at the files' printed precision (the bounds of
    tests/test_reference_replay.py), one csb_ba_set_graph and 57 csb_ba_add_frame; then the files themselves, format included."""
    frames, boxes, truth, out_obj, out_cam = seq
    fake = _FakeCtx(csb)
    nd = node.run_sequence(fake, csb, frames, boxes, truth, blur_generation=3)
    assert fake.calls[:7] == ["blur3", "edlines", "blur4", "detect", "set_graph", "upload", "optimize"]
    assert fake.calls.count("set_graph") == 1 and fake.calls.count("add_frame") == 57
    assert nd.n_cuboid_edges == sum(1 for b in boxes if len(b)) == fake.calls.count("detect")
    obj = nd.object_history_minimal()
    dpos = np.linalg.norm(obj[:, :3] - out_obj[:, :3], axis=1)
    dyaw = np.abs(np.angle(np.exp(1j * (obj[:, 5] - out_obj[:, 5]))))
    dscale = np.abs(obj[:, 6:9] - out_obj[:, 6:9]).max(axis=1)
    assert dpos.max() < 1e-5 and dyaw.max() < 1e-4 and dscale.max() < 2e-6, (dpos.max(), dyaw.max(), dscale.max())
    Twc = nd.cam_poses_Twc()
    assert np.linalg.norm(Twc[:, :3] - out_cam[:, 1:4], axis=1).max() < 2e-5
    dq = np.minimum(np.abs(Twc[:, 3:7] - out_cam[:, 4:8]).max(axis=1), np.abs(Twc[:, 3:7] + out_cam[:, 4:8]).max(axis=1))
    assert dq.max() < 2e-6
    # the oracle replay (tests/replay.py) is the same recipe written against the oracle's SE(3) functions: the two chains differ by the last
    # bits of a few pose products, which the delta = 1e-9 numeric Jacobians of 58 chained LM solves carry to the 1e-7 level (observed 3e-7)
    ref = replay.run(replay.OracleBackend(), frames, boxes, truth)
    assert np.abs(np.array(nd.history) - ref["cube10"]).max() < 1e-5 and np.abs(Twc - ref["Twc"]).max() < 1e-5
    # files: the camera file parses back to the track (timestamps exactly as ros::Time prints them); the object file has the reference's
    # nine columns -- position, yaw and scale within the printed digits of the committed one
    node.write_results(str(tmp_path), truth[:, 0], Twc, obj)
    lines = open(os.path.join(tmp_path, "output_cam_poses.txt")).read().splitlines()
    assert lines[0] == "# timestamp tx ty tz qx qy qz qw" and len(lines) == 59
    assert lines[1].split()[0] == "1341841278.842700005"
    got = np.array([[float(x) for x in l.split()] for l in lines[1:]])
    assert np.abs(got[:, 1:] - out_cam[:, 1:8]).max() < 2e-5 and np.abs(got[:, 0] - truth[:, 0]).max() < 1e-6
    rows = open(os.path.join(tmp_path, "output_obj_poses.txt")).read().splitlines()
    assert len(rows) == 58 and all(len(r.split()) == 9 and r.endswith(" ") for r in rows)
    got = np.array([[float(x) for x in r.split()] for r in rows])
    assert np.abs(got[:, [0, 1, 2, 6, 7, 8]] - out_obj[:, [0, 1, 2, 6, 7, 8]]).max() < 2.5e-5   # both sides rounded to six digits


def test_base_folder_reader(seq, tmp_path):
    """read_base_folder on the reference's folder layout (rebuilt from the committed fixture) gives what the node's tests run on"""
    frames, boxes, truth, _, _ = seq
    f2, b2, t2 = node.read_base_folder(replay.write_base_folder(str(tmp_path / "data")))
    assert len(f2) == 58 and all(np.array_equal(a, b) for a, b in zip(frames, f2))
    assert all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(boxes, b2)) and sum(len(b) == 0 for b in b2) == 7
    assert np.array_equal(truth, t2)


def test_txt_formats_are_the_references_own(tmp_path):
    """Eigen's operator<< (6 significant digits, columns padded to the widest coefficient) and ros::Time's: the reference's first rows,
    re-written from their parsed values, come back byte for byte."""
    row_obj = "    -1.42494      0.41679     0.259904  5.69518e-65 -2.77634e-64      3.06894     0.431456     0.328831     0.259904 "
    row_cam = "1341841278.842700005    -2.5508    0.9872    1.1019 -0.487105  0.767307 -0.351903  0.223902"
    node.write_results(str(tmp_path), [float(row_cam.split()[0])], [[float(x) for x in row_cam.split()[1:]]], [[float(x) for x in row_obj.split()]])
    assert open(os.path.join(tmp_path, "output_obj_poses.txt")).read() == row_obj + "\n"
    assert open(os.path.join(tmp_path, "output_cam_poses.txt")).read().splitlines()[1] == row_cam


def test_first_frame_without_a_cuboid_is_an_error(seq, csb):
    frames, boxes, truth, _, _ = seq
    nd = node.ObjectSlamNode(_FakeCtx(csb), csb, truth[0, 1:8])
    blank = np.full_like(frames[0], 128)
    with pytest.raises(RuntimeError):
        nd.add_frame(blank, np.zeros((0, 5)))


def test_offline_mode_matches_the_oracle_replay(csb, tmp_path):
    """online_detect_mode = false (main_obj.cpp:686-712): saved cuboids + saved camera poses -> the same graph growth and optimiser calls.
    Against the oracle's replay of that loop (tests/test_ba_gpu.py::tum_graph) on the reference's own tables; the tables also go through the
    txt reader."""
    import test_ba_gpu
    d = np.load(os.path.join(replay.GOLD, "tum_ba.npz"))
    for name, key, fmt in (("detect_cuboids_saved.txt", "det", "%.10g"), ("pop_cam_poses_saved.txt", "pop", "%.10f"), ("truth_cam_poses.txt", "truth", "%.4f")):
        np.savetxt(os.path.join(tmp_path, name), d[key], fmt=fmt)
    det, pop, truth = node.read_offline_tables(str(tmp_path))
    assert np.allclose(det, d["det"], rtol=1e-9) and np.allclose(pop, d["pop"], atol=1e-9) and np.array_equal(truth, d["truth"])
    fake = _FakeCtx(csb)
    nd = node.run_offline(fake, csb, d["det"], d["pop"], d["truth"])
    assert fake.calls.count("set_graph") == 1 and fake.calls.count("add_frame") == 57 and "detect" not in fake.calls and "edlines" not in fake.calls
    assert nd.n_cuboid_edges == len(d["det"])
    ref = test_ba_gpu.tum_graph(d)
    assert np.abs(np.array(nd.cams) - ref["cams7"]).max() < 1e-5 and np.abs(nd.cube - ref["cubes10"][0]).max() < 1e-5


def test_detect_3d_cuboid_mirror(seq, csb):
    """csb.detect_3d_cuboid: the reference class's member names and defaults (detect_3d_cuboid.h:74-118); a BGR frame is converted like
    cv2.cvtColor (bit-identical); the call packs what csb_detect_batch_gray expects (checked by the stand-in context's assertions) and
    returns one list per box."""
    import cv2
    frames, boxes, truth, _, _ = seq
    det = csb.detect_3d_cuboid(_FakeCtx(csb))
    assert (det.consider_config_1, det.consider_config_2, det.whether_sample_cam_roll_pitch, det.whether_sample_bbox_height, det.max_cuboid_num,
            det.nominal_skew_ratio, det.max_cut_skew) == (True, True, True, False, 1, 1.0, 3.0)
    with pytest.raises(csb.CsbError):
        det.detect_cuboid(frames[0], np.eye(4), boxes[0], np.zeros((0, 4)))          # no calibration yet
    det.set_calibration(node.K_TUM)
    det.whether_sample_cam_roll_pitch = False
    det.nominal_skew_ratio = 2
    d = np.load(os.path.join(replay.GOLD, "tum_online.npz"))
    bgr = cv2.imdecode(d["jpeg"][d["jpeg_off"][0]:d["jpeg_off"][1]], 1)
    T0 = node.pose_matrix(node.se3_from_vector7(truth[0, 1:8]))
    lines = np.asarray(det._ctx.edlines_detect_batch(frames[0][None], 15.0, True)[0][0], np.float64)
    from_bgr = det.detect_cuboid(bgr, T0, boxes[0], lines)
    from_gray = det.detect_cuboid(frames[0], T0, boxes[0], lines)
    assert len(from_bgr) == len(boxes[0]) == 1 and len(from_bgr[0]) == 1
    assert from_bgr[0][0].rank_index == from_gray[0][0].rank_index and np.array_equal(np.array(from_bgr[0][0].pos), np.array(from_gray[0][0].pos))
    assert det.detect_cuboid(frames[0], T0, np.zeros((0, 5)), lines) == []
    with pytest.raises(csb.CsbError):
        det.detect_cuboid(frames[0].astype(np.float32), T0, boxes[0], lines)
