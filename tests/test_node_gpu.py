"""cube_slam_wu_b200.node on the B200: `python -m cube_slam_wu_b200.node --base-folder <object_slam/data>` on the reference's bundled TUM
sequence (the folder rebuilt from the committed fixture), every stage through the C ABI, against the output files the reference's own node
committed (object_slam/data/output_obj_poses.txt, output_cam_poses.txt; main_obj.cpp:305-336)."""
import os

import numpy as np
import pytest

import replay
from cube_slam_wu_b200 import node

pytestmark = pytest.mark.gpu


def test_node_cli_reproduces_the_reference_output_files(tmp_path):
    _, _, truth, out_obj, out_cam = replay.load_sequence()
    base = replay.write_base_folder(str(tmp_path / "data"))
    out = str(tmp_path / "out")
    nd = node.main(["--base-folder", base, "--out", out, "--blur-generation", "3"])   # the author's OpenCV generation
    assert len(nd.cams) == 58 and nd.n_cuboid_edges == 51
    # in memory: the bounds tests/test_replay_gpu.py and tests/test_reference_replay.py hold the replays to (the files' printed precision)
    obj = nd.object_history_minimal()
    dpos = np.linalg.norm(obj[:, :3] - out_obj[:, :3], axis=1)
    dyaw = np.abs(np.angle(np.exp(1j * (obj[:, 5] - out_obj[:, 5]))))
    dscale = np.abs(obj[:, 6:9] - out_obj[:, 6:9]).max(axis=1)
    assert dpos.max() < 1e-5 and dyaw.max() < 1e-4 and dscale.max() < 2e-6, (dpos.max(), dyaw.max(), dscale.max())
    Twc = nd.cam_poses_Twc()
    dcam = np.linalg.norm(Twc[:, :3] - out_cam[:, 1:4], axis=1)
    dq = np.minimum(np.abs(Twc[:, 3:7] - out_cam[:, 4:8]).max(axis=1), np.abs(Twc[:, 3:7] + out_cam[:, 4:8]).max(axis=1))
    assert dcam.max() < 2e-5 and dq.max() < 2e-6, (dcam.max(), dq.max())
    # the files: same layout as the reference's, values within one unit of the last printed digit of the committed ones
    cam_lines = open(os.path.join(out, "output_cam_poses.txt")).read().splitlines()
    assert cam_lines[0] == "# timestamp tx ty tz qx qy qz qw" and len(cam_lines) == 59 and cam_lines[1].split()[0] == "1341841278.842700005"
    got = np.array([[float(x) for x in l.split()] for l in cam_lines[1:]])
    sgn = np.sign((got[:, 4:8] * out_cam[:, 4:8]).sum(axis=1))[:, None]
    assert np.abs(got[:, 1:4] - out_cam[:, 1:4]).max() < 2.5e-5 and np.abs(sgn * got[:, 4:8] - out_cam[:, 4:8]).max() < 2.5e-6
    rows = open(os.path.join(out, "output_obj_poses.txt")).read().splitlines()
    assert len(rows) == 58 and all(len(r.split()) == 9 for r in rows)
    got = np.array([[float(x) for x in r.split()] for r in rows])
    assert np.abs(got[:, [0, 1, 2, 6, 7, 8]] - out_obj[:, [0, 1, 2, 6, 7, 8]]).max() < 2.5e-5
    print("node on the GPU vs output_obj_poses.txt: pos %.1e yaw %.1e scale %.1e; vs output_cam_poses.txt: %.1e m" % (dpos.max(), dyaw.max(), dscale.max(), dcam.max()))
