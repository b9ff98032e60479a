"""object_slam's node in ONLINE mode on the B200 library: the host side of the reference's `incremental_build_graph`
(object_slam/src/main_obj.cpp:479-841, online_detect_mode = true) with every compute stage behind the C ABI.

    python -m cube_slam_wu_b200.node --base-folder <object_slam/data> [--out <dir>] [--lsd] [--blur-generation 3] [--offline]

reads the reference's data folder as its node does (raw_imgs/%04d_rgb_raw.jpg, filter_2d_obj_txts/%04d_yolo2_0.15.txt,
truth_cam_poses.txt: main_obj.cpp:879-893, 585-620) and writes output_cam_poses.txt / output_obj_poses.txt in the reference's formats
(main_obj.cpp:305-336).  Per frame (one landmark, perfect association, as in the reference):

  constant-velocity pose prediction (:545-564)
  -> csb_edlines_detect_batch / csb_lsd_detect_batch   = line_lbd_detect::detect_filter_lines, line_length_thres 15 (:503-505, 596)
  -> csb_detect_batch_gray                             = detect_3d_cuboid::detect_cuboid with the FIRST frame's pose as transToWolrd,
                                                         roll / pitch sampling for every frame but the first, skew ratio 2 (:494, 612-618)
  -> the measurement in the camera frame, the sampled roll / pitch applied (:643-679), meas_quality (:732)
  -> csb_ba_set_graph (first frame) / csb_ba_add_frame  = camera vertex (first one fixed), EdgeSE3Cuboid with information
                                                         (2 meas_quality)^2, EdgeSE3Expmap odometry with identity information (:738-799)
  -> csb_ba_optimize(5)                                 = graph.optimize(5) (:803)

Only bookkeeping happens here (SE(3) products of a handful of poses, the graph's index lists); nothing on this path computes with the CPU
oracle -- the tests drive this class with the real Context on the GPU (tests/test_node_gpu.py) and with a stand-in on the CPU
(tests/test_node.py)."""
import argparse
import os

import numpy as np

from . import graph, synth

K_TUM = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]])   # main_obj.cpp:484-486
_IDENT = np.array([0, 0, 0, 0, 0, 0, 1.0])


def se3_from_vector7(v):
    """g2o::SE3Quat(Vector7d): x y z qx qy qz qw, rotation normalised with w >= 0 (se3quat.h:58-70, 88-93)"""
    v = np.asarray(v, np.float64)
    q = v[3:7] / np.linalg.norm(v[3:7])
    return np.concatenate([v[:3], -q if q[3] < 0 else q])


def pose_matrix(p7):
    x, y, z, qx, qy, qz, qw = p7
    T = np.eye(4)
    T[:3, :3] = [[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                 [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                 [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]]
    T[:3, 3] = [x, y, z]
    return T


def quat_to_euler_zyx(q):
    """roll, pitch, yaw of x y z w (matrix_utils.cpp:38-51; se3quat.h:184-194)"""
    qx, qy, qz, qw = q
    return (np.arctan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy)), np.arcsin(2 * (qw * qy - qz * qx)),
            np.arctan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz)))


def cuboid_from_minimal(v9):
    """g2o::cuboid::fromMinimalVector (g2o_Object.h:37-42): x y z roll pitch yaw + half scales -> x y z qx qy qz qw + half scales"""
    v9 = np.asarray(v9, np.float64)
    return np.concatenate([se3_from_vector7(np.concatenate([v9[:3], synth.quat_from_euler(v9[3], v9[4], v9[5])])), v9[6:9]])


def cuboid_to_minimal(c10):
    """g2o::cuboid::toMinimalVector (g2o_Object.h:136-142)"""
    return np.concatenate([c10[:3], quat_to_euler_zyx(c10[3:7]), c10[7:10]])


def cuboid_transform_to(c10, Twc7):
    """g2o::cuboid::transform_to (g2o_Object.h:126-132): pose = Twc^-1 * pose"""
    return np.concatenate([graph.se3_mul(graph.se3_inv(Twc7), c10[:7]), c10[7:10]])


def cuboid_transform_from(c10, Twc7):
    """g2o::cuboid::transform_from (g2o_Object.h:117-122): pose = Twc * pose"""
    return np.concatenate([graph.se3_mul(Twc7, c10[:7]), c10[7:10]])


class ObjectSlamNode:
    """The state `incremental_build_graph` keeps between frames: camera estimates (world -> camera), the landmark, the device graph."""

    def __init__(self, ctx, csb, first_cam_pose_Twc, K=K_TUM, use_lsd=False, line_length_thres=15.0, nominal_skew_ratio=2.0, lm_iterations=5,
                 blur_generation=4):
        self.ctx, self.csb, self.K = ctx, csb, np.asarray(K, np.float64)
        self.use_lsd, self.line_length_thres, self.skew, self.lm_iterations = bool(use_lsd), float(line_length_thres), float(nominal_skew_ratio), int(lm_iterations)
        self.blur_generation = int(blur_generation)   # csb_set_blur_generation: 3 = the 8-bit Gaussian of OpenCV <= 3.4.0, the author's build
        self.Twc0 = se3_from_vector7(first_cam_pose_Twc)   # fixed_init_cam_pose_Twc (:528): only the first truth pose is used
        self.T0 = pose_matrix(self.Twc0)
        self.eul0 = quat_to_euler_zyx(self.Twc0[3:7])     # cam_pose_raw.euler_angle after set_cam_pose(transToWolrd)
        self.detector = csb.detect_3d_cuboid(ctx)          # detect_cuboid_obj (:494-500)
        self.detector.set_calibration(self.K)
        self.detector.whether_sample_bbox_height = False
        self.detector.nominal_skew_ratio = self.skew
        self.cams = []            # optimised world -> camera poses (VertexSE3Expmap estimates), one per frame
        self.cube = None          # the landmark (VertexCuboid estimate), 10 doubles
        self.history = []         # the landmark after every frame's optimisation (cube_pose_opti_history)
        self.n_lines = []
        self.n_cuboid_edges = 0

    # -- stages behind the C ABI ---------------------------------------------------------------------------------------------------
    def detect_lines(self, gray):
        if self.use_lsd:
            out, _ = self.ctx.lsd_detect_batch(gray[None], self.line_length_thres, True)
        else:
            self.ctx.set_blur_generation(self.blur_generation)
            try:
                out, _ = self.ctx.edlines_detect_batch(gray[None], self.line_length_thres, True)
            finally:
                self.ctx.set_blur_generation(4)
        return np.ascontiguousarray(np.asarray(out[0], np.float64)).reshape(-1, 4)

    def detect_cuboid(self, gray, boxes, lines, sample_roll_pitch):
        """the best cuboid of the frame's first 2D box (frames_cuboids[0][0]; the reference's frames carry at most one box), or None"""
        d = self.detector
        d.whether_sample_cam_roll_pitch = bool(sample_roll_pitch)                    # (:624)
        found = d.detect_cuboid(gray, self.T0, boxes, lines)
        return found[0][0] if len(found) and len(found[0]) else None

    # -- one frame -----------------------------------------------------------------------------------------------------------------
    def _predict(self):
        """constant-velocity prediction of the new frame's camera (:545-564): (Twc, odometry measurement)"""
        f = len(self.cams)
        odom = _IDENT.copy()
        if f == 0:
            return self.Twc0, odom
        prev = self.cams[f - 1]
        if f > 1:
            odom = graph.se3_mul(prev, graph.se3_inv(self.cams[f - 2]))
        return graph.se3_inv(graph.se3_mul(odom, prev)), odom

    def _grow_and_optimize(self, Twc, odom, meas, proposal_error):
        """vertices / edges of the new frame (:738-799) and graph.optimize (:803); meas = the cuboid in the camera frame or None"""
        f = len(self.cams)
        ec = None
        if meas is not None:
            quality = (1 - proposal_error + 0.5) / 2                                 # meas_quality (:732)
            ec = (np.zeros(1, np.int32), meas.reshape(1, 10), ((2 * quality) ** 2 * np.eye(9)).reshape(1, 81))
        cam = graph.se3_inv(Twc)
        if f == 0:
            if meas is None:
                raise RuntimeError("object_slam node: the first frame has no cuboid -- the reference initialises its landmark from it (main_obj.cpp:745-751)")
            self.cube = cuboid_transform_from(meas, Twc)
            self.ctx.ba_set_graph(np.ones(1, np.int32), np.zeros(1, np.int32), ec=(np.zeros(1, np.int32), ec[0], ec[1], ec[2]), ep=None, eo=None)
        else:
            eo = (np.array([f - 1], np.int32), odom.reshape(1, 7), np.eye(6).reshape(1, 36))
            idx = self.ctx.ba_add_frame(cam, cam_fixed=False, ec=ec, eo=eo)
            if idx != f:
                raise RuntimeError("csb_ba_add_frame returned camera %d for frame %d" % (idx, f))
        self.n_cuboid_edges += 0 if ec is None else 1
        self.ctx.ba_upload_estimates(np.array(self.cams + [cam]), self.cube.reshape(1, 10))
        cams, cubes, _ = self.ctx.ba_optimize(self.lm_iterations)
        self.cams = [np.array(cams[i]) for i in range(f + 1)]
        self.cube = np.array(cubes[0])
        self.history.append(self.cube.copy())
        return self.cube

    def add_frame(self, gray, boxes):
        """online_detect_mode = true.  gray: (h, w) uint8; boxes: (k, 5) x y w h prob, 0-based.  Returns the landmark estimate after this
        frame's optimisation."""
        Twc, odom = self._predict()
        lines = self.detect_lines(gray)
        self.n_lines.append(len(lines))
        sample = len(self.cams) != 0
        best = self.detect_cuboid(gray, boxes, lines, sample) if len(boxes) else None
        if best is None:
            return self._grow_and_optimize(Twc, odom, None, 0.0)
        cube_ground = cuboid_from_minimal([best.pos[0], best.pos[1], best.pos[2], 0, 0, best.rotY, best.scale[0], best.scale[1], best.scale[2]])
        meas = cuboid_transform_to(cube_ground, Twc)
        if sample:   # the detector's own camera: the first pose with the sampled roll / pitch (:655-672)
            Tn = np.eye(4)
            Tn[:3, :3] = np.asarray(synth.euler_zyx_to_rot(self.eul0[0] + best.camera_roll_delta, self.eul0[1] + best.camera_pitch_delta, self.eul0[2]))
            Tn[:3, 3] = self.T0[:3, 3]
            meas = cuboid_transform_to(cube_ground, graph.pose7_from_matrix(Tn))
        return self._grow_and_optimize(Twc, odom, meas, best.normalized_error)

    def add_frame_offline(self, saved_cuboid, init_cam_pose_Twc):
        """online_detect_mode = false (:686-712): the frame's cuboid comes from detect_cuboids_saved.txt (x y z yaw sx sy sz error, in the
        ground frame of the saved camera pose pop_cam_poses_saved.txt) or is None; only the graph and the optimiser run."""
        Twc, odom = self._predict()
        if saved_cuboid is None:
            return self._grow_and_optimize(Twc, odom, None, 0.0)
        m = np.asarray(saved_cuboid, np.float64)
        cube_ground = cuboid_from_minimal([m[0], m[1], m[2], 0, 0, m[3], m[4], m[5], m[6]])
        return self._grow_and_optimize(Twc, odom, cuboid_transform_to(cube_ground, se3_from_vector7(init_cam_pose_Twc)), m[7])

    # -- results -------------------------------------------------------------------------------------------------------------------
    def cam_poses_Twc(self):
        return np.array([graph.se3_inv(c) for c in self.cams])

    def object_history_minimal(self):
        return np.array([cuboid_to_minimal(c) for c in self.history])


# ---- the reference's files -------------------------------------------------------------------------------------------------------
def read_base_folder(base_folder):
    """frames (gray), per-frame boxes (0-based x y w h prob) and the truth poses of object_slam/data (main_obj.cpp:585-620, 879-893)"""
    import cv2
    truth = np.loadtxt(os.path.join(base_folder, "truth_cam_poses.txt")).reshape(-1, 8)
    frames, boxes = [], []
    for f in range(truth.shape[0]):
        img = cv2.imread(os.path.join(base_folder, "raw_imgs", "%04d_rgb_raw.jpg" % f), 1)
        if img is None:
            raise FileNotFoundError("raw_imgs/%04d_rgb_raw.jpg" % f)
        frames.append(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))   # detect_filter_lines / detect_cuboid convert BGR -> gray themselves
        p = os.path.join(base_folder, "filter_2d_obj_txts", "%04d_yolo2_0.15.txt" % f)
        b = np.loadtxt(p).reshape(-1, 5) if os.path.exists(p) and os.path.getsize(p) > 0 else np.zeros((0, 5))
        b = b.copy()
        b[:, :2] -= 1                                          # "change matlab coordinate to c++" (:620)
        boxes.append(b)
    return frames, boxes, truth


def read_offline_tables(base_folder):
    """detect_cuboids_saved.txt (frame x y z yaw sx sy sz error, at most one row per frame), pop_cam_poses_saved.txt and truth_cam_poses.txt
    (time x y z qx qy qz qw): the inputs of the reference's offline mode (main_obj.cpp:879-893)"""
    rd = lambda n, c: np.loadtxt(os.path.join(base_folder, n)).reshape(-1, c)
    return rd("detect_cuboids_saved.txt", 9), rd("pop_cam_poses_saved.txt", 8), rd("truth_cam_poses.txt", 8)


def run_offline(ctx, csb, det, pop, truth, n_frames=None, **node_args):
    node = ObjectSlamNode(ctx, csb, truth[0, 1:8], **node_args)
    row = 0
    for f in range(len(truth) if n_frames is None else n_frames):
        has = row < len(det) and int(det[row, 0]) == f          # "not all frame has observation" (:689-690)
        node.add_frame_offline(det[row, 1:9] if has else None, pop[f, 1:8])
        row += int(has)
    return node


def _eigen_row(values):
    """Eigen's default operator<< of a row vector: 6 significant digits, every coefficient right-aligned to the widest, one space between"""
    s = ["%.6g" % v for v in values]
    w = max(len(x) for x in s)
    return " ".join(x.rjust(w) for x in s)


def _ros_time(t):
    """operator<<(ros::Time) of ros::Time(double): seconds '.' nine digits of nanoseconds"""
    sec = int(np.floor(t))
    nsec = int(round((t - sec) * 1e9))
    if nsec >= 1000000000:
        sec, nsec = sec + 1, nsec - 1000000000
    return "%d.%09d" % (sec, nsec)


def write_results(out_folder, timestamps, cam_poses_Twc, object_history_minimal):
    """output_cam_poses.txt / output_obj_poses.txt as the reference's node writes them (main_obj.cpp:305-336)"""
    os.makedirs(out_folder, exist_ok=True)
    with open(os.path.join(out_folder, "output_cam_poses.txt"), "w") as f:
        f.write("# timestamp tx ty tz qx qy qz qw\n")
        for t, p in zip(timestamps, cam_poses_Twc):
            f.write(_ros_time(t) + "  " + _eigen_row(p) + "\n")
    with open(os.path.join(out_folder, "output_obj_poses.txt"), "w") as f:
        for v in object_history_minimal:
            f.write(_eigen_row(v) + " \n")


def run_sequence(ctx, csb, frames, boxes, truth, n_frames=None, **node_args):
    node = ObjectSlamNode(ctx, csb, truth[0, 1:8], **node_args)
    for f in range(len(frames) if n_frames is None else n_frames):
        node.add_frame(frames[f], boxes[f])
    return node


def main(argv=None):
    import cube_slam_wu_b200 as csb
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--base-folder", required=True, help="the reference's object_slam/data layout")
    ap.add_argument("--out", default=None, help="where output_cam_poses.txt / output_obj_poses.txt go (default: the base folder, like the reference)")
    ap.add_argument("--lsd", action="store_true", help="line_lbd_obj.use_LSD = true (the reference's node runs EDLines)")
    ap.add_argument("--blur-generation", type=int, default=4, choices=[3, 4])
    ap.add_argument("--offline", action="store_true", help="online_detect_mode = false: cuboids from detect_cuboids_saved.txt, poses from pop_cam_poses_saved.txt")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    ctx = csb.Context(a.device)   # raises if the CUDA library is not built or no device is usable: there is no CPU path
    if a.offline:
        det, pop, truth = read_offline_tables(a.base_folder)
        node = run_offline(ctx, csb, det, pop, truth)
    else:
        frames, boxes, truth = read_base_folder(a.base_folder)
        node = run_sequence(ctx, csb, frames, boxes, truth, use_lsd=a.lsd, blur_generation=a.blur_generation)
    write_results(a.out or a.base_folder, truth[:, 0], node.cam_poses_Twc(), node.object_history_minimal())
    print("%d frames, %d cuboid edges, landmark %s" % (len(node.cams), node.n_cuboid_edges, _eigen_row(node.object_history_minimal()[-1])))
    return node


if __name__ == "__main__":
    main()
