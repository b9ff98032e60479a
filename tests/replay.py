"""Replay of the reference's object_slam node in ONLINE mode (object_slam/src/main_obj.cpp:479-841, online_detect_mode = true) on its bundled TUM
sequence, with the stages supplied by a backend (the CPU oracles, or the GPU library through the C ABI).  TEST INFRASTRUCTURE.

Per frame (main_obj.cpp):  constant-velocity pose prediction (:545-558) -> line_lbd_detect::detect_filter_lines with use_LSD = false and
line_length_thres 15 (:503-505, 596) -> detect_cuboid on the YOLO box with the FIRST frame's pose as transToWolrd, roll / pitch sampling on for
every frame but the first, nominal_skew_ratio 2 (:494, 612-618) -> measurement in the camera frame with the sampled roll / pitch applied
(:643-679) -> graph: camera vertex (first fixed), EdgeSE3Cuboid with information (2 meas_quality)^2, EdgeSE3Expmap odometry (:738-799) ->
optimize(5) (:803) -> the landmark estimate after every frame = one row of output_obj_poses.txt, the final camera poses = output_cam_poses.txt."""
import os

import cv2
import numpy as np

import oracle_lib as O
from cube_slam_wu_b200 import graph, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K_TUM = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]])   # main_obj.cpp:484-486


def load_sequence():
    d = np.load(os.path.join(GOLD, "tum_online.npz"))
    ba = np.load(os.path.join(GOLD, "tum_ba.npz"))
    frames = []
    for f in range(len(d["jpeg_off"]) - 1):
        img = cv2.imdecode(d["jpeg"][d["jpeg_off"][f]:d["jpeg_off"][f + 1]], 1)
        frames.append(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))          # detect_filter_lines / detect_cuboid convert BGR -> gray themselves
    boxes = []
    for f in range(len(frames)):
        b = d["boxes"][d["boxes"][:, 0] == f][:, 1:].copy()
        b[:, :2] -= 1                                                  # "change matlab coordinate to c++" (:608)
        boxes.append(b)
    return frames, boxes, ba["truth"], ba["out_obj"], ba["out_cam"]


def write_base_folder(dst):
    """the committed TUM fixture laid out like the reference's object_slam/data (what its node and cube_slam_wu_b200.node read): the JPEG
    files byte for byte, the YOLO boxes (1-based, tab separated, empty file = no box), truth_cam_poses.txt with its four decimals"""
    d = np.load(os.path.join(GOLD, "tum_online.npz"))
    truth = np.load(os.path.join(GOLD, "tum_ba.npz"))["truth"]
    os.makedirs(os.path.join(dst, "raw_imgs")); os.makedirs(os.path.join(dst, "filter_2d_obj_txts"))
    for f in range(len(d["jpeg_off"]) - 1):
        with open(os.path.join(dst, "raw_imgs", "%04d_rgb_raw.jpg" % f), "wb") as fh:
            fh.write(d["jpeg"][d["jpeg_off"][f]:d["jpeg_off"][f + 1]].tobytes())
        with open(os.path.join(dst, "filter_2d_obj_txts", "%04d_yolo2_0.15.txt" % f), "w") as fh:
            for b in d["boxes"][d["boxes"][:, 0] == f][:, 1:]:
                fh.write("%d\t%d\t%d\t%d\t%.2f\n" % (b[0], b[1], b[2], b[3], b[4]))
    with open(os.path.join(dst, "truth_cam_poses.txt"), "w") as fh:
        for r in truth:
            fh.write("\t".join("%.4f" % v for v in r) + "\t\n")
    return dst


def pose_mat(v7):
    x, y, z, qx, qy, qz, qw = v7
    R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                  [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                  [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = [x, y, z]
    return T


def quat_to_euler_zyx(q):  # matrix_utils.cpp:38-51; q = x y z w
    qx, qy, qz, qw = q
    return (np.arctan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy)), np.arcsin(2 * (qw * qy - qz * qx)),
            np.arctan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz)))


class OracleBackend:
    """every stage from the CPU oracles (distance maps: cv2, the reference's own calls box_proposal_detail.cpp:320-327)"""

    def __init__(self, use_lsd=False, blur_generation=3, roi_view=True):
        self.use_lsd = use_lsd
        self.blur_generation = blur_generation   # 3: the 8-bit Gaussian taps of OpenCV <= 3.4.0, what the reference's author ran (oracle_lbd.cpp)
        self.roi_view = roi_view                 # False: Canny on an isolated copy of the ROI (what cv2 does with a numpy slice)

    def lines(self, gray):
        if self.use_lsd:
            return O.lsd_detect(gray).astype(np.float64)
        O.lbd_set_blur_generation(self.blur_generation)
        try:
            return O.edlines_detect(gray)[0].astype(np.float64)
        finally:
            O.lbd_set_blur_generation(4)

    def best_cuboid(self, gray, T0, box, lines, sample):
        H, W = gray.shape
        tasks = O.plan(box, W, H)
        # cv::Canny on the cv::Mat ROI VIEW gray_img(object_bbox): its Sobel sees the ROI's real neighbours in the parent frame (a python slice
        # would replicate the ROI's own border instead).  With the view semantics the replay matches output_obj_poses.txt to 5e-7, with
        # isolated slices only to 7e-5 -- the reference's OpenCV evidently filtered across the ROI border.
        roi_map = synth.dist_map_for_roi_reference if self.roi_view else synth.dist_map_for_roi
        maps = [roi_map(gray, t.left, t.top, t.width, t.height) for t in tasks]
        P = O.default_params(whether_sample_cam_roll_pitch=int(sample), nominal_skew_ratio=2.0)
        R = O.detect_frame(K_TUM, T0, W, H, box, lines, maps, P)
        b = R.boxes[0]
        if not len(b["sorted"]):
            return None
        c = b["raw"][b["sorted"][0]]
        return dict(pos=np.array(c.pos), rotY=c.rotY, scale=np.array(c.scale), err=c.normalized_error, droll=c.camera_roll_delta, dpitch=c.camera_pitch_delta,
                    rank_index=int(b["sorted"][0]), dist=c.edge_distance_error, angle=c.edge_angle_error)

    def optimize(self, cams, fixed, cube, ec, eo):
        E = O.ba_edges(ec=(ec[0], ec[1], np.array(ec[2]), np.array(ec[3])),
                       eo=(eo[0], eo[1], np.array(eo[2]).reshape(-1, 7), np.array(eo[3]).reshape(-1, 36)) if len(eo[0]) else None)
        c2, q2, _, _ = O.ba_optimize(np.array(cams), fixed, cube.reshape(1, 10), [0], E, 5)
        return c2, q2[0]


class GpuBackend:
    """every stage from libcubeslam_b200.so through the C ABI: csb_edlines_detect_batch (or csb_lsd_detect_batch), csb_detect_batch_gray
    (Canny + distance transform + proposals on the GPU), csb_ba_set_graph + csb_ba_optimize (LM on the device)"""

    def __init__(self, ctx, csb, use_lsd=False, blur_generation=3):
        self.ctx, self.csb, self.use_lsd, self.blur_generation = ctx, csb, use_lsd, blur_generation

    def lines(self, gray):
        if self.use_lsd:
            out, _ = self.ctx.lsd_detect_batch(gray[None], 15.0, True)
        else:
            self.ctx.set_blur_generation(self.blur_generation)   # csb_set_blur_generation: the author's OpenCV generation (see OracleBackend)
            try:
                out, _ = self.ctx.edlines_detect_batch(gray[None], 15.0, True)
            finally:
                self.ctx.set_blur_generation(4)
        return np.ascontiguousarray(out[0].astype(np.float64)).reshape(-1, 4)

    def best_cuboid(self, gray, T0, box, lines, sample):
        csb, ctx = self.csb, self.ctx
        H, W = gray.shape
        params = csb.DetectParams.default(whether_sample_cam_roll_pitch=int(sample), nominal_skew_ratio=2.0)
        frames = csb.make_frames([K_TUM], [T0], W, H, [(0, len(box))], [(0, len(lines))])
        boxes = np.ascontiguousarray(box, np.float64).reshape(-1, 5)
        lines = np.ascontiguousarray(lines, np.float64).reshape(-1, 4) if len(lines) else np.zeros((0, 4))
        tasks, n_tasks, n_map = csb.detect_plan(frames, boxes, params)
        cub, ncub, _ = ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, np.ascontiguousarray(gray.ravel(), np.uint8), params)
        if ncub[0] < 1:
            return None
        c = cub[0]
        return dict(pos=np.array(c.pos[:]), rotY=c.rotY, scale=np.array(c.scale[:]), err=c.normalized_error, droll=c.camera_roll_delta,
                    dpitch=c.camera_pitch_delta, rank_index=int(c.rank_index), dist=c.edge_distance_error, angle=c.edge_angle_error)

    def optimize(self, cams, fixed, cube, ec, eo):
        """The node adds one keyframe per frame and optimises again (main_obj.cpp:738-803): after the first frame the device graph grows by
        csb_ba_add_frame (the frame's camera, its cuboid edge if the box gave one, the odometry edge from the previous frame); the first frame
        -- or anything that is not 'the graph so far plus one frame' -- goes through csb_ba_set_graph."""
        ctx = self.ctx
        n = len(cams)
        state = getattr(self, "_ba_state", None)
        grown = (state is not None and state == (n - 1, sum(1 for c in ec[0] if c < n - 1), sum(1 for j in eo[1] if j < n - 1)) and n >= 2)
        if grown:
            m_ec = [k for k, c in enumerate(ec[0]) if c == n - 1]
            m_eo = [k for k, j in enumerate(eo[1]) if j == n - 1]
            grown = all(k >= state[1] for k in m_ec) and all(k >= state[2] for k in m_eo)
        if grown:
            ecf = (np.array([ec[1][k] for k in m_ec], np.int32), np.array([ec[2][k] for k in m_ec]).reshape(-1, 10), np.array([ec[3][k] for k in m_ec]).reshape(-1, 81)) if m_ec else None
            eof = (np.array([eo[0][k] for k in m_eo], np.int32), np.array([eo[2][k] for k in m_eo]).reshape(-1, 7), np.array([eo[3][k] for k in m_eo]).reshape(-1, 36)) if m_eo else None
            idx = ctx.ba_add_frame(np.array(cams[-1]), cam_fixed=bool(fixed[-1]), ec=ecf, eo=eof)
            assert idx == n - 1
            self.n_add_frame = getattr(self, "n_add_frame", 0) + 1
        else:
            ecs = (np.array(ec[0], np.int32), np.array(ec[1], np.int32), np.array(ec[2]).reshape(-1, 10), np.array(ec[3]).reshape(-1, 81)) if len(ec[0]) else None
            eos = (np.array(eo[0], np.int32), np.array(eo[1], np.int32), np.array(eo[2]).reshape(-1, 7), np.array(eo[3]).reshape(-1, 36)) if len(eo[0]) else None
            ctx.ba_set_graph(np.array(fixed, np.int32), np.zeros(1, np.int32), ec=ecs, ep=None, eo=eos)
        self._ba_state = (n, len(ec[0]), len(eo[0]))
        # the caller's estimates (in the checked replay they are the ORACLE's, so that every frame compares like with like)
        ctx.ba_upload_estimates(np.array(cams), cube.reshape(1, 10))
        c2, q2, _ = ctx.ba_optimize(5)
        return c2, q2[0]


class CheckedBackend:
    """The oracle backend drives the replay; every stage is ALSO run on the GPU backend with the oracle's inputs and compared, so the first
    stage that differs is named with its frame: line tables bit for bit, the best cuboid's ranking index exactly and its floats to 1e-9, the
    estimates after optimize(5) to `lm_tol` (the LM drift of one frame, recorded in self.lm_drift)."""

    def __init__(self, oracle_backend, gpu_backend, lm_tol=1e-5):
        self.o, self.g, self.lm_tol = oracle_backend, gpu_backend, lm_tol
        self.frame, self.lm_drift, self.n_cuboids = 0, [], 0

    def lines(self, gray):
        a, b = self.o.lines(gray), self.g.lines(gray)
        assert a.shape == b.shape and np.array_equal(a, b), "frame %d: line tables differ (%s vs %s)" % (self.frame, a.shape, b.shape)
        return a

    def best_cuboid(self, gray, T0, box, lines, sample):
        a, b = self.o.best_cuboid(gray, T0, box, lines, sample), self.g.best_cuboid(gray, T0, box, lines, sample)
        assert (a is None) == (b is None), "frame %d: only one side found a cuboid" % self.frame
        if a is not None:
            assert a["rank_index"] == b["rank_index"], "frame %d: best proposal index %d (oracle) vs %d (GPU)" % (self.frame, a["rank_index"], b["rank_index"])
            for k in ("pos", "rotY", "scale", "err", "droll", "dpitch", "dist", "angle"):
                d = float(np.abs(np.asarray(a[k]) - np.asarray(b[k])).max())
                assert d <= 1e-9, "frame %d: cuboid field %s differs by %g" % (self.frame, k, d)
            self.n_cuboids += 1
        return a

    def optimize(self, cams, fixed, cube, ec, eo):
        c2, q2 = self.o.optimize(cams, fixed, cube, ec, eo)
        g2, gq = self.g.optimize(cams, fixed, cube, ec, eo)
        d = max(float(np.abs(np.asarray(g2)[:len(cams)] - np.asarray(c2)[:len(cams)]).max()), float(np.abs(gq - q2).max()))
        assert d <= self.lm_tol, "frame %d: estimates after optimize(5) differ by %g" % (self.frame, d)
        self.lm_drift.append(d)
        self.frame += 1
        return c2, q2


def run(backend, frames, boxes, truth, n_frames=None):
    cv2.setNumThreads(1)
    try:
        cv2.ipp.setUseIPP(False)
    except Exception:
        pass
    N = len(frames) if n_frames is None else n_frames
    ident = np.array([0, 0, 0, 0, 0, 0, 1.0])
    Twc0 = O.se3_mul(truth[0, 1:8], ident)            # g2o::SE3Quat(Vector7d) keeps w >= 0
    T0 = pose_mat(Twc0)
    eul0 = quat_to_euler_zyx(Twc0[3:7])               # detect_cuboid_obj.cam_pose_raw.euler_angle after set_cam_pose(transToWolrd)
    cams, fixed, ec, eo, cube, hist, n_lines = [], [], ([], [], [], []), ([], [], [], []), None, [], []
    for f in range(N):
        odom = ident.copy()
        if f == 0:
            Twc = Twc0
        else:
            prev = cams[f - 1]
            if f > 1:
                odom = O.se3_mul(prev, O.se3_inverse(cams[f - 2]))
            Twc = O.se3_inverse(O.se3_mul(odom, prev))
        gray, box = frames[f], boxes[f]
        lines = backend.lines(gray)
        n_lines.append(len(lines))
        sample = f != 0
        best = backend.best_cuboid(gray, T0, box, lines, sample) if len(box) else None   # transToWolrd is the first frame's pose either way
        if best is not None:
            cg = O.cuboid_from_minimal([best["pos"][0], best["pos"][1], best["pos"][2], 0, 0, best["rotY"], best["scale"][0], best["scale"][1], best["scale"][2]])
            meas = O.cuboid_transform_to(cg, Twc)
            if sample:
                Rn = np.asarray(synth.euler_zyx_to_rot(eul0[0] + best["droll"], eul0[1] + best["dpitch"], eul0[2]))
                Tn = np.eye(4); Tn[:3, :3] = Rn; Tn[:3, 3] = T0[:3, 3]
                meas = O.cuboid_transform_to(cg, graph.pose7_from_matrix(Tn))
            q = (1 - best["err"] + 0.5) / 2
        if f == 0:
            cube = O.cuboid_transform_from(meas, Twc)
        cams.append(O.se3_inverse(Twc)); fixed.append(1 if f == 0 else 0)
        if best is not None:
            ec[0].append(f); ec[1].append(0); ec[2].append(meas); ec[3].append((np.eye(9) * (2 * q) ** 2).ravel())
        if f > 0:
            eo[0].append(f - 1); eo[1].append(f); eo[2].append(odom); eo[3].append(np.eye(6).ravel())
        c2, cube = backend.optimize(cams, fixed, cube, ec, eo)
        cams = [c2[i] for i in range(f + 1)]
        hist.append(cube.copy())
    hist = np.array(hist)
    obj9 = np.concatenate([hist[:, :3], np.zeros((len(hist), 2)), 2 * np.arctan2(hist[:, 5:6], hist[:, 6:7]), hist[:, 7:10]], axis=1)  # x y z . . yaw scale
    Twc = np.array([O.se3_inverse(c) for c in cams])
    return dict(obj=obj9, cube10=hist, Twc=Twc, n_lines=np.array(n_lines))
