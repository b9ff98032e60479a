"""Seeded synthetic inputs for the benchmark and the parity tests.

* KITTI-shaped proposal batches (config #2 / #5): 1242x375 frames, 8 boxes per frame, ~280 line segments, distance maps
  produced the way the reference produces them (cv2.Canny(80,200) + cv2.distanceTransform(DIST_L2, 3) per ROI,
  detect_3d_cuboid/src/box_proposal_detail.cpp:320-327) from a rendered line image.
* camera-cuboid BA graphs (config #4): keyframes on a loop, cuboids on the ground, noisy relative measurements with
  90-degree yaw flips (exercises cuboid::min_log_error, g2o_Object.h:76-101).

Nothing here is on the measured path: generation happens before the timed region.
"""
import math

import numpy as np

KITTI_K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.22], [0, 0, 1.0]])
KITTI_W, KITTI_H = 1242, 375


def euler_zyx_to_rot(roll, pitch, yaw):
    cp, sp, sr, cr, sy, cy = math.cos(pitch), math.sin(pitch), math.sin(roll), math.cos(roll), math.sin(yaw), math.cos(yaw)
    return np.array([[cp * cy, sr * sp * cy - cr * sy, cr * sp * cy + sr * sy],
                     [cp * sy, sr * sp * sy + cr * cy, cr * sp * sy - sr * cy],
                     [-sp, sr * cp, cr * cp]])


def make_kitti_batch(n_frames, boxes_per_frame=8, seed=20260925, lines_per_box=30, bg_lines=40, img_w=KITTI_W, img_h=KITTI_H,
                     box_w=(60, 320), box_h=(50, 200), poses_only=False):
    """Returns dict(K (F,3,3), T (F,4,4), boxes (F*B,5), lines (M,4), box_ranges, line_ranges, images list of u8 gray).
    poses_only: skip the rendering (images = None); K / T / boxes / lines are the same as in the full batch."""
    import cv2
    rng = np.random.default_rng(seed)
    Ks, Ts, boxes, lines, box_ranges, line_ranges, images = [], [], [], [], [], [], []
    for f in range(n_frames):
        height = rng.uniform(1.5, 1.8)
        tilt = math.radians(rng.uniform(-3, 3))
        roll_opt = math.radians(rng.uniform(-2, 2))
        R = euler_zyx_to_rot(-math.pi / 2 + tilt, roll_opt, 0.0)  # optical frame (z forward, y down) -> world z-up, cf. main.cpp:43-46
        T = np.eye(4)
        T[:3, :3] = R
        T[2, 3] = height
        KinvR = KITTI_K @ np.linalg.inv(R)
        b0, l0 = len(boxes), len(lines)
        img = np.zeros((img_h, img_w), np.uint8)
        for _ in range(boxes_per_frame):
            w = int(rng.uniform(*box_w)); h = int(rng.uniform(*box_h))
            w = min(w, img_w - 42); h = min(h, img_h - 42)
            x = int(rng.uniform(20, img_w - w - 21)); y = int(rng.uniform(20, img_h - h - 21))
            boxes.append([x, y, w, h, rng.uniform(0.3, 0.95)])
            yaw = -math.pi / 2 + math.radians(rng.uniform(-45, 45))
            vps = []
            for d in ((math.cos(yaw), math.sin(yaw), 0), (-math.sin(yaw), math.cos(yaw), 0), (0, 0, 1)):
                v = KinvR @ np.array(d)
                vps.append(v[:2] / v[2])
            for i in range(lines_per_box):
                mx, my = rng.uniform(x, x + w), rng.uniform(y, y + h)
                if i < lines_per_box // 2 + 3:
                    vp = vps[i % 3]
                    ang = math.atan2(vp[1] - my, vp[0] - mx) + math.radians(rng.uniform(-8, 8))
                else:
                    ang = rng.uniform(-math.pi / 2, math.pi / 2)
                ln = rng.uniform(35, max(40.0, 0.8 * min(w, h)))
                dx, dy = 0.5 * ln * math.cos(ang), 0.5 * ln * math.sin(ang)
                p = [mx - dx, my - dy, mx + dx, my + dy]
                p[0] = min(max(p[0], x - 8), x + w + 8); p[2] = min(max(p[2], x - 8), x + w + 8)
                p[1] = min(max(p[1], y - 8), y + h + 8); p[3] = min(max(p[3], y - 8), y + h + 8)
                if p[2] < p[0]:
                    p = [p[2], p[3], p[0], p[1]]
                lines.append(p)
        for _ in range(bg_lines):
            mx, my = rng.uniform(0, img_w - 1), rng.uniform(0, img_h - 1)
            ang = rng.uniform(-math.pi / 2, math.pi / 2); ln = rng.uniform(15, 200)
            p = [mx - 0.5 * ln * math.cos(ang), my - 0.5 * ln * math.sin(ang), mx + 0.5 * ln * math.cos(ang), my + 0.5 * ln * math.sin(ang)]
            p = [min(max(p[0], 0), img_w - 1), min(max(p[1], 0), img_h - 1), min(max(p[2], 0), img_w - 1), min(max(p[3], 0), img_h - 1)]
            if p[2] < p[0]:
                p = [p[2], p[3], p[0], p[1]]
            lines.append(p)
        salt = rng.random((img_h, img_w)) < 0.005   # drawn in every mode: the random stream must not depend on poses_only
        if not poses_only:
            for p in lines[l0:]:
                cv2.line(img, (int(round(p[0])), int(round(p[1]))), (int(round(p[2])), int(round(p[3]))), 255, 1)
            img[salt] = 255
            images.append(img)
        Ks.append(KITTI_K.copy()); Ts.append(T)
        box_ranges.append((b0, len(boxes))); line_ranges.append((l0, len(lines)))
    return dict(K=np.array(Ks), T=np.array(Ts), boxes=np.array(boxes, np.float64).reshape(-1, 5), lines=np.array(lines, np.float64).reshape(-1, 4),
                box_ranges=box_ranges, line_ranges=line_ranges, images=None if poses_only else images, img_w=img_w, img_h=img_h)


def dist_map_for_roi(gray, left, top, width, height):
    """cv::Canny(gray(ROI),80,200) + cv::distanceTransform(255-canny, CV_DIST_L2, 3) (box_proposal_detail.cpp:320-327)."""
    import cv2
    roi = np.ascontiguousarray(gray[top:top + height, left:left + width])
    can = cv2.Canny(roi, 80, 200)
    return cv2.distanceTransform(255 - can, cv2.DIST_L2, 3).astype(np.float32)


def dist_maps_for_tasks(images, tasks, n_tasks, n_map_floats, frame_of=lambda t: t.frame_id, roi_of=None, offset_of=lambda t: t.map_offset):
    """Packed float32 buffer with one distance map per task at the task's map_offset."""
    buf = np.zeros(max(int(n_map_floats), 1) + 16, np.float32)
    for i in range(n_tasks):
        t = tasks[i]
        left, top, w, h = roi_of(t) if roi_of else (t.roi_left, t.roi_top, t.roi_width, t.roi_height)
        dm = dist_map_for_roi(images[frame_of(t)], left, top, w, h)
        o = int(offset_of(t))
        buf[o:o + w * h] = dm.ravel()
    return buf


# ---- small SE(3) / quaternion helpers (numpy, generation only) ------------------------------------
def quat_mul(a, b):  # x y z w
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_rot(q, v):
    qv = np.array(q[:3]); uv = 2 * np.cross(qv, v)
    return v + q[3] * uv + np.cross(qv, uv)


def quat_from_euler(roll, pitch, yaw):
    sy, cy, sp, cp, sr, cr = math.sin(yaw / 2), math.cos(yaw / 2), math.sin(pitch / 2), math.cos(pitch / 2), math.sin(roll / 2), math.cos(roll / 2)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


def se3_mul(a, b):  # 7-vectors x y z qx qy qz qw
    q = quat_mul(a[3:], b[3:])
    if q[3] < 0:
        q = -q
    q = q / np.linalg.norm(q)
    return np.concatenate([a[:3] + quat_rot(a[3:], b[:3]), q])


def se3_inv(a):
    qc = np.array([-a[3], -a[4], -a[5], a[6]])
    return np.concatenate([quat_rot(qc, -a[:3]), qc])


def project_bbox(cube10, Tcw7, K, min_depth=1.0):
    """cuboid::projectOntoImageBbox (g2o_Object.h:181-197) in numpy; None if a corner is closer than min_depth."""
    body = np.array([[1, 1, -1, -1, 1, 1, -1, -1], [1, -1, -1, 1, 1, -1, -1, 1], [-1, -1, -1, -1, 1, 1, 1, 1.0]])
    pts = []
    for k in range(8):
        pw = cube10[:3] + quat_rot(cube10[3:7], body[:, k] * cube10[7:])
        pc = Tcw7[:3] + quat_rot(Tcw7[3:7], pw)
        if pc[2] < min_depth:
            return None
        uv = K @ pc
        pts.append(uv[:2] / uv[2])
    pts = np.array(pts)
    mn, mx = pts.min(0), pts.max(0)
    return np.array([(mn[0] + mx[0]) / 2, (mn[1] + mx[1]) / 2, mx[0] - mn[0], mx[1] - mn[1]])


def make_ba_graph(n_cam=200, n_cube=50, obs_per_cube=80, seed=20260926, with_proj=False):
    """Config #4: returns dict with cams7 (world->camera), cubes10, fixed flags and edge tuples ec / ep / eo."""
    rng = np.random.default_rng(seed)
    radius = 40.0 / (2 * math.pi)
    cams_wc = []
    for i in range(n_cam):
        th = 2 * math.pi * i / n_cam
        pos = np.array([radius * math.cos(th), radius * math.sin(th), 1.6])
        heading = th + math.pi / 2 + math.radians(rng.uniform(-5, 5))
        # optical frame: z forward (heading), y down
        q = quat_mul(quat_from_euler(0, 0, heading - math.pi / 2), quat_from_euler(-math.pi / 2 - math.radians(10), 0, 0))
        cams_wc.append(np.concatenate([pos, q / np.linalg.norm(q)]))
    cams_wc = np.array(cams_wc)
    cubes = []
    for j in range(n_cube):
        r = radius + rng.uniform(-4, 6); th = rng.uniform(0, 2 * math.pi)
        s = rng.uniform(0.2, 2.0, 3)
        q = quat_from_euler(0, 0, rng.uniform(-math.pi, math.pi))
        cubes.append(np.concatenate([[r * math.cos(th), r * math.sin(th), s[2]], q, s]))
    cubes = np.array(cubes)
    ec_cam, ec_cube, ec_meas, ec_info = [], [], [], []
    ep_meas, ep_info, ep_K, ep_ok = [], [], [], []
    K = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]])  # main_obj.cpp:484-486
    for j in range(n_cube):
        # the obs_per_cube nearest keyframes see cuboid j
        d = np.linalg.norm(cams_wc[:, :2] - cubes[j, :2], axis=1)
        for i in np.sort(np.argsort(d)[:obs_per_cube]):
            Tcw = se3_inv(cams_wc[i])
            local = se3_mul(Tcw, cubes[j, :7])
            scale = cubes[j, 7:].copy()
            # noise: 5 cm, 2 deg, 5 %
            dq = quat_from_euler(*np.radians(rng.normal(0, 2, 3)))
            noisy = se3_mul(local, np.concatenate([rng.normal(0, 0.05, 3), dq]))
            scale = scale * (1 + rng.normal(0, 0.05, 3))
            if rng.random() < 0.25:  # different front face: +-90 / 180 deg about the object's z axis
                kq = rng.choice([-1, 1, 2])
                noisy = se3_mul(noisy, np.concatenate([[0, 0, 0], quat_from_euler(0, 0, kq * math.pi / 2)]))
                if kq != 2:
                    scale = scale[[1, 0, 2]]
            ec_cam.append(i); ec_cube.append(j)
            ec_meas.append(np.concatenate([noisy, scale]))
            qual = rng.uniform(0.5, 1.0)
            ec_info.append((np.eye(9) * (2 * qual) ** 2).ravel())  # main_obj.cpp:732, 775-780
            if with_proj:
                bb = project_bbox(cubes[j], Tcw, K)
                ep_ok.append(bb is not None)
                ep_meas.append((bb if bb is not None else np.zeros(4)) + rng.normal(0, 2.0, 4))
                ep_info.append(np.eye(4).ravel()); ep_K.append(K.ravel())
    order = np.lexsort((ec_cube, ec_cam))  # edge ids follow frames, like main_obj.cpp:768 (id = frame index)
    ec_cam = np.array(ec_cam, np.int32)[order]; ec_cube = np.array(ec_cube, np.int32)[order]
    ec_meas = np.array(ec_meas)[order]; ec_info = np.array(ec_info)[order]
    eo_i = np.arange(n_cam - 1, dtype=np.int32); eo_j = eo_i + 1
    eo_meas = []
    cams_cw = np.array([se3_inv(c) for c in cams_wc])
    for i in range(n_cam - 1):
        m = se3_mul(cams_cw[i + 1], se3_inv(cams_cw[i]))  # odom_val = T_cw(i+1) * T_cw(i)^-1, main_obj.cpp:560
        m = se3_mul(m, np.concatenate([rng.normal(0, 0.01, 3), quat_from_euler(*np.radians(rng.normal(0, 0.3, 3)))]))
        eo_meas.append(m)
    eo_info = np.tile(np.eye(6).ravel(), (n_cam - 1, 1))
    # initial estimates: perturbed truth
    cams0 = np.array([se3_mul(np.concatenate([rng.normal(0, 0.03, 3), quat_from_euler(*np.radians(rng.normal(0, 1, 3)))]), c) for c in cams_cw])
    cams0[0] = cams_cw[0]
    cubes0 = cubes.copy()
    cubes0[:, :3] += rng.normal(0, 0.05, (n_cube, 3)); cubes0[:, 7:] *= (1 + rng.normal(0, 0.05, (n_cube, 3)))
    cam_fixed = np.zeros(n_cam, np.int32); cam_fixed[0] = 1
    out = dict(cams7=cams0, cubes10=cubes0, cam_fixed=cam_fixed, cube_fixed=np.zeros(n_cube, np.int32),
               ec=(ec_cam, ec_cube, ec_meas, ec_info), eo=(eo_i, eo_j, np.array(eo_meas), eo_info), ep=None)
    if with_proj:
        ok = np.array(ep_ok)[order]  # EdgeSE3CuboidProj only where the cuboid is fully in front of the camera
        out["ep"] = (ec_cam[ok].copy(), ec_cube[ok].copy(), np.array(ep_meas)[order][ok], np.array(ep_info)[order][ok], np.array(ep_K)[order][ok])
    return out


def dist_map_for_roi_reference(gray, left, top, width, height, return_edges=False):
    """What the reference's C++ computes for `cv::Canny(gray_img(object_bbox), ...)` + `cv::distanceTransform(..., CV_DIST_L2, 3)`
    (box_proposal_detail.cpp:320-327): on a cv::Mat ROI *view* the Sobel filter of cv::Canny sees the ROI's real neighbours inside the
    parent image (python slices lose that), and OpenCV's own 3x3 chamfer transform is the fixed-point one (a cv2 build with IPP swaps in
    a closed-source float variant, so IPP is switched off around the call).  This is the parity target of the GPU path (distmap.cu)."""
    import cv2
    dx = cv2.Sobel(gray, cv2.CV_16S, 1, 0, ksize=3, borderType=cv2.BORDER_REPLICATE)
    dy = cv2.Sobel(gray, cv2.CV_16S, 0, 1, ksize=3, borderType=cv2.BORDER_REPLICATE)
    sl = (slice(top, top + height), slice(left, left + width))
    edges = cv2.Canny(np.ascontiguousarray(dx[sl]), np.ascontiguousarray(dy[sl]), 80, 200)
    ipp = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        dm = cv2.distanceTransform(255 - edges, cv2.DIST_L2, 3).astype(np.float32)
    finally:
        cv2.ipp.setUseIPP(ipp)
    return (dm, edges) if return_edges else dm


def make_lsd_frames(n_frames, img_w=640, img_h=480, seed=20260927, n_polys=10, n_lines=25, noise_sigma=3.0, texture=0.0):
    """Synthetic gray frames for the line detector (BASELINE config #3: 640x480): filled quadrilaterals and thick line segments on a
    shaded background, blurred by a small PSF, plus white sensor noise (and optionally band-limited texture).  uint8 (n, h, w)."""
    import cv2
    rng = np.random.default_rng(seed)
    out = np.zeros((n_frames, img_h, img_w), np.uint8)
    yy, xx = np.mgrid[0:img_h, 0:img_w]
    for f in range(n_frames):
        gx, gy = rng.uniform(-0.05, 0.05, 2)
        img = np.clip(rng.uniform(60, 160) + gx * (xx - img_w / 2) + gy * (yy - img_h / 2), 0, 255).astype(np.uint8)
        for _ in range(n_polys):
            cx, cy = rng.uniform(0, img_w), rng.uniform(0, img_h)
            sz = rng.uniform(20, 160, 2)
            ang = rng.uniform(0, np.pi)
            c, s = np.cos(ang), np.sin(ang)
            base = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64) * sz * 0.5 * rng.uniform(0.6, 1.0, (4, 2))
            pts = np.c_[cx + base[:, 0] * c - base[:, 1] * s, cy + base[:, 0] * s + base[:, 1] * c]
            cv2.fillPoly(img, [np.round(pts).astype(np.int32)], int(rng.integers(0, 256)), lineType=cv2.LINE_AA)
        for _ in range(n_lines):
            p1 = (int(rng.integers(0, img_w)), int(rng.integers(0, img_h)))
            ln, ang = rng.uniform(15, 250), rng.uniform(0, 2 * np.pi)
            p2 = (int(p1[0] + ln * np.cos(ang)), int(p1[1] + ln * np.sin(ang)))
            cv2.line(img, p1, p2, int(rng.integers(0, 256)), int(rng.integers(1, 6)), lineType=cv2.LINE_AA)
        img = cv2.GaussianBlur(img, (0, 0), 0.8).astype(np.float64)
        if texture > 0:
            img += cv2.GaussianBlur(rng.normal(0, 1, img.shape), (0, 0), 2.0) * texture * 6.0
        img += rng.normal(0, noise_sigma, img.shape)
        out[f] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out
