"""ctypes bindings for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

The oracle is the checker: tests, __graft_entry__.smoke() and bench.py's CPU legs may load it,
the product (cube_slam_wu_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


class Params(C.Structure):
    _fields_ = [("consider_config_1", C.c_int), ("consider_config_2", C.c_int),
                ("whether_sample_cam_roll_pitch", C.c_int), ("whether_sample_bbox_height", C.c_int),
                ("max_cuboid_num", C.c_int), ("leak_cam_state", C.c_int),
                ("nominal_skew_ratio", C.c_double), ("max_cut_skew", C.c_double), ("libm_atan2", C.c_int), ("reserved", C.c_int)]


class Task(C.Structure):
    _fields_ = [("box_id", C.c_int), ("hs_id", C.c_int), ("down_expand", C.c_int), ("left", C.c_int), ("top", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("n_top", C.c_int), ("map_offset", C.c_longlong)]


class Cuboid(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("scale", C.c_double * 3), ("rotY", C.c_double), ("box_config_type", C.c_double * 2),
                ("box_corners_3d_world", C.c_double * 24), ("rect_detect_2d", C.c_double * 4),
                ("edge_distance_error", C.c_double), ("edge_angle_error", C.c_double), ("normalized_error", C.c_double),
                ("skew_ratio", C.c_double), ("down_expand_height", C.c_double), ("camera_roll_delta", C.c_double),
                ("camera_pitch_delta", C.c_double), ("box_corners_2d", C.c_int * 16), ("task_id", C.c_int), ("raw_cube_ind", C.c_int)]


class BAEdges(C.Structure):
    _fields_ = [("n_ec", C.c_int), ("ec_cam", C.c_void_p), ("ec_cube", C.c_void_p), ("ec_meas10", C.c_void_p), ("ec_info81", C.c_void_p),
                ("n_ep", C.c_int), ("ep_cam", C.c_void_p), ("ep_cube", C.c_void_p), ("ep_meas4", C.c_void_p), ("ep_info16", C.c_void_p), ("ep_K9", C.c_void_p),
                ("n_eo", C.c_int), ("eo_i", C.c_void_p), ("eo_j", C.c_void_p), ("eo_meas7", C.c_void_p), ("eo_info36", C.c_void_p)]


class BAOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ec_err", "ec_Ji", "ec_Jj", "ep_err", "ep_Ji", "ep_Jj", "eo_err", "eo_Ji", "eo_Jj",
                                          "H_cam", "b_cam", "H_cube", "b_cube", "ec_Hij", "ep_Hij", "eo_Hij")]


def build(force=False):
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_proposal.cpp", "oracle_ba.cpp", "oracle_lsd.cpp", "oracle_lbd.cpp", "oracle_edlines.cpp", "oracle_math.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_detect_frame.restype = C.c_void_p
        L.orc_num_scored.restype = C.c_longlong
        L.orc_num_enum.restype = C.c_longlong
        L.orc_detect_batch.restype = C.c_longlong
        L.orc_ba_linearize.restype = C.c_double
        L.orc_det_atan2.restype = C.c_double
        L.orc_det_atan2.argtypes = [C.c_double, C.c_double]
        L.orc_lsd_fast_atan2.restype = C.c_float
        L.orc_lsd_fast_atan2.argtypes = [C.c_float, C.c_float]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(**kw):
    p = Params(1, 1, 1, 0, 1, 1, 1.0, 3.0, 0, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def plan(boxes, img_w, img_h, sample_height=False):
    boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 5)
    cap = 3 * len(boxes) + 1
    tasks = (Task * cap)()
    n = lib().orc_plan(_p(boxes), len(boxes), img_w, img_h, int(sample_height), tasks, cap)
    assert n >= 0
    return [tasks[i] for i in range(n)]


class FrameResult:
    """Python view of one oracle detect_cuboid() run."""

    def __init__(self, h):
        L = lib()
        self.n_scored = L.orc_num_scored(h)
        self.n_enum = L.orc_num_enum(h)
        self.tasks = []
        for t in range(L.orc_num_tasks(h)):
            tk = Task(); nv = C.c_int(); ne = C.c_int(); nm = C.c_int(); nk = C.c_int()
            L.orc_task_info(h, t, C.byref(tk), C.byref(nv), C.byref(ne), C.byref(nm), C.byref(nk))
            d = dict(task=tk, n_valid=nv.value, n_enum=ne.value,
                     rows=np.zeros((nv.value, 9)), corners=np.zeros((nv.value, 16)), hyp_id=np.zeros(nv.value, np.int32),
                     merged=np.zeros((nm.value, 4)), keep=np.zeros(nk.value, np.int32), norm_score=np.zeros(nk.value))
            L.orc_task_data(h, t, _p(d["rows"]), _p(d["corners"]), _p(d["hyp_id"]), _p(d["merged"]), _p(d["keep"]), _p(d["norm_score"]))
            self.tasks.append(d)
        self.boxes = []


def detect_frame(K, T, img_w, img_h, boxes, lines, dist_maps, params=None, n_boxes=None):
    L = lib()
    params = params or default_params()
    K = np.ascontiguousarray(K, np.float64); T = np.ascontiguousarray(T, np.float64)
    boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 5)
    lines = np.ascontiguousarray(lines, np.float64).reshape(-1, 4)
    dist_maps = np.ascontiguousarray(dist_maps, np.float32)
    h = C.c_void_p(L.orc_detect_frame(_p(K), _p(T), img_w, img_h, _p(boxes), len(boxes), _p(lines), len(lines), _p(dist_maps), C.byref(params)))
    try:
        R = FrameResult(h)
        for b in range(len(boxes)):
            nr = C.c_int(); ns = C.c_int()
            L.orc_box_info(h, b, C.byref(nr), C.byref(ns))
            raw = (Cuboid * max(nr.value, 1))()
            comb = np.zeros(nr.value); srt = np.zeros(ns.value, np.int32)
            L.orc_box_data(h, b, raw, _p(comb), _p(srt))
            R.boxes.append(dict(raw=[raw[i] for i in range(nr.value)], combined=comb, sorted=srt))
        return R
    finally:
        L.orc_free(h)


def ba_edges(ec=None, ep=None, eo=None):
    """ec=(cam,cube,meas10,info81), ep=(cam,cube,meas4,info16,K9), eo=(i,j,meas7,info36); arrays kept alive on the struct."""
    E = BAEdges()
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(a, dt); keep.append(a); return a.ctypes.data_as(C.c_void_p)
    if ec is not None:
        E.n_ec = len(ec[0]); E.ec_cam = arr(ec[0], np.int32); E.ec_cube = arr(ec[1], np.int32); E.ec_meas10 = arr(ec[2], np.float64); E.ec_info81 = arr(ec[3], np.float64)
    if ep is not None:
        E.n_ep = len(ep[0]); E.ep_cam = arr(ep[0], np.int32); E.ep_cube = arr(ep[1], np.int32); E.ep_meas4 = arr(ep[2], np.float64); E.ep_info16 = arr(ep[3], np.float64); E.ep_K9 = arr(ep[4], np.float64)
    if eo is not None:
        E.n_eo = len(eo[0]); E.eo_i = arr(eo[0], np.int32); E.eo_j = arr(eo[1], np.int32); E.eo_meas7 = arr(eo[2], np.float64); E.eo_info36 = arr(eo[3], np.float64)
    E._keep = keep
    return E


def ba_linearize(cams7, cam_fixed, cubes10, cube_fixed, E, n_threads=1):
    cams7 = np.ascontiguousarray(cams7, np.float64).reshape(-1, 7); cubes10 = np.ascontiguousarray(cubes10, np.float64).reshape(-1, 10)
    cam_fixed = np.ascontiguousarray(cam_fixed, np.int32); cube_fixed = np.ascontiguousarray(cube_fixed, np.int32)
    nc, nq = len(cams7), len(cubes10)
    out = dict(ec_err=np.zeros((E.n_ec, 9)), ec_Ji=np.zeros((E.n_ec, 54)), ec_Jj=np.zeros((E.n_ec, 81)),
               ep_err=np.zeros((E.n_ep, 4)), ep_Ji=np.zeros((E.n_ep, 24)), ep_Jj=np.zeros((E.n_ep, 36)),
               eo_err=np.zeros((E.n_eo, 6)), eo_Ji=np.zeros((E.n_eo, 36)), eo_Jj=np.zeros((E.n_eo, 36)),
               H_cam=np.zeros((nc, 36)), b_cam=np.zeros((nc, 6)), H_cube=np.zeros((nq, 81)), b_cube=np.zeros((nq, 9)),
               ec_Hij=np.zeros((E.n_ec, 54)), ep_Hij=np.zeros((E.n_ep, 54)), eo_Hij=np.zeros((E.n_eo, 36)))
    O = BAOut(*[_p(out[n]) for n, _ in BAOut._fields_])
    out["chi2"] = lib().orc_ba_linearize(nc, _p(cams7), _p(cam_fixed), nq, _p(cubes10), _p(cube_fixed), C.byref(E), C.byref(O), n_threads)
    return out


def ba_set_delta(delta=1e-9):
    """step of the central differences (reference: 1e-9, base_binary_edge.hpp:147); test knob, see oracle_ba.cpp"""
    lib().orc_ba_set_delta(C.c_double(delta))


def ba_optimize(cams7, cam_fixed, cubes10, cube_fixed, E, iterations):
    cams7 = np.array(cams7, np.float64).reshape(-1, 7).copy(); cubes10 = np.array(cubes10, np.float64).reshape(-1, 10).copy()
    cam_fixed = np.ascontiguousarray(cam_fixed, np.int32); cube_fixed = np.ascontiguousarray(cube_fixed, np.int32)
    chi = C.c_double()
    it = lib().orc_ba_optimize(len(cams7), _p(cams7), _p(cam_fixed), len(cubes10), _p(cubes10), _p(cube_fixed), C.byref(E), iterations, C.byref(chi))
    return cams7, cubes10, it, chi.value


def _v(n):
    return np.zeros(n, np.float64)


def cuboid_from_minimal(v9):
    o = _v(10); lib().orc_cuboid_from_minimal(_p(np.ascontiguousarray(v9, np.float64)), _p(o)); return o


def cuboid_transform_to(c10, Twc7):
    o = _v(10); lib().orc_cuboid_transform_to(_p(np.ascontiguousarray(c10, np.float64)), _p(np.ascontiguousarray(Twc7, np.float64)), _p(o)); return o


def cuboid_transform_from(c10, Twc7):
    o = _v(10); lib().orc_cuboid_transform_from(_p(np.ascontiguousarray(c10, np.float64)), _p(np.ascontiguousarray(Twc7, np.float64)), _p(o)); return o


def se3_inverse(a7):
    o = _v(7); lib().orc_se3_inverse(_p(np.ascontiguousarray(a7, np.float64)), _p(o)); return o


def se3_mul(a7, b7):
    o = _v(7); lib().orc_se3_mul(_p(np.ascontiguousarray(a7, np.float64)), _p(np.ascontiguousarray(b7, np.float64)), _p(o)); return o


def se3_log(a7):
    o = _v(6); lib().orc_se3_log(_p(np.ascontiguousarray(a7, np.float64)), _p(o)); return o


def se3_exp(u6):
    o = _v(7); lib().orc_se3_exp(_p(np.ascontiguousarray(u6, np.float64)), _p(o)); return o


# ---- LSD line detector (oracle/oracle_lsd.cpp) ------------------------------------------------------------------------------------
def lsd_gauss_kernel():
    k = np.zeros(7); lib().orc_lsd_gauss_kernel(_p(k)); return k


def lsd_maps(gray, gauss7=None):
    """scaled image, gradient norm, level-line angle (-1024 = undefined) as the reference's double pipeline computes them."""
    gray = np.ascontiguousarray(gray, np.uint8); h, w = gray.shape
    W, H = int(np.rint(w * 0.8)), int(np.rint(h * 0.8))
    sc = np.zeros((H, W)); mg = np.zeros((H, W)); an = np.zeros((H, W)); sw = C.c_int(); sh = C.c_int()
    lib().orc_lsd_maps(_p(gray), w, h, _p(gauss7), _p(sc), _p(mg), _p(an), C.byref(sw), C.byref(sh))
    assert (sw.value, sh.value) == (W, H)
    return sc, mg, an


def lsd_detect(gray, seed_order=0, libm_trig=0, refine=2, mode=1, length_thres=15.0, gauss7=None, scaled=None, cap=20000, details=False):
    """seed_order 0 = reference (raster), 1 = cv2 4.x; mode 1 = detect_filter_lines output, 0 = raw LSD segments."""
    gray = np.ascontiguousarray(gray, np.uint8); h, w = gray.shape
    out = np.zeros((cap, 4), np.float32); det = np.zeros((cap, 12))
    if scaled is not None:
        scaled = np.ascontiguousarray(scaled, np.float64)
    n = lib().orc_lsd_detect(_p(gray), w, h, _p(gauss7), _p(scaled), int(seed_order), int(libm_trig), int(refine), int(mode), C.c_float(length_thres),
                             _p(out), _p(det), cap)
    assert 0 <= n <= cap
    return (out[:n].copy(), det[:n].copy()) if details else out[:n].copy()


def lsd_spec_sim(gray, K):
    """orc_lsd_spec_sim: wave-speculative seed processing (study, see oracle/oracle_lsd.cpp).  Returns (segments, stats[6])."""
    gray = np.ascontiguousarray(gray, np.uint8); h, w = gray.shape
    out = np.zeros((20000, 4), np.float32); st = np.zeros(6)
    n = lib().orc_lsd_spec_sim(_p(gray), w, h, int(K), _p(out), 20000, _p(st))
    return out[:n].copy(), st


# ---- LBD line descriptor (oracle/oracle_lbd.cpp) -----------------------------------------------------------------------------------
def lbd_set_blur_generation(gen):
    """4 (default): cv2 4.x integer taps of the 8-bit 5x5 Gaussian; 3: OpenCV <= 3.4.0's (what the reference's author ran).  Affects the
    LBD and EDLines oracles (both start from BinaryDescriptor's blurred frame)."""
    lib().orc_lbd_set_blur_generation(int(gen))


def lbd_gradients(gray):
    """blurred frame (u8) and the int16 Sobel images BinaryDescriptor::computeSobel produces."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    b = np.zeros((h, w), np.uint8); dx = np.zeros((h, w), np.int16); dy = np.zeros((h, w), np.int16)
    lib().orc_lbd_gradients(_p(gray), w, h, _p(b), _p(dx), _p(dy))
    return b, dx, dy


def lbd_weights():
    g = np.zeros(63, np.float32); l = np.zeros(21, np.float32)
    lib().orc_lbd_weights(_p(g), _p(l))
    return g, l


def lbd_describe(gray, lines, libm_trig=0):
    """72-float and 32-byte LBD descriptors of `lines` (n x 4 float32) + key-line fields {angle, numOfPixels, lineLength}."""
    gray = np.ascontiguousarray(gray, np.uint8)
    lines = np.ascontiguousarray(lines, np.float32).reshape(-1, 4)
    h, w = gray.shape
    n = len(lines)
    d72 = np.zeros((n, 72), np.float32); d32 = np.zeros((n, 32), np.uint8); kl = np.zeros((n, 3), np.float32)
    if n:
        lib().orc_lbd_describe(_p(gray), w, h, _p(lines), n, int(libm_trig), _p(d72), _p(d32), _p(kl))
    return d72, d32, kl


# ---- EDLines line detector (oracle/oracle_edlines.cpp; no GPU path yet) -------------------------------------------------------------
def edlines_maps(gray, anchor_cap=400000):
    """thresholded gradient map (int16), direction map (255 / 0), anchors (n, 2) [x, y] in the reference's scan order."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    g = np.zeros((h, w), np.int16); d = np.zeros((h, w), np.uint8); a = np.zeros((anchor_cap, 2), np.uint32)
    n = lib().orc_edlines_maps(_p(gray), w, h, _p(g), _p(d), _p(a), anchor_cap)
    return g, d, a[:max(n, 0)].astype(np.int64)


def edlines_chains(gray):
    """edge chains of EDLineDetector::EdgeDrawing: list of (k_i, 2) [x, y] arrays."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = w * h
    xy = np.zeros((cap, 2), np.uint32); sid = np.zeros(cap // 16 + 2, np.uint32); npx = C.c_int()
    n = lib().orc_edlines_chains(_p(gray), w, h, _p(xy), cap, _p(sid), len(sid) - 1, C.byref(npx))
    return [xy[sid[i]:sid[i + 1]].astype(np.int64) for i in range(max(n, 0))]


def edlines_detect(gray, filter=True, length_thres=15.0, cap=20000, libm_trig=False):
    """key lines (n, 4) float32 [x1 y1 x2 y2] and (n, 3) {direction, numOfPixels, lineLength}."""
    lib().orc_edlines_set_libm_trig(int(libm_trig))
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.zeros((cap, 4), np.float32); ex = np.zeros((cap, 3), np.float32)
    n = lib().orc_edlines_detect(_p(gray), w, h, int(filter), C.c_float(length_thres), _p(out), _p(ex), cap)
    return out[:max(n, 0)].copy(), ex[:max(n, 0)].copy()


def lbd_describe_keylines(gray, lines, direction, num_px, libm_trig=0):
    """LBD descriptors of key lines with detector-supplied fields (EDLines: direction = lineDirection_, numOfPixels = chain-segment pixels)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    lines = np.ascontiguousarray(lines, np.float32).reshape(-1, 4)
    fields = np.ascontiguousarray(np.stack([np.asarray(direction, np.float32), np.asarray(num_px, np.float32)], axis=1), np.float32)
    h, w = gray.shape
    n = len(lines)
    d72 = np.zeros((n, 72), np.float32); d32 = np.zeros((n, 32), np.uint8)
    if n:
        lib().orc_lbd_describe_keylines(_p(gray), w, h, _p(lines), _p(fields), n, int(libm_trig), _p(d72), _p(d32))
    return d72, d32
