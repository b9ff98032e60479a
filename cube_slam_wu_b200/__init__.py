"""cube_slam_wu_b200 -- B200-native CubeSLAM hot path (cuboid proposal scoring + cuboid-BA linearisation).

The product is the C-ABI shared library `libcubeslam_b200.so` (include/cubeslam_b200.h); this package is the thin
ctypes binding used by the tests and the benchmark.  There is no CPU fallback: if the CUDA library is missing or no
GPU is usable, the calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSB_LIB") or os.path.join(_PKG, "libcubeslam_b200.so")  # CSB_LIB: development override (A/B timing of library builds)
_LIB = None

CSB_OK, CSB_ERR_INVALID, CSB_ERR_CUDA, CSB_ERR_CAPACITY, CSB_ERR_STATE = 0, -1, -2, -3, -4
CSB_OPT_GRAY_GATHER = 1


class CsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cubeslam_b200 error %d: %s" % (code, msg))
        self.code = code


class DetectParams(C.Structure):
    _fields_ = [("consider_config_1", C.c_int32), ("consider_config_2", C.c_int32), ("whether_sample_cam_roll_pitch", C.c_int32),
                ("whether_sample_bbox_height", C.c_int32), ("max_cuboid_num", C.c_int32), ("reserved", C.c_int32),
                ("nominal_skew_ratio", C.c_double), ("max_cut_skew", C.c_double)]

    @staticmethod
    def default(**kw):
        p = DetectParams(1, 1, 1, 0, 1, 0, 1.0, 3.0)  # detect_3d_cuboid.h:109-116
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class Frame(C.Structure):
    _fields_ = [("Kalib", C.c_double * 9), ("transToWolrd", C.c_double * 16), ("img_width", C.c_int32), ("img_height", C.c_int32),
                ("box_begin", C.c_int32), ("box_end", C.c_int32), ("line_begin", C.c_int32), ("line_end", C.c_int32)]


class Task(C.Structure):
    _fields_ = [("frame_id", C.c_int32), ("box_id", C.c_int32), ("hs_id", C.c_int32), ("down_expand", C.c_int32),
                ("roi_left", C.c_int32), ("roi_top", C.c_int32), ("roi_width", C.c_int32), ("roi_height", C.c_int32),
                ("n_top", C.c_int32), ("n_enum", C.c_int32), ("map_offset", C.c_int64)]


class Cuboid(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("scale", C.c_double * 3), ("rotY", C.c_double), ("box_config_type", C.c_double * 2),
                ("box_corners_3d_world", C.c_double * 24), ("rect_detect_2d", C.c_double * 4),
                ("edge_distance_error", C.c_double), ("edge_angle_error", C.c_double), ("normalized_error", C.c_double), ("skew_ratio", C.c_double),
                ("down_expand_height", C.c_double), ("camera_roll_delta", C.c_double), ("camera_pitch_delta", C.c_double),
                ("box_corners_2d", C.c_int32 * 16), ("task_id", C.c_int32), ("raw_cube_ind", C.c_int32), ("rank_index", C.c_int32), ("reserved", C.c_int32)]


class DetectStats(C.Structure):
    _fields_ = [("n_enumerated", C.c_int64), ("n_scored", C.c_int64), ("n_kept", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("n_kernel_launches", C.c_int32), ("n_tasks_smem_map", C.c_int32),
                ("gpu_ms_prep", C.c_float), ("gpu_ms_score", C.c_float), ("gpu_ms_select", C.c_float), ("gpu_ms_recover", C.c_float),
                ("gpu_ms_rank", C.c_float), ("gpu_ms_distmap", C.c_float)]


class BAGraph(C.Structure):
    _fields_ = [("n_cam", C.c_int32), ("n_cube", C.c_int32), ("cam_fixed", C.c_void_p), ("cube_fixed", C.c_void_p),
                ("n_ec", C.c_int32), ("ec_cam", C.c_void_p), ("ec_cube", C.c_void_p), ("ec_meas", C.c_void_p), ("ec_info", C.c_void_p),
                ("n_ep", C.c_int32), ("ep_cam", C.c_void_p), ("ep_cube", C.c_void_p), ("ep_meas", C.c_void_p), ("ep_info", C.c_void_p), ("ep_K", C.c_void_p),
                ("n_eo", C.c_int32), ("eo_cam_i", C.c_void_p), ("eo_cam_j", C.c_void_p), ("eo_meas", C.c_void_p), ("eo_info", C.c_void_p)]


BA_OUT_FIELDS = ("ec_err", "ec_Ji", "ec_Jj", "ep_err", "ep_Ji", "ep_Jj", "eo_err", "eo_Ji", "eo_Jj", "H_cam", "b_cam", "H_cube", "b_cube",
                 "ec_Hij", "ep_Hij", "eo_Hij", "chi2")


class BAOptimizeStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("trials", C.c_int32), ("n_kernel_launches", C.c_int32), ("schur_dim", C.c_int32),
                ("chi2", C.c_double), ("lambda_", C.c_double), ("gpu_ms", C.c_float), ("n_launches", C.c_int32)]


class BAFrame(C.Structure):
    _fields_ = [("cam7", C.c_void_p), ("cam_fixed", C.c_int32), ("n_new_cubes", C.c_int32), ("new_cubes10", C.c_void_p), ("new_cube_fixed", C.c_void_p),
                ("n_ec", C.c_int32), ("ec_cube", C.c_void_p), ("ec_meas", C.c_void_p), ("ec_info", C.c_void_p),
                ("n_eo", C.c_int32), ("eo_cam_i", C.c_void_p), ("eo_meas", C.c_void_p), ("eo_info", C.c_void_p)]


class BAOutput(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BA_OUT_FIELDS]


class LsdParams(C.Structure):
    _fields_ = [("line_length_thres", C.c_float), ("filter", C.c_int32), ("max_lines", C.c_int32), ("unit_link_deg", C.c_int32)]


class LsdStats(C.Structure):
    _fields_ = [("n_lines", C.c_int64), ("n_regions", C.c_int64), ("n_region_px", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("scaled_width", C.c_int32), ("scaled_height", C.c_int32), ("n_kernel_launches", C.c_int32), ("reserved", C.c_int32),
                ("gpu_ms_maps", C.c_float), ("gpu_ms_grow", C.c_float), ("grow_cycles", C.c_int64 * 5), ("n_merge_rounds", C.c_int64), ("n_unit_conflicts", C.c_int64)]


class LbdStats(C.Structure):
    _fields_ = [("n_lines", C.c_int64), ("n_samples", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("n_kernel_launches", C.c_int32), ("reserved", C.c_int32), ("gpu_ms_grad", C.c_float), ("gpu_ms_describe", C.c_float)]


class EdlinesStats(C.Structure):
    _fields_ = [("n_lines", C.c_int64), ("n_anchors", C.c_int64), ("n_chain_px", C.c_int64), ("n_chains", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("n_kernel_launches", C.c_int32), ("n_frames_failed", C.c_int32),
                ("gpu_ms_maps", C.c_float), ("gpu_ms_draw", C.c_float), ("gpu_ms_fit", C.c_float), ("reserved", C.c_float)]


def lib():
    """Load the CUDA library; raises if it has not been built (there is no fallback path)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise CsbError(CSB_ERR_STATE, "libcubeslam_b200.so is not built: run `python -m cube_slam_wu_b200.build` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        L.csb_last_error.restype = C.c_char_p
        L.csb_version.restype = C.c_char_p
        _LIB = L
    return _LIB


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def make_frames(Ks, Ts, img_w, img_h, box_ranges, line_ranges):
    n = len(Ks)
    fr = (Frame * n)()
    for i in range(n):
        fr[i].Kalib[:] = np.asarray(Ks[i], np.float64).ravel().tolist()
        fr[i].transToWolrd[:] = np.asarray(Ts[i], np.float64).ravel().tolist()
        fr[i].img_width, fr[i].img_height = int(img_w), int(img_h)
        fr[i].box_begin, fr[i].box_end = int(box_ranges[i][0]), int(box_ranges[i][1])
        fr[i].line_begin, fr[i].line_end = int(line_ranges[i][0]), int(line_ranges[i][1])
    return fr


def detect_plan(frames, boxes, params):
    """csb_detect_plan(): host-only task / ROI planning.  Returns (tasks ctypes array, n_map_floats)."""
    L = lib()
    boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 5)
    nt = C.c_int(); nm = C.c_int64()
    rc = L.csb_detect_plan(frames, len(frames), _p(boxes), len(boxes), C.byref(params), None, 0, C.byref(nt), C.byref(nm))
    if rc != CSB_OK:
        raise CsbError(rc, "csb_detect_plan failed")
    tasks = (Task * max(nt.value, 1))()
    rc = L.csb_detect_plan(frames, len(frames), _p(boxes), len(boxes), C.byref(params), tasks, nt.value, C.byref(nt), C.byref(nm))
    if rc != CSB_OK:
        raise CsbError(rc, "csb_detect_plan failed")
    return tasks, nt.value, nm.value


class Context:
    """csb_context wrapper (one CUDA device + stream)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        rc = lib().csb_create(C.byref(self._h), int(device))
        if rc != CSB_OK:
            raise CsbError(rc, "csb_create failed: no usable CUDA device %d (this library has no CPU fallback)" % device)
        if stream is not None:
            self._chk(lib().csb_set_stream(self._h, C.c_void_p(int(stream))))

    def close(self):
        if self._h:
            lib().csb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != CSB_OK:
            raise CsbError(rc, (lib().csb_last_error(self._h) or b"").decode())

    def set_blur_generation(self, generation):
        """csb_set_blur_generation(): 4 (default) = OpenCV 4.x taps of the 8-bit 5x5 Gaussian in front of LBD / EDLines, 3 = OpenCV <= 3.4.0's."""
        self._chk(lib().csb_set_blur_generation(self._h, int(generation)))

    def set_option(self, option, value):
        """csb_set_option(): CSB_OPT_GRAY_GATHER (1) = fetch only the ROI segments of pinned gray frames (default on)."""
        self._chk(lib().csb_set_option(self._h, int(option), int(value)))

    def synchronize(self):
        self._chk(lib().csb_synchronize(self._h))

    # ---- proposal half -------------------------------------------------------------------------
    def _args(self, frames, boxes, lines, tasks, n_tasks, dist_maps, n_map_floats, params):
        nb = boxes.shape[0] if hasattr(boxes, "shape") else self._nb
        nl = lines.shape[0] if hasattr(lines, "shape") else self._nl
        return (self._h, frames, len(frames), _p(boxes), nb, _p(lines), nl, tasks, n_tasks, _p(dist_maps), C.c_int64(n_map_floats), C.byref(params))

    def detect_batch(self, frames, boxes, lines, tasks, n_tasks, dist_maps, n_map_floats, params, want_stats=True):
        """csb_detect_batch(): host buffers in, host cuboids out (copies inside).  boxes (n,5) f64, lines (m,4) f64, dist_maps f32."""
        nb = boxes.shape[0]
        kmax = params.max_cuboid_num
        cub = (Cuboid * max(nb * kmax, 1))()
        ncub = np.zeros(max(nb, 1), np.int32)
        st = DetectStats()
        self._chk(lib().csb_detect_batch(*self._args(frames, boxes, lines, tasks, n_tasks, dist_maps, n_map_floats, params), cub, _p(ncub),
                                         C.byref(st) if want_stats else None))
        return cub, ncub[:nb], st

    def detect_batch_gray(self, frames, boxes, lines, tasks, n_tasks, gray, params, want_stats=True):
        """csb_detect_batch_gray(): like detect_batch but from packed uint8 gray frames (Canny + distance transform on the GPU)."""
        nb = boxes.shape[0]
        kmax = params.max_cuboid_num
        cub = (Cuboid * max(nb * kmax, 1))()
        ncub = np.zeros(max(nb, 1), np.int32)
        st = DetectStats()
        self._nb, self._kmax = nb, kmax
        self._chk(lib().csb_detect_batch_gray(self._h, frames, len(frames), _p(boxes), nb, _p(lines), lines.shape[0], tasks, n_tasks, _p(gray),
                                              C.c_int64(gray.size), C.byref(params), cub, _p(ncub), C.byref(st) if want_stats else None))
        return cub, ncub[:nb], st

    def detect_upload_gray(self, frames, boxes, lines, tasks, n_tasks, gray, params):
        self._nb, self._kmax = boxes.shape[0], params.max_cuboid_num
        self._chk(lib().csb_detect_upload_gray(self._h, frames, len(frames), _p(boxes), boxes.shape[0], _p(lines), lines.shape[0], tasks, n_tasks, _p(gray),
                                               C.c_int64(gray.size), C.byref(params)))

    def debug_map(self, task_index, task, edges=False):
        """Distance map (and optionally the 0/1/2 Canny map) of tasks[task_index] after a run."""
        n = task.roi_width * task.roi_height
        dm = np.zeros(n, np.float32)
        ed = np.zeros(n, np.uint8) if edges else None
        self._chk(lib().csb_detect_debug_map(self._h, int(task_index), _p(dm), _p(ed), n))
        shape = (task.roi_height, task.roi_width)
        return (dm.reshape(shape), ed.reshape(shape)) if edges else dm.reshape(shape)

    def score_phases(self, reset=True):
        out = np.zeros(12, np.uint64)
        self._chk(lib().csb_detect_debug_score_phases(self._h, _p(out), int(reset)))
        return out

    def debug_atan2(self, y, x):
        y = np.ascontiguousarray(y, np.float64); x = np.ascontiguousarray(x, np.float64)
        out = np.zeros_like(y); nf = C.c_int()
        self._chk(lib().csb_detect_debug_atan2(self._h, _p(y), _p(x), _p(out), int(y.size), C.byref(nf)))
        return out, nf.value

    def detect_upload(self, frames, boxes, lines, tasks, n_tasks, dist_maps, n_map_floats, params):
        self._nb, self._kmax = boxes.shape[0], params.max_cuboid_num
        self._chk(lib().csb_detect_upload(*self._args(frames, boxes, lines, tasks, n_tasks, dist_maps, n_map_floats, params)))

    def detect_run(self, timed=False):
        self._chk(lib().csb_detect_run(self._h, int(timed)))

    def detect_download(self):
        cub = (Cuboid * max(self._nb * self._kmax, 1))()
        ncub = np.zeros(max(self._nb, 1), np.int32)
        st = DetectStats()
        self._chk(lib().csb_detect_download(self._h, cub, _p(ncub), C.byref(st)))
        return cub, ncub[:self._nb], st

    def debug_task(self, task_id, capacity):
        nv = C.c_int32(); nm = C.c_int32(); nk = C.c_int32()
        self._chk(lib().csb_detect_debug_task(self._h, task_id, C.byref(nv), C.byref(nm), C.byref(nk), None, None, None, None, None, None, None, 0))
        cap = max(nv.value, nm.value, nk.value, 1)
        out = dict(n_valid=nv.value, n_merged=nm.value, n_keep=nk.value, hyp_id=np.zeros(cap, np.int32), dist=np.zeros(cap), angle=np.zeros(cap),
                   corners=np.zeros((cap, 16)), merged=np.zeros((cap, 4)), keep=np.zeros(cap, np.int32), norm_score=np.zeros(cap))
        self._chk(lib().csb_detect_debug_task(self._h, task_id, C.byref(nv), C.byref(nm), C.byref(nk), _p(out["hyp_id"]), _p(out["dist"]), _p(out["angle"]),
                                              _p(out["corners"]), _p(out["merged"]), _p(out["keep"]), _p(out["norm_score"]), cap))
        for k in ("hyp_id", "dist", "angle", "corners"):
            out[k] = out[k][:nv.value]
        out["merged"] = out["merged"][:nm.value]
        out["keep"] = out["keep"][:nk.value]
        out["norm_score"] = out["norm_score"][:nk.value]
        return out

    # ---- BA half -------------------------------------------------------------------------------
    def ba_set_graph(self, cam_fixed, cube_fixed, ec=None, ep=None, eo=None):
        """ec=(cam,cube,meas10,info81)  ep=(cam,cube,meas4,info16,K9)  eo=(i,j,meas7,info36)"""
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dt); keep.append(a); return a.ctypes.data_as(C.c_void_p)
        g = BAGraph()
        g.n_cam, g.n_cube = len(cam_fixed), len(cube_fixed)
        self._ba_n_cam, self._ba_n_cube = len(cam_fixed), len(cube_fixed)
        g.cam_fixed, g.cube_fixed = arr(cam_fixed, np.int32), arr(cube_fixed, np.int32)
        if ec is not None and len(ec[0]):
            g.n_ec = len(ec[0]); g.ec_cam = arr(ec[0], np.int32); g.ec_cube = arr(ec[1], np.int32); g.ec_meas = arr(ec[2], np.float64); g.ec_info = arr(ec[3], np.float64)
        if ep is not None and len(ep[0]):
            g.n_ep = len(ep[0]); g.ep_cam = arr(ep[0], np.int32); g.ep_cube = arr(ep[1], np.int32); g.ep_meas = arr(ep[2], np.float64); g.ep_info = arr(ep[3], np.float64); g.ep_K = arr(ep[4], np.float64)
        if eo is not None and len(eo[0]):
            g.n_eo = len(eo[0]); g.eo_cam_i = arr(eo[0], np.int32); g.eo_cam_j = arr(eo[1], np.int32); g.eo_meas = arr(eo[2], np.float64); g.eo_info = arr(eo[3], np.float64)
        self._chk(lib().csb_ba_set_graph(self._h, C.byref(g)))
        self._ba_dims = (g.n_cam, g.n_cube, g.n_ec, g.n_ep, g.n_eo)

    def ba_add_frame(self, cam7, cam_fixed=False, new_cubes10=None, new_cube_fixed=None, ec=None, eo=None):
        """csb_ba_add_frame(): one more camera with its edges.  ec=(cube,meas10,info81)  eo=(cam_i,meas7,info36).  Returns the camera index."""
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dt); keep.append(a); return a.ctypes.data_as(C.c_void_p)
        f = BAFrame()
        f.cam7 = arr(cam7, np.float64); f.cam_fixed = int(bool(cam_fixed))
        if new_cubes10 is not None and len(new_cubes10):
            f.n_new_cubes = len(new_cubes10); f.new_cubes10 = arr(new_cubes10, np.float64)
            f.new_cube_fixed = arr(new_cube_fixed if new_cube_fixed is not None else np.zeros(len(new_cubes10)), np.int32)
        if ec is not None and len(ec[0]):
            f.n_ec = len(ec[0]); f.ec_cube = arr(ec[0], np.int32); f.ec_meas = arr(ec[1], np.float64); f.ec_info = arr(ec[2], np.float64)
        if eo is not None and len(eo[0]):
            f.n_eo = len(eo[0]); f.eo_cam_i = arr(eo[0], np.int32); f.eo_meas = arr(eo[1], np.float64); f.eo_info = arr(eo[2], np.float64)
        idx = C.c_int32(-1)
        self._chk(lib().csb_ba_add_frame(self._h, C.byref(f), C.byref(idx)))
        nc, nq, nec, nep, neo = self._ba_dims
        self._ba_dims = (nc + 1, nq + f.n_new_cubes, nec + f.n_ec, nep, neo + f.n_eo)
        self._ba_n_cam, self._ba_n_cube = self._ba_dims[0], self._ba_dims[1]
        return idx.value

    def _ba_out(self, jacobians):
        nc, nq, nec, nep, neo = self._ba_dims
        shapes = dict(ec_err=(nec, 9), ec_Ji=(nec, 54), ec_Jj=(nec, 81), ep_err=(nep, 4), ep_Ji=(nep, 24), ep_Jj=(nep, 36), eo_err=(neo, 6), eo_Ji=(neo, 36),
                      eo_Jj=(neo, 36), H_cam=(nc, 36), b_cam=(nc, 6), H_cube=(nq, 81), b_cube=(nq, 9), ec_Hij=(nec, 54), ep_Hij=(nep, 54), eo_Hij=(neo, 36), chi2=(1,))
        out = {k: np.zeros(v) for k, v in shapes.items()}
        O = BAOutput()
        for k in BA_OUT_FIELDS:
            if not jacobians and k.endswith(("_Ji", "_Jj")):
                continue
            setattr(O, k, _p(out[k]))
        return out, O

    def ba_linearize(self, cams7, cubes10, jacobians=True):
        cams7 = np.ascontiguousarray(cams7, np.float64); cubes10 = np.ascontiguousarray(cubes10, np.float64)
        out, O = self._ba_out(jacobians)
        self._chk(lib().csb_ba_linearize(self._h, _p(cams7), _p(cubes10), C.byref(O)))
        return out

    def ba_upload_estimates(self, cams7, cubes10):
        cams7 = np.ascontiguousarray(cams7, np.float64); cubes10 = np.ascontiguousarray(cubes10, np.float64)
        self._chk(lib().csb_ba_upload_estimates(self._h, _p(cams7), _p(cubes10)))

    def ba_set_jacobian_mode(self, analytic):
        """csb_ba_set_jacobian_mode(): False = central differences like the reference (default), True = closed form."""
        self._chk(lib().csb_ba_set_jacobian_mode(self._h, 1 if analytic else 0))

    def ba_run(self):
        self._chk(lib().csb_ba_run(self._h))

    def ba_optimize(self, iterations):
        """csb_ba_optimize(): LM on the device from the uploaded estimates.  Returns (cams7, cubes10, stats)."""
        cams = np.zeros((self._ba_n_cam, 7)); cubes = np.zeros((self._ba_n_cube, 10))
        st = BAOptimizeStats()
        self._chk(lib().csb_ba_optimize(self._h, int(iterations), _p(cams), _p(cubes), C.byref(st)))
        return cams, cubes, st

    def ba_download(self, jacobians=False):
        out, O = self._ba_out(jacobians)
        self._chk(lib().csb_ba_download(self._h, C.byref(O)))
        return out

    # ---- line detection (LSD) ------------------------------------------------------------------
    def _lsd_params(self, line_length_thres, filter, max_lines, unit_link_deg=0):
        self._lsd_p = LsdParams(float(line_length_thres), int(filter), int(max_lines), int(unit_link_deg))
        return self._lsd_p

    @staticmethod
    def _lsd_gray(gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        if gray.ndim == 2:
            gray = gray[None]
        assert gray.ndim == 3
        return gray

    def lsd_detect_batch(self, gray, line_length_thres=15.0, filter=True, max_lines=4096, unit_link_deg=0):
        """csb_lsd_detect_batch(): gray (n, h, w) uint8 -> list of (k_i, 4) float32 arrays [x1 y1 x2 y2], stats."""
        gray = self._lsd_gray(gray)
        n, h, w = gray.shape
        P = self._lsd_params(line_length_thres, filter, max_lines, unit_link_deg)
        lines = np.zeros((n, max_lines, 4), np.float32); cnt = np.zeros(n, np.int32); st = LsdStats()
        self._chk(lib().csb_lsd_detect_batch(self._h, _p(gray), n, w, h, C.byref(P), _p(lines), _p(cnt), C.byref(st)))
        return [lines[i, :cnt[i]].copy() for i in range(n)], st

    def lsd_upload(self, gray, line_length_thres=15.0, filter=True, max_lines=4096, unit_link_deg=0):
        gray = self._lsd_gray(gray)
        n, h, w = gray.shape
        self._lsd_n = n
        self._chk(lib().csb_lsd_upload(self._h, _p(gray), n, w, h, C.byref(self._lsd_params(line_length_thres, filter, max_lines, unit_link_deg))))

    def lsd_run(self, timed=False):
        self._chk(lib().csb_lsd_run(self._h, int(timed)))

    def lsd_download(self):
        n, cap = self._lsd_n, self._lsd_p.max_lines
        lines = np.zeros((n, cap, 4), np.float32); cnt = np.zeros(n, np.int32); st = LsdStats()
        self._chk(lib().csb_lsd_download(self._h, _p(lines), _p(cnt), C.byref(st)))
        return [lines[i, :cnt[i]].copy() for i in range(n)], st

    def lsd_debug_maps(self, frame, scaled_shape):
        H, W = scaled_shape
        sc = np.zeros((H, W)); mg = np.zeros((H, W)); an = np.zeros((H, W))
        self._chk(lib().csb_lsd_debug_maps(self._h, int(frame), _p(sc), _p(mg), _p(an)))
        return sc, mg, an

    # ---- line detection (EDLines, use_LSD = false) ----------------------------------------------
    def edlines_detect_batch(self, gray, line_length_thres=15.0, filter=True, max_lines=4096):
        """csb_edlines_detect_batch(): gray (n, h, w) uint8 -> list of (k_i, 4) float32 arrays [x1 y1 x2 y2], stats."""
        gray = self._lsd_gray(gray)
        n, h, w = gray.shape
        P = LsdParams(float(line_length_thres), int(filter), int(max_lines), 0)
        lines = np.zeros((n, max_lines, 4), np.float32); cnt = np.zeros(n, np.int32); st = EdlinesStats()
        self._chk(lib().csb_edlines_detect_batch(self._h, _p(gray), n, w, h, C.byref(P), _p(lines), _p(cnt), C.byref(st)))
        return [lines[i, :cnt[i]].copy() for i in range(n)], st

    def edlines_detect_describe_batch(self, gray, line_length_thres=15.0, filter=True, max_lines=4096, want_float=False):
        """csb_edlines_upload / run / describe / download*: lines and their LBD descriptors (detect_descrip_lines with use_LSD = false)."""
        gray = self._lsd_gray(gray)
        n, h, w = gray.shape
        P = LsdParams(float(line_length_thres), int(filter), int(max_lines), 0)
        self._chk(lib().csb_edlines_upload(self._h, _p(gray), n, w, h, C.byref(P)))
        self._chk(lib().csb_edlines_run(self._h, 0))
        self._chk(lib().csb_edlines_describe(self._h, int(want_float)))
        lines = np.zeros((n, max_lines, 4), np.float32); cnt = np.zeros(n, np.int32); st = EdlinesStats()
        self._chk(lib().csb_edlines_download(self._h, _p(lines), _p(cnt), C.byref(st)))
        total = int(cnt.sum())
        d32 = np.zeros((max(total, 1), 32), np.uint8); d72 = np.zeros((max(total, 1), 72), np.float32) if want_float else None
        cnt2 = np.zeros(n, np.int32)
        self._chk(lib().csb_edlines_download_descriptors(self._h, _p(d32), _p(d72), _p(cnt2), C.c_int64(max(total, 1))))
        out = {"lines": [lines[i, :cnt[i]].copy() for i in range(n)], "desc": self._lbd_split(d32[:total], cnt), "stats": st}
        if want_float:
            out["desc_float"] = self._lbd_split(d72[:total], cnt)
        return out

    # ---- line descriptors (LBD) ----------------------------------------------------------------
    @staticmethod
    def _lbd_pack(lines_per_frame):
        off = np.zeros(len(lines_per_frame) + 1, np.int32)
        for i, l in enumerate(lines_per_frame):
            off[i + 1] = off[i] + len(l)
        flat = np.zeros((max(int(off[-1]), 1), 4), np.float32)
        for i, l in enumerate(lines_per_frame):
            if len(l):
                flat[off[i]:off[i + 1]] = np.asarray(l, np.float32).reshape(-1, 4)
        return flat, off

    @staticmethod
    def _lbd_split(arr, cnt):
        out, o = [], 0
        for k in cnt:
            out.append(arr[o:o + k].copy()); o += k
        return out

    def lbd_describe_batch(self, gray, lines_per_frame, want_float=False):
        """csb_lbd_describe_batch(): gray (n, h, w) uint8 + per-frame (k_i, 4) float32 line arrays -> per-frame (k_i, 32) uint8 descriptors
        (and (k_i, 72) float32 ones with want_float), stats."""
        gray = self._lsd_gray(gray)
        n, h, w = gray.shape
        assert len(lines_per_frame) == n
        flat, off = self._lbd_pack(lines_per_frame)
        total = int(off[-1])
        d32 = np.zeros((max(total, 1), 32), np.uint8); d72 = np.zeros((max(total, 1), 72), np.float32) if want_float else None
        st = LbdStats()
        self._chk(lib().csb_lbd_describe_batch(self._h, _p(gray), n, w, h, _p(flat), _p(off), _p(d32), _p(d72), C.byref(st)))
        cnt = np.diff(off)
        if want_float:
            return self._lbd_split(d32[:total], cnt), self._lbd_split(d72[:total], cnt), st
        return self._lbd_split(d32[:total], cnt), st

    def lbd_upload(self, gray, lines_per_frame, want_float=False):
        gray = self._lsd_gray(gray)
        n, h, w = gray.shape
        flat, off = self._lbd_pack(lines_per_frame)
        self._lbd_cap = int(off[-1])
        self._lbd_n = n
        self._chk(lib().csb_lbd_upload(self._h, _p(gray), n, w, h, _p(flat), _p(off), int(want_float)))

    def lbd_run(self, timed=False):
        self._chk(lib().csb_lbd_run(self._h, int(timed)))

    def lbd_run_on_lsd(self, want_float=False, timed=False):
        """csb_lbd_run_on_lsd(): descriptors of the segments the last lsd_run() left on the device."""
        self._lbd_cap = self._lsd_n * self._lsd_p.max_lines
        self._lbd_n = self._lsd_n
        self._chk(lib().csb_lbd_run_on_lsd(self._h, int(want_float), int(timed)))

    def lbd_download(self, want_float=False, keylines=False):
        cap = max(self._lbd_cap, 1)
        d32 = np.zeros((cap, 32), np.uint8)
        d72 = np.zeros((cap, 72), np.float32) if want_float else None
        kl = np.zeros((cap, 4), np.float32) if keylines else None
        cnt = np.zeros(self._lbd_n, np.int32); st = LbdStats()
        self._chk(lib().csb_lbd_download(self._h, _p(d32), _p(d72), _p(kl), _p(cnt), C.c_int64(cap), C.byref(st)))
        out = {"desc": self._lbd_split(d32, cnt), "stats": st, "n_lines": cnt}
        if want_float:
            out["desc_float"] = self._lbd_split(d72, cnt)
        if keylines:
            out["keylines"] = self._lbd_split(kl, cnt)
        return out

    def lbd_debug_gradients(self, frame, shape):
        h, w = shape
        dx = np.zeros((h, w), np.int16); dy = np.zeros((h, w), np.int16)
        self._chk(lib().csb_lbd_debug_gradients(self._h, int(frame), _p(dx), _p(dy)))
        return dx, dy


class detect_3d_cuboid:
    """Host-side mirror of class detect_3d_cuboid (reference: detect_3d_cuboid/include/detect_3d_cuboid/detect_3d_cuboid.h:74-118): same
    member names, defaults and meaning -- set_calibration, the configuration flags, detect_cuboid(img, transToWolrd, obj_bbox_coors, edges)
    -- computing on the GPU through csb_detect_batch_gray (Canny + distance transform + the proposal sweep + scoring + ranking + 3D
    recovery).  The plotting / saving switches of the reference are accepted and ignored (no images are produced)."""

    def __init__(self, ctx):
        self._ctx = ctx
        self.Kalib = None
        self.whether_plot_detail_images = self.whether_plot_final_images = self.whether_save_final_images = self.print_details = False
        self.consider_config_1 = True                # detect_3d_cuboid.h:109-116
        self.consider_config_2 = True
        self.whether_sample_cam_roll_pitch = True
        self.whether_sample_bbox_height = False
        self.max_cuboid_num = 1
        self.nominal_skew_ratio = 1.0
        self.max_cut_skew = 3.0

    def set_calibration(self, Kalib):
        self.Kalib = np.asarray(Kalib, np.float64).reshape(3, 3).copy()

    def params(self):
        return DetectParams(int(self.consider_config_1), int(self.consider_config_2), int(self.whether_sample_cam_roll_pitch),
                            int(self.whether_sample_bbox_height), int(self.max_cuboid_num), 0, float(self.nominal_skew_ratio), float(self.max_cut_skew))

    def detect_cuboid(self, img, transToWolrd, obj_bbox_coors, edges):
        """box_proposal_detail.cpp:65-861.  img: (h, w) uint8 gray or (h, w, 3) BGR (converted like cv::cvtColor BGR2GRAY); transToWolrd: 4x4
        camera-to-world; obj_bbox_coors: (k, 5) x y w h prob, 0-based; edges: (n, 4) x1 y1 x2 y2.  Returns all_object_cuboids: one list per
        2D box holding up to max_cuboid_num Cuboid records, best first (empty list: no valid proposal for that box)."""
        if self.Kalib is None:
            raise CsbError(CSB_ERR_STATE, "detect_3d_cuboid: set_calibration() has not been called")
        img = np.asarray(img)
        if img.ndim == 3:   # cv::cvtColor(BGR2GRAY), 8-bit fixed point of OpenCV 4.x: (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15
            b, g, r = (img[..., i].astype(np.int32) for i in range(3))   # (bit-identical to cv2 4.13, tests/test_node.py)
            img = ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)
        if img.ndim != 2 or img.dtype != np.uint8:
            raise CsbError(CSB_ERR_INVALID, "detect_3d_cuboid: img must be uint8, (h, w) or (h, w, 3)")
        h, w = img.shape
        boxes = np.ascontiguousarray(obj_bbox_coors, np.float64).reshape(-1, 5)
        lines = np.ascontiguousarray(edges, np.float64).reshape(-1, 4) if len(edges) else np.zeros((0, 4))
        if len(boxes) == 0:
            return []
        p = self.params()
        frames = make_frames([self.Kalib], [np.asarray(transToWolrd, np.float64).reshape(4, 4)], w, h, [(0, len(boxes))], [(0, len(lines))])
        tasks, n_tasks, _ = detect_plan(frames, boxes, p)
        cub, ncub, _ = self._ctx.detect_batch_gray(frames, boxes, lines, tasks, n_tasks, np.ascontiguousarray(img.ravel()), p)
        k = max(1, int(self.max_cuboid_num))
        return [[cub[b * k + j] for j in range(int(ncub[b]))] for b in range(len(boxes))]


class line_lbd_detect:
    """Host-side mirror of class line_lbd_detect (reference: line_lbd/include/line_lbd/line_lbd_allclass.h:20-60) for its LSD branch:
    same member names and meaning (use_LSD, line_length_thres, detect_filter_lines), computing on the GPU through the C ABI."""

    def __init__(self, ctx, numoctaves=1, octaveratio=2.0):
        if numoctaves != 1:
            raise CsbError(CSB_ERR_INVALID, "only one octave is supported (every caller in the reference uses one: main_obj.cpp:503, detect_lines.cpp:61)")
        self._ctx = ctx
        self.use_LSD = True            # line_lbd_allclass.cpp:125 defaults to False (EDLines)
        self.line_length_thres = 50.0  # line_lbd_allclass.cpp:126; both callers overwrite it with 15
        self.max_lines = 4096

    def detect_filter_lines(self, gray_img):
        """line_lbd_allclass.cpp:221-235: gray image(s) -> linesmat_out rows [x1 y1 x2 y2] float32 (one array per frame)."""
        single = np.asarray(gray_img).ndim == 2
        if not self.use_LSD:
            out, _ = self._ctx.edlines_detect_batch(gray_img, self.line_length_thres, True, self.max_lines)
            return out[0] if single else out
        out, _ = self._ctx.lsd_detect_batch(gray_img, self.line_length_thres, True, self.max_lines)
        return out[0] if single else out

    def detect_descrip_lines(self, gray_img):
        """line_lbd_allclass.cpp:263-281 (the KeyLine overload: octave 0, lineLength > line_length_thres): gray image(s) -> (lines rows
        [x1 y1 x2 y2] float32, line_descrips rows of 32 bytes); the segments stay on the device between the two stages."""
        single = np.asarray(gray_img).ndim == 2
        c = self._ctx
        if not self.use_LSD:
            out = c.edlines_detect_describe_batch(gray_img, self.line_length_thres, True, self.max_lines)
            return (out["lines"][0], out["desc"][0]) if single else (out["lines"], out["desc"])
        c.lsd_upload(gray_img, self.line_length_thres, True, self.max_lines)
        c.lsd_run()
        c.lbd_run_on_lsd()
        lines, _ = c.lsd_download()
        desc = c.lbd_download()["desc"]
        return (lines[0], desc[0]) if single else (lines, desc)
