"""Host-side graph assembly from observation records: the caller's bookkeeping between the two halves of the path (BASELINE config #5).

Mirrors the graph recipe of the reference's object_slam node (object_slam/src/main_obj.cpp:738-803) for the records
`csb_detect_observations_device` writes (csrc/observe.cu: 16 doubles per 2D box = frame, box, valid, meas_quality, the cuboid measurement in
the local camera frame x y z qx qy qz qw sx sy sz, normalized_error, 0), gathered over the ranks:

  * one VertexSE3Expmap per frame holding the world->camera pose, the first one fixed (:753-762);
  * one VertexCuboid per landmark, initialised from its first valid observation: cube_local_meas.transform_from(Twc) (:745-751;
    g2o_Object.h:134-140).  The reference's data set has a single landmark; here column 1 of a record names the landmark (data
    association is the caller's: it is not part of the reference);
  * one EdgeSE3Cuboid per valid record, information diag((2 meas_quality)^2) (:766-781);
  * one EdgeSE3Expmap between consecutive frames, information I6 (:786-799); its measurement is T_j T_i^-1 of the given poses.

Pure numpy (vectorised); the result goes to csb_ba_set_graph / csb_ba_linearize / csb_ba_optimize.  Nothing here computes on the path.
"""
import numpy as np


def _qmul(a, b):  # (..., 4) x y z w
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _qrot(q, v):
    qv = np.concatenate([v, np.zeros(v.shape[:-1] + (1,))], axis=-1)
    qc = q * np.array([-1.0, -1.0, -1.0, 1.0])
    return _qmul(_qmul(q, qv), qc)[..., :3]


def _wpos(q):
    """g2o::SE3Quat keeps w >= 0 (normalizeRotation: se3quat.h:58-70 and after every product / inverse)"""
    return np.where(q[..., 3:4] < 0, -q, q)


def se3_mul(a, b):
    """(..., 7) x y z qx qy qz qw"""
    return np.concatenate([a[..., :3] + _qrot(a[..., 3:7], b[..., :3]), _wpos(_qmul(a[..., 3:7], b[..., 3:7]))], axis=-1)


def se3_inv(a):
    qc = _wpos(a[..., 3:7] * np.array([-1.0, -1.0, -1.0, 1.0]))
    return np.concatenate([-_qrot(qc, a[..., :3]), qc], axis=-1)


def rot_to_quat(R):
    """Eigen::Quaterniond(Matrix3d) (the branch structure of Eigen's quaternion-from-matrix); (3, 3) -> x y z w."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([(R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s, w])
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[i] = 0.5 * s
    s = 0.5 / s
    q[3] = (R[k, j] - R[j, k]) * s
    q[j] = (R[j, i] + R[i, j]) * s
    q[k] = (R[k, i] + R[i, k]) * s
    return q


def pose7_from_matrix(T):
    T = np.asarray(T, np.float64)
    return np.concatenate([T[:3, 3], rot_to_quat(T[:3, :3])])


def globalise_records(rec, frames_per_step, boxes_per_rank):
    """rec: (world, steps, boxes_per_rank, 16) records as gathered over the ranks (rank-major), every rank having pushed `steps` passes over
    its own `frames_per_step` frames.  Rewrites, in place, column 0 to the global frame index ((rank * steps + step) * frames_per_step + local
    frame; -1 stays -1) and column 1 to the global landmark index (rank * boxes_per_rank + box: every 2D box of a rank's frames is its own
    landmark, observed again at every pass).  Returns (records as (n, 16), number of landmarks)."""
    world, steps = rec.shape[0], rec.shape[1]
    step_base = (np.arange(world)[:, None] * steps + np.arange(steps)[None, :]) * frames_per_step
    has_frame = rec[..., 0] >= 0
    rec[..., 0] = np.where(has_frame, rec[..., 0] + step_base[:, :, None], -1)
    rec[..., 1] = rec[..., 1] + (np.arange(world) * boxes_per_rank)[:, None, None]
    return rec.reshape(-1, 16), world * boxes_per_rank


def assemble_graph(records, cams_wc7, n_landmarks, cams_est_wc7=None):
    """records: (n, 16) observation records whose column 0 already holds the GLOBAL frame index (row of cams_wc7) and whose column 1
    holds the landmark index (taken modulo n_landmarks).  cams_wc7: (n_frames, 7) camera-to-world poses: they give the odometry
    measurements.  cams_est_wc7 (default: the same poses): the camera ESTIMATES the vertices start from and the landmarks are initialised
    with (the node uses its current estimate there, main_obj.cpp:745-751).  Returns the dict of arrays Context.ba_set_graph / ba_linearize
    take, plus 'landmark_seen'."""
    records = np.asarray(records, np.float64).reshape(-1, 16)
    cams_true = np.asarray(cams_wc7, np.float64).reshape(-1, 7)
    cams_wc7 = cams_true if cams_est_wc7 is None else np.asarray(cams_est_wc7, np.float64).reshape(-1, 7)
    n_frames = len(cams_wc7)
    cams_cw = se3_inv(cams_wc7)                    # g2o vertices store world -> camera (main_obj.cpp:760)
    true_cw = se3_inv(cams_true)
    valid = (records[:, 2] == 1) & (records[:, 0] >= 0) & (records[:, 0] < n_frames)
    rec = records[valid]
    order = np.lexsort((rec[:, 1], rec[:, 0]))     # frame after frame, box after box: the order the node adds its edges in
    rec = rec[order]
    frame = rec[:, 0].astype(np.int64)
    lm = rec[:, 1].astype(np.int64) % n_landmarks
    meas = rec[:, 4:14]
    q = rec[:, 3]
    # landmark initialisation from the first valid observation (frame order)
    cubes = np.zeros((n_landmarks, 10))
    cubes[:, 6] = 1.0
    cubes[:, 7:10] = 1.0
    seen = np.zeros(n_landmarks, bool)
    if len(lm):
        ul, fi = np.unique(lm, return_index=True)  # first occurrence per landmark in frame order
        cubes[ul, :7] = se3_mul(cams_wc7[frame[fi]], meas[fi, :7])   # cuboid::transform_from: pose = Twc * local pose, scale unchanged
        cubes[ul, 7:10] = meas[fi, 7:10]
        seen[ul] = True
    info = np.zeros((len(rec), 81))
    info[:, ::10] = ((2.0 * q) ** 2)[:, None]
    ec = (frame.astype(np.int32), lm.astype(np.int32), np.ascontiguousarray(meas), info)
    if n_frames > 1:
        i0 = np.arange(n_frames - 1, dtype=np.int32)
        odo = se3_mul(true_cw[1:], se3_inv(true_cw[:-1]))     # e = log(M T_i T_j^-1) = 0 for M = T_j T_i^-1 (types_six_dof_expmap.h:90-99)
        eo = (i0, i0 + 1, np.ascontiguousarray(odo), np.tile(np.eye(6).ravel(), (n_frames - 1, 1)))
    else:
        eo = None
    cam_fixed = np.zeros(n_frames, np.int32)
    cam_fixed[0] = 1
    return dict(cams7=np.ascontiguousarray(cams_cw), cubes10=cubes, cam_fixed=cam_fixed, cube_fixed=(~seen).astype(np.int32), ec=ec, ep=None, eo=eo,
                landmark_seen=seen)
