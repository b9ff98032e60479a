// observe.cu -- per-box observation records for camera-object graph assembly (the producer side of the one NCCL allgather).
//
// Follows object_slam/src/main_obj.cpp:643-679 and :732: the best cuboid of a 2D box (ground frame) becomes a g2o::cuboid
// measurement in the local camera frame, plus meas_quality = (1 - normalized_error + 0.5) / 2.
// Record layout (16 doubles): frame_id, box_id, valid, meas_quality, x y z qx qy qz qw sx sy sz, normalized_error, 0.
#include <cuda_runtime.h>

#include "context.h"
#include "csb_math.cuh"

namespace csb {

__global__ void k_observe(DetectBuffers B, int n_boxes, double* out) {
    int box = blockIdx.x * blockDim.x + threadIdx.x;
    if (box >= n_boxes) return;
    double* r = out + 16 * (size_t)box;
    const int t0 = B.box_task_begin[box], t1 = B.box_task_begin[box + 1];
    for (int i = 0; i < 16; i++) r[i] = 0;
    r[1] = (double)box;
    r[0] = -1;
    if (t1 > t0) r[0] = (double)B.ttab[t0].frame_id;
    if (B.n_cuboids[box] < 1) return;
    const csb_cuboid& c = B.cuboids[(size_t)box * B.dc.max_cuboid_num];
    const FrameTab& ft = B.ftab[B.ttab[t0].frame_id];
    // cube_ground_value.fromMinimalVector([pos, 0, 0, rotY, scale])  (g2o_Object.h:36-41; zyx_euler_to_quat matrix_utils.cpp:19-33)
    double sy = sin(c.rotY * 0.5), cy = cos(c.rotY * 0.5), sp = sin(0.0 * 0.5), cp = cos(0.0 * 0.5), sr = sin(0.0 * 0.5), cr = cos(0.0 * 0.5);
    Quat q{cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy};
    SE3 cube_pose = se3_make(q, V3{c.pos[0], c.pos[1], c.pos[2]});
    // camera pose the measurement is expressed in: the sampled roll/pitch applied to the raw pose (main_obj.cpp:667-676)
    M3 R;
    const double* T0 = ft.Tnew[0];
    if (ft.sample_rp) R = euler_zyx_to_rot(ft.euler_raw[0] + c.camera_roll_delta, ft.euler_raw[1] + c.camera_pitch_delta, ft.euler_raw[2]);
    else R = M3{{T0[0], T0[1], T0[2], T0[4], T0[5], T0[6], T0[8], T0[9], T0[10]}};
    SE3 Twc = se3_make(quat_from_rot(R), V3{T0[3], T0[7], T0[11]});
    SE3 local = se3_mul(se3_inverse(Twc), cube_pose);  // cuboid::transform_to, g2o_Object.h:126-132
    r[2] = 1.0;
    r[3] = (1 - c.normalized_error + 0.5) / 2;
    r[4] = local.t.x; r[5] = local.t.y; r[6] = local.t.z; r[7] = local.r.x; r[8] = local.r.y; r[9] = local.r.z; r[10] = local.r.w;
    r[11] = c.scale[0]; r[12] = c.scale[1]; r[13] = c.scale[2];
    r[14] = c.normalized_error;
}

}  // namespace csb

extern "C" int csb_detect_observations_device(csb_context* c, void* dev_out) {
    if (!c || !dev_out) return CSB_ERR_INVALID;
    DetectState& d = c->det;
    if (!d.ran) { c->err = "csb_detect_observations_device before csb_detect_run"; return CSB_ERR_STATE; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    if (d.n_boxes == 0) return CSB_OK;
    csb::k_observe<<<(d.n_boxes + 127) / 128, 128, 0, c->stream>>>(d.B, d.n_boxes, (double*)dev_out);
    CSB_CUDA(c, cudaGetLastError());
    return CSB_OK;
}
