/* cubeslam_b200.h -- C ABI of the B200-native CubeSLAM hot path.
 *
 * Drop-in boundary.  Two halves:
 *
 *  (1) proposal half -- everything inside detect_3d_cuboid::detect_cuboid()
 *      (reference: detect_3d_cuboid/include/detect_3d_cuboid/detect_3d_cuboid.h:74-118,
 *       detect_3d_cuboid/src/box_proposal_detail.cpp:65-861).  cv::Canny + cv::distanceTransform
 *      (box_proposal_detail.cpp:320-327) run on the GPU when the caller hands over the gray frames
 *      (csb_detect_batch_gray / csb_detect_upload_gray); a caller that computes the distance maps itself
 *      asks csb_detect_plan() which ROIs need one and passes them to csb_detect_batch().
 *
 *  (2) BA half -- what g2o's BlockSolver::buildSystem() does for the cuboid graph
 *      (reference: object_slam/Thirdparty/g2o/g2o/core/block_solver.hpp:501-560 calling
 *       base_binary_edge.hpp:54-120,130-205 on the edge/vertex classes of
 *       object_slam/include/object_slam/g2o_Object.h:202-292 and types_six_dof_expmap.h:59-142).
 *
 * All entry points take plain pointers and sizes, return CSB_OK (0) or a negative error code, and never
 * throw.  Buffers are caller-owned.  A context is bound to one CUDA device and one stream; it is not
 * re-entrant (like the reference's detect_3d_cuboid object, which mutates cam_pose during the sweep).
 * There is no CPU fallback: every entry point that computes fails with CSB_ERR_CUDA if no device is usable.
 */
#ifndef CUBESLAM_B200_H
#define CUBESLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSB_OK 0
#define CSB_ERR_INVALID (-1)  /* bad argument / inconsistent sizes */
#define CSB_ERR_CUDA (-2)     /* CUDA runtime error (see csb_last_error) */
#define CSB_ERR_CAPACITY (-3) /* an output buffer or internal cap is too small */
#define CSB_ERR_STATE (-4)    /* call order violated (e.g. run before upload) */

typedef struct csb_context csb_context;

int csb_create(csb_context** out, int device_ordinal);
void csb_destroy(csb_context* ctx);
const char* csb_last_error(const csb_context* ctx);
/* Use an existing CUDA stream (cudaStream_t) for all work of this context; NULL = context-owned stream. */
int csb_set_stream(csb_context* ctx, void* cuda_stream);
int csb_synchronize(csb_context* ctx);
const char* csb_version(void);
/* Tuning switches (defaults in brackets).
 *   CSB_OPT_GRAY_GATHER [1]: csb_detect_upload_gray / csb_detect_batch_gray with a PINNED, 16-byte aligned gray buffer (cudaHostAlloc /
 *   cudaHostRegister) fetch only the 128-byte segments of the frames that the box ROIs (+ the one-pixel Sobel halo) touch, by a kernel
 *   reading the caller's buffer over PCIe, instead of copying whole frames with the copy engine; 0 = always copy whole frames.  Results
 *   are identical either way; pageable buffers are always copied as a whole. */
#define CSB_OPT_GRAY_GATHER 1
int csb_set_option(csb_context* ctx, int option, int value);

/* ------------------------------------------------------------------------------------------------
 * Proposal half
 * ------------------------------------------------------------------------------------------------ */

/* Public mode flags of class detect_3d_cuboid (detect_3d_cuboid.h:95-117).  Plot/print flags have no
 * meaning here and are handled (ignored) by the C++ adapter. */
typedef struct csb_detect_params {
    int32_t consider_config_1;             /* :109 default 1 */
    int32_t consider_config_2;             /* :110 default 1 */
    int32_t whether_sample_cam_roll_pitch; /* :111 default 1 */
    int32_t whether_sample_bbox_height;    /* :112 default 0 */
    int32_t max_cuboid_num;                /* :114 default 1 */
    int32_t reserved;
    double nominal_skew_ratio; /* :115 default 1 */
    double max_cut_skew;       /* :116 default 3 */
} csb_detect_params;

/* One detect_cuboid() call = one frame: calibration (set_calibration), camera pose (transToWolrd, row-major
 * 4x4), image size, and the row ranges of this frame's 2D boxes and line segments in the batch arrays. */
typedef struct csb_frame {
    double Kalib[9];
    double transToWolrd[16];
    int32_t img_width, img_height;
    int32_t box_begin, box_end;   /* rows of boxes[] : x y w h prob (0-based pixels), obj_bbox_coors */
    int32_t line_begin, line_end; /* rows of lines[] : x1 y1 x2 y2, all_lines_raw */
} csb_frame;

/* One (2D box, height sample) unit of work.  The caller must supply a float32 distance map of
 * roi_height x roi_width for each task, computed as the reference does:
 *   cv::Canny(gray(cv::Rect(roi_left, roi_top, roi_width, roi_height)), e, 80, 200);
 *   cv::distanceTransform(255 - e, dist_map, CV_DIST_L2, 3);
 * packed at dist_maps[map_offset .. map_offset + roi_width*roi_height). */
typedef struct csb_task {
    int32_t frame_id, box_id; /* box_id indexes boxes[] (batch-global row) */
    int32_t hs_id, down_expand;
    int32_t roi_left, roi_top, roi_width, roi_height;
    int32_t n_top, n_enum; /* top-edge samples; hypotheses enumerated for this task */
    int64_t map_offset;    /* in floats */
} csb_task;

/* Mirror of class cuboid (detect_3d_cuboid.h:20-41), POD. */
typedef struct csb_cuboid {
    double pos[3];
    double scale[3];
    double rotY;
    double box_config_type[2];       /* configuration id, vp1 left(1)/right(2) */
    double box_corners_3d_world[24]; /* 3x8 row-major */
    double rect_detect_2d[4];
    double edge_distance_error, edge_angle_error, normalized_error, skew_ratio;
    double down_expand_height, camera_roll_delta, camera_pitch_delta;
    int32_t box_corners_2d[16]; /* 2x8 row-major */
    int32_t task_id;            /* provenance: task (index into tasks[]) ... */
    int32_t raw_cube_ind;       /* ... and index in that task's compacted valid-proposal list */
    int32_t rank_index;         /* index in the box's raw_obj_proposals list (what sort_idx_small holds) */
    int32_t reserved;
} csb_cuboid;

typedef struct csb_detect_stats {
    int64_t n_enumerated; /* hypotheses enumerated */
    int64_t n_scored;     /* hypotheses that passed all geometric checks and received both scores */
    int64_t n_kept;       /* proposals that survived fuse_normalize_scores_v2 */
    int64_t h2d_bytes, d2h_bytes;
    int32_t n_kernel_launches, n_tasks_smem_map; /* tasks whose distance map was staged in shared memory */
    float gpu_ms_prep, gpu_ms_score, gpu_ms_select, gpu_ms_recover, gpu_ms_rank, gpu_ms_distmap; /* CUDA-event times of the last timed run */
} csb_detect_stats;

/* Host-only integer logic of box_proposal_detail.cpp:143-256: enumerates tasks and ROI rectangles.
 * tasks_out may be NULL to query the count. */
int csb_detect_plan(const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const csb_detect_params* params,
                    csb_task* tasks_out, int max_tasks, int* n_tasks_out, int64_t* n_map_floats_out);

/* The batched equivalent of calling detect_cuboid() once per frame, with HOST buffers (copies in and out
 * happen inside).  cuboids_out holds n_boxes * max_cuboid_num entries (box-major, best first),
 * n_cuboids_out[n_boxes] the number filled per box (0 = empty ObjectSet).  stats may be NULL. */
int csb_detect_batch(csb_context* ctx, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines,
                     int n_lines, const csb_task* tasks, int n_tasks, const float* dist_maps, int64_t n_map_floats,
                     const csb_detect_params* params, csb_cuboid* cuboids_out, int32_t* n_cuboids_out, csb_detect_stats* stats);

/* Device-resident variant: upload once, run the kernels any number of times, download results.
 * csb_detect_run() is asynchronous on the context stream; timed != 0 records per-kernel CUDA events. */
int csb_detect_upload(csb_context* ctx, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines,
                      int n_lines, const csb_task* tasks, int n_tasks, const float* dist_maps, int64_t n_map_floats,
                      const csb_detect_params* params);
int csb_detect_run(csb_context* ctx, int timed);
int csb_detect_download(csb_context* ctx, csb_cuboid* cuboids_out, int32_t* n_cuboids_out, csb_detect_stats* stats);

/* Gray-frame variants: instead of caller-computed distance maps the caller passes the 8-bit gray frames
 * (img_height x img_width bytes each, packed in frame order).  Every csb_detect_run() then starts with, per task ROI,
 *   cv::Canny(gray(roi), e, 80, 200)  +  cv::distanceTransform(255 - e, dist, CV_DIST_L2, 3)      (box_proposal_detail.cpp:320-327)
 * computed on the device with OpenCV's own algorithms (Sobel 3x3 that sees the ROI's true neighbours like a cv::Mat ROI view,
 * L1 magnitude, NMS, hysteresis; 16.16 fixed-point 3x3 chamfer transform) -- bit-identical to cv2 4.13 with IPP disabled.
 * One CTA per ROI: 256 threads in a batch, 1024 when the call holds at most one ROI per SM (one frame per call: latency); the environment
 * variable CSB_DISTMAP_CTA=256|1024 forces one of the two (tests). */
int csb_detect_upload_gray(csb_context* ctx, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines,
                           int n_lines, const csb_task* tasks, int n_tasks, const uint8_t* gray, int64_t n_gray_bytes,
                           const csb_detect_params* params);
int csb_detect_batch_gray(csb_context* ctx, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines,
                          int n_lines, const csb_task* tasks, int n_tasks, const uint8_t* gray, int64_t n_gray_bytes,
                          const csb_detect_params* params, csb_cuboid* cuboids_out, int32_t* n_cuboids_out, csb_detect_stats* stats);
/* Parity/debug: the distance map (and, in gray mode, the 0/1/2 Canny map: 2 = edge) of one task after a run. */
int csb_detect_debug_map(csb_context* ctx, int task_id, float* dist_map_out, uint8_t* edges_out, int capacity);
/* Profiling/debug, only in a library built with -DCSB_SCORE_PHASES (otherwise CSB_ERR_CUDA, "not supported"): SM cycles spent per phase
 * of the scoring kernel (thread 0 of every CTA, summed over CTAs and tasks) since the last reset: [0] task fetch / chunk wait, [1] line
 * tables + vanishing points, [2] VP support, [3] corner construction + rejection, [4] prefix sums, [5] wait for the distance map,
 * [6] scoring, [7] exit; [8..10] VP-support units decided by the float / double / exact tier, [11] unused.  The buffer holds 12 entries. */
int csb_detect_debug_score_phases(csb_context* ctx, uint64_t* cycles12, int reset);
/* Parity/debug: the six-at-a-time atan2 of the scoring kernel (groups of six operands, n a multiple of 6) with its scalar fallback;
 * n_fallback (optional) receives the number of groups that took the fallback.  Must equal the scalar det_atan2 bit for bit. */
int csb_detect_debug_atan2(csb_context* ctx, const double* y, const double* x, double* out, int n, int* n_fallback);

/* Per-box observation records for camera-object graph assembly (object_slam/src/main_obj.cpp:643-679, :732): the best
 * cuboid of each 2D box as a g2o::cuboid measurement in the local camera frame.  Writes n_boxes x 16 doubles into a
 * DEVICE buffer (asynchronously, on the context stream) so that the records can be all-gathered across GPUs (NCCL)
 * without a host round trip:  frame_id, box_id, valid, meas_quality, x y z qx qy qz qw sx sy sz, normalized_error, 0. */
int csb_detect_observations_device(csb_context* ctx, void* device_out);

/* Parity/debug access to the intermediate per-task lists of the last run (all pointers optional):
 *  n_valid, hyp_id[n_valid], dist_err[n_valid], angle_err[n_valid]      <- all_configs_error_one_objH cols 4,5
 *  corners[n_valid*16] (x0..x7,y0..y7)                                   <- all_box_corners_2d_one_objH (recomputed)
 *  merged_lines[n_merged*4], n_keep, keep[n_keep], norm_score[n_keep]    <- merge_break_lines / fuse_normalize_scores_v2 */
int csb_detect_debug_task(csb_context* ctx, int task_id, int32_t* n_valid, int32_t* n_merged, int32_t* n_keep, int32_t* hyp_id,
                          double* dist_err, double* angle_err, double* corners, double* merged_lines, int32_t* keep, double* norm_score,
                          int capacity);

/* ------------------------------------------------------------------------------------------------
 * BA half
 * ------------------------------------------------------------------------------------------------ */

/* Camera-object graph in SoA form.
 *  cameras : VertexSE3Expmap estimates (world->camera), 7 doubles x y z qx qy qz qw (SE3Quat::toVector)
 *  cuboids : VertexCuboid estimates (object->world), 10 doubles x y z qx qy qz qw sx sy sz (cuboid::toVector)
 *  ec      : EdgeSE3Cuboid      (vertex0 = camera, vertex1 = cuboid), measurement 10, information 9x9 row-major
 *  ep      : EdgeSE3CuboidProj  (camera, cuboid), measurement 4 (cx cy w h), information 4x4, Kalib 3x3 row-major
 *  eo      : EdgeSE3Expmap      (camera i, camera j), measurement 7, information 6x6
 * Edges are accumulated in the order ec, ep, eo, each in array order (= g2o's id order for main_obj.cpp's ids). */
typedef struct csb_ba_graph {
    int32_t n_cam, n_cube;
    const int32_t* cam_fixed;  /* n_cam  : Vertex::fixed() */
    const int32_t* cube_fixed; /* n_cube */
    int32_t n_ec;
    const int32_t *ec_cam, *ec_cube;
    const double *ec_meas, *ec_info;
    int32_t n_ep;
    const int32_t *ep_cam, *ep_cube;
    const double *ep_meas, *ep_info, *ep_K;
    int32_t n_eo;
    const int32_t *eo_cam_i, *eo_cam_j;
    const double *eo_meas, *eo_info;
} csb_ba_graph;

/* Outputs of one linearisation; any pointer may be NULL.  Jacobians and Hessian blocks are column-major
 * like g2o's Eigen blocks: J is D x dim(vertex); H_cam[i] 6x6, H_cube[j] 9x9, *_Hij = A^T Omega B with
 * rows = vertex0 dim, cols = vertex1 dim.  b_* = -J^T Omega e accumulated per vertex. */
typedef struct csb_ba_output {
    double *ec_err, *ec_Ji, *ec_Jj; /* n_ec x 9, x 54, x 81 */
    double *ep_err, *ep_Ji, *ep_Jj; /* n_ep x 4, x 24, x 36 */
    double *eo_err, *eo_Ji, *eo_Jj; /* n_eo x 6, x 36, x 36 */
    double *H_cam, *b_cam;          /* n_cam x 36, x 6 */
    double *H_cube, *b_cube;        /* n_cube x 81, x 9 */
    double *ec_Hij, *ep_Hij, *eo_Hij; /* n_ec x 54, n_ep x 54, n_eo x 36 */
    double* chi2;                   /* 1 : sum of e^T Omega e (SparseOptimizer::activeRobustChi2) */
} csb_ba_output;

/* buildStructure(): upload topology, measurements and information matrices; build per-vertex adjacency. */
int csb_ba_set_graph(csb_context* ctx, const csb_ba_graph* graph);
/* One more keyframe for an online caller (object_slam/src/main_obj.cpp:738-803 adds, per frame, a VertexSE3Expmap, the EdgeSE3Cuboid
 * measurements of that frame, an EdgeSE3Expmap to the previous frame and -- when a landmark is seen for the first time -- a VertexCuboid,
 * then calls optimize() again).  The new camera gets index n_cam, the new cuboids n_cube .. n_cube + n_new_cubes - 1; ec edges start at the
 * new camera, eo edges end at it.  The result is the same device state csb_ba_set_graph would build from the concatenated arrays (same
 * edge order, so the same sums), but nothing is reallocated (arrays grow geometrically) and the estimates on the device -- e.g. the ones
 * csb_ba_optimize left there -- are kept; the new vertices take the estimates passed here.  Start from csb_ba_set_graph (an empty
 * graph, all counts 0, is allowed). */
typedef struct csb_ba_frame {
    const double* cam7;            /* estimate of the new camera: world->camera, x y z qx qy qz qw */
    int32_t cam_fixed;
    int32_t n_new_cubes;
    const double* new_cubes10;     /* n_new_cubes x 10 */
    const int32_t* new_cube_fixed; /* n_new_cubes */
    int32_t n_ec;
    const int32_t* ec_cube;        /* n_ec: cuboid index (may be one of the new ones) */
    const double *ec_meas, *ec_info; /* n_ec x 10, n_ec x 81 */
    int32_t n_eo;
    const int32_t* eo_cam_i;       /* n_eo: the other (earlier) camera; the edge is (cam_i -> new camera) */
    const double *eo_meas, *eo_info; /* n_eo x 7, n_eo x 36 */
} csb_ba_frame;
int csb_ba_add_frame(csb_context* ctx, const csb_ba_frame* frame, int32_t* cam_index_out);
/* computeActiveErrors() + buildSystem() with HOST vertex estimates in, HOST blocks out. */
int csb_ba_linearize(csb_context* ctx, const double* cams7, const double* cubes10, const csb_ba_output* out);
/* Device-resident: upload estimates once, linearise repeatedly (async), download when needed. */
int csb_ba_upload_estimates(csb_context* ctx, const double* cams7, const double* cubes10);
int csb_ba_run(csb_context* ctx);
int csb_ba_download(csb_context* ctx, const csb_ba_output* out);

/* How the Jacobians of EdgeSE3Cuboid and EdgeSE3Expmap are obtained.  NUMERIC (default) is the reference's own
 * definition: BaseBinaryEdge::linearizeOplus, central differences with delta = 1e-9 (base_binary_edge.hpp:130-205; upstream's exact
 * linearizeOplus is commented out, types_six_dof_expmap.h:141).  ANALYTIC evaluates the closed form (derivative of the SE(3) logarithm,
 * the yaw variant chosen by cuboid::min_log_error held fixed); it agrees with NUMERIC to the latter's round-off (~1e-7 relative).
 * EdgeSE3CuboidProj always uses central differences.  Applies to csb_ba_linearize / csb_ba_run / csb_ba_optimize. */
#define CSB_BA_JACOBIAN_NUMERIC 0
#define CSB_BA_JACOBIAN_ANALYTIC 1
int csb_ba_set_jacobian_mode(csb_context* ctx, int mode);

/* SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg on the device (replaces
 * Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189 + block_solver.hpp:353-486 + linear_solver_dense.h:65-113
 * as configured at object_slam/src/main_obj.cpp:512-517, 803).  Works on the estimates uploaded with csb_ba_upload_estimates();
 * every trial (cuboid elimination, reduced camera system, Cholesky, back-substitution, oplus, chi2) runs on the device, the host
 * only takes the scalar accept/reject decision.  The optimised estimates stay on the device and are copied to cams7_out /
 * cubes10_out (either may be NULL). */
typedef struct csb_ba_optimize_stats {
    int32_t iterations;        /* outer LM iterations run */
    int32_t trials;            /* linear solves (lambda trials) over all iterations */
    int32_t n_kernel_launches; /* kernels executed */
    int32_t schur_dim;         /* size of the reduced camera system (6 x free cameras) */
    double chi2;               /* activeRobustChi2 after the last accepted step */
    double lambda;
    float gpu_ms;              /* device time of the whole call */
    int32_t n_launches;        /* host launch calls: a linearisation and an LM trial are one CUDA-graph launch each */
} csb_ba_optimize_stats;
int csb_ba_optimize(csb_context* ctx, int iterations, double* cams7_out, double* cubes10_out, csb_ba_optimize_stats* stats);

/* ------------------------------------------------------------------------------------------------
 * Line detection: the LSD branch of line_lbd_detect::detect_filter_lines
 * ------------------------------------------------------------------------------------------------
 * Replaces, for a batch of equally sized 8-bit gray frames, what
 *   line_lbd_detect::detect_filter_lines(const cv::Mat& gray_img, cv::Mat& linesmat_out)   line_lbd/class/line_lbd_allclass.cpp:221-235
 * does with use_LSD = true (line_lbd/launch/line_detect.launch:9): LSDDetector::detectImpl (line_lbd/libs/LSDDetector.cpp:154-293, one
 * octave) -> LineSegmentDetectorImpl::detect with LSD_REFINE_ADV and the default parameters (line_lbd/libs/lsd.cpp:414-1167:
 * Gaussian blur + 0.8 down-scaling, level-line angles, region growing in raster seed order, rectangle fit, refinement, NFA search)
 * -> border clamp / border filter (LSDDetector.cpp:80-101, 206-232) -> length filter (line_lbd_allclass.cpp:200-208).
 * Output rows are linesmat_out's: float32 [x1 y1 x2 y2] in image coordinates, in the reference's order (keylines_to_mat, :28-38). */
typedef struct csb_lsd_params {
    float line_length_thres; /* line_lbd_detect::line_length_thres; both callers set 15 (main_obj.cpp:505, detect_lines.cpp:66) */
    int32_t filter;          /* 1: detect_filter_lines output; 0: every segment LineSegmentDetectorImpl::detect returns */
    int32_t max_lines;       /* capacity (rows) per frame of lines_out */
    int32_t unit_link_deg;   /* 0 = default (45).  Tuning/test knob of the work partition only: the result does not depend on it */
} csb_lsd_params;

typedef struct csb_lsd_stats {
    int64_t n_lines;        /* segments written over the whole batch */
    int64_t n_regions;      /* seeds that started a region (region_grow calls from flsd; units that cannot yield a segment are skipped) */
    int64_t n_region_px;    /* pixels accepted by those calls and by the re-growing of refine(), redone units included */
    int64_t h2d_bytes, d2h_bytes;
    int32_t scaled_width, scaled_height;
    int32_t n_kernel_launches, reserved;
    float gpu_ms_maps, gpu_ms_grow; /* CUDA-event times of the last timed run: streaming kernels / region kernel */
    int64_t grow_cycles[5];         /* SM cycles summed over warps: region_grow, region2rect, refine, rect_improve, whole seed loop */
    int64_t n_merge_rounds;         /* frames x rounds in which interacting units had to be merged and redone (0 = first partition valid) */
    int64_t n_unit_conflicts;       /* units that found a pixel of another unit aligned */
} csb_lsd_stats;

/* Host buffers in and out: gray = n_frames x height x width bytes; lines_out = n_frames x max_lines x 4 floats; n_lines_out[n_frames].
 * Returns CSB_ERR_CAPACITY (after filling what fits) if any frame produced more than max_lines segments. */
int csb_lsd_detect_batch(csb_context* ctx, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params,
                         float* lines_out, int32_t* n_lines_out, csb_lsd_stats* stats);
/* Device-resident variant: upload once, run any number of times (asynchronous on the context stream), download. */
int csb_lsd_upload(csb_context* ctx, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params);
int csb_lsd_run(csb_context* ctx, int timed);
int csb_lsd_download(csb_context* ctx, float* lines_out, int32_t* n_lines_out, csb_lsd_stats* stats);
/* Parity/debug: per-pixel maps of one frame after a run (scaled_width x scaled_height doubles each, any pointer may be NULL):
 * the blurred + down-scaled image, the gradient norm, the level-line angle in radians (-1024 = undefined). */
int csb_lsd_debug_maps(csb_context* ctx, int frame, double* scaled_out, double* modgrad_out, double* angles_out);

/* ------------------------------------------------------------------------------------------------
 * Line descriptors (BASELINE config #3): the LBD half of line_lbd_detect::detect_descrip_lines
 * ------------------------------------------------------------------------------------------------
 * Replaces, for a batch of equally sized 8-bit gray frames and their line segments, what
 *   line_lbd_detect::detect_descrip_lines(gray_img, lines_mat, line_descrips)             line_lbd/class/line_lbd_allclass.cpp:239-281
 * does after the detector: lbd->compute(gray_img, keylines, descrips) = BinaryDescriptor::computeImpl (line_lbd/libs/binary_descriptor.cpp:
 * 607-794): Gaussian blur 5x5 + Sobel (:347-402), computeLBD (:1150-1512), binaryConversion (:405-417, 766-773), with the key-line fields
 * (numOfPixels, angle) filled as LSDDetector::detectImpl does (line_lbd/libs/LSDDetector.cpp:80-101, 239-245).  One octave.
 * Output rows are line_descrips': 32 bytes per line (CV_8UC1, returnFloatDescr = false) and, on request, the 72 floats behind them
 * (returnFloatDescr = true), one row per input line, frame after frame.  A frame without lines yields no rows (computeImpl returns early).
 * Not provided: line_lbd_detect::get_line_descriptors (its mat_to_keylines reads KeyLine fields before setting them) and the matcher. */
/* The LBD descriptor and EDLines both start from BinaryDescriptor's blurred frame: cv::GaussianBlur(img, Size(5, 5), 1) on CV_8U
 * (binary_descriptor.cpp:356, 815-816).  OpenCV is not part of the reference tree and the 8-bit Gaussian exists in two generations that differ
 * in the integer taps: 4 (default) = OpenCV >= 3.4.1 / 4.x, {14, 62, 104, 62, 14} / 256; 3 = OpenCV <= 3.4.0, {14, 63, 103, 63, 14} / 256 -- the
 * generation the reference's committed object_slam outputs were produced with (tests/test_reference_replay.py). */
int csb_set_blur_generation(csb_context* ctx, int generation);

typedef struct csb_lbd_stats {
    int64_t n_lines;    /* descriptors computed */
    int64_t n_samples;  /* gradient samples gathered: 63 rows x numOfPixels per line */
    int64_t h2d_bytes, d2h_bytes;
    int32_t n_kernel_launches, reserved;
    float gpu_ms_grad, gpu_ms_describe; /* CUDA-event times of the last timed run: blur + Sobel / prefix + descriptors */
} csb_lbd_stats;

/* Host buffers in and out.  lines = float32 [x1 y1 x2 y2] rows of all frames back to back; line_offsets[n_frames + 1] = first row of each
 * frame (line_offsets[0] is the base); desc_out = rows x 32 bytes; desc_float_out = rows x 72 floats or NULL. */
int csb_lbd_describe_batch(csb_context* ctx, const uint8_t* gray, int n_frames, int width, int height, const float* lines,
                           const int32_t* line_offsets, uint8_t* desc_out, float* desc_float_out, csb_lbd_stats* stats);
/* Device-resident variant. */
int csb_lbd_upload(csb_context* ctx, const uint8_t* gray, int n_frames, int width, int height, const float* lines, const int32_t* line_offsets,
                   int want_float);
int csb_lbd_run(csb_context* ctx, int timed);
/* detect_descrip_lines without a host round trip: descriptors of the segments the last csb_lsd_run of this context left on the device
 * (its frames and its segment table are read in place).  With csb_lsd_params.filter = 1 the rows are those of the KeyLine overload
 * (line_lbd_allclass.cpp:263-281: octave 0, lineLength > line_length_thres); a descriptor does not depend on the other lines. */
int csb_lbd_run_on_lsd(csb_context* ctx, int want_float, int timed);
/* Rows of all frames back to back (n_lines_out[f] rows for frame f; any pointer may be NULL).  keylines_out = rows x 4 floats
 * {angle, numOfPixels, lineLength, 0}.  CSB_ERR_CAPACITY if the batch holds more than capacity_rows lines (nothing is copied). */
int csb_lbd_download(csb_context* ctx, uint8_t* desc_out, float* desc_float_out, float* keylines_out, int32_t* n_lines_out,
                     int64_t capacity_rows, csb_lbd_stats* stats);
/* Parity/debug: the int16 Sobel images of one frame after a run (width x height each). */
int csb_lbd_debug_gradients(csb_context* ctx, int frame, int16_t* dx_out, int16_t* dy_out);

/* ------------------------------------------------------------------------------------------------
 * Line detection, EDLines branch: line_lbd_detect::detect_filter_lines with use_LSD = false
 * ------------------------------------------------------------------------------------------------
 * What object_slam selects (object_slam/src/main_obj.cpp:503-505): BinaryDescriptor::detect (line_lbd/libs/binary_descriptor.cpp:486-590,
 * one octave) -> OctaveKeyLines (:796-1148: GaussianBlur 5x5, EDline, end-point order) -> EDLineDetector::EdgeDrawing (:1583-2380) /
 * EDline (:2383-2630) / LineValidation_ (:2793-2873) with the constructor's parameters (:1515-1525: gradient threshold 80, anchor
 * threshold 8, scan interval 2, minimum line length 15, fit error 1.6) -> filter_lines (line_lbd_allclass.cpp:200-208).
 * Same output rows as csb_lsd_*: float32 [x1 y1 x2 y2] (startPoint, endPoint), in the reference's order.  Takes csb_lsd_params
 * (line_length_thres, filter: 1 = octave 0 and lineLength > line_length_thres, 0 = every key line; max_lines; unit_link_deg unused).
 * STATUS: the device code (csrc/edlines_dev.cuh) is executed on the host against the oracle by the CPU tests; see DESIGN.md 3e. */
typedef struct csb_edlines_stats {
    int64_t n_lines;     /* segments written over the whole batch */
    int64_t n_anchors, n_chain_px, n_chains;
    int64_t h2d_bytes, d2h_bytes;
    int32_t n_kernel_launches; /* kernels executed */
    int32_t n_frames_failed; /* frames that hit EdgeDrawing's capacity errors (the reference prints "Line Detection not finished": no lines) */
    float gpu_ms_maps, gpu_ms_draw, gpu_ms_fit, reserved;
} csb_edlines_stats;

int csb_edlines_detect_batch(csb_context* ctx, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params,
                             float* lines_out, int32_t* n_lines_out, csb_edlines_stats* stats);
int csb_edlines_upload(csb_context* ctx, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params);
int csb_edlines_run(csb_context* ctx, int timed);
int csb_edlines_download(csb_context* ctx, float* lines_out, int32_t* n_lines_out, csb_edlines_stats* stats);
/* detect_descrip_lines with use_LSD = false (line_lbd_allclass.cpp:239-281): LBD descriptors of the key lines of the last csb_edlines_run,
 * computed on the device from the detector's own fields (direction = lineDirection_, numOfPixels = pixels of the fitted chain segment,
 * end points as projected; binary_descriptor.cpp:1045-1140, 1150-1512).  Rows of all frames back to back (n_lines_out[f] rows for frame f):
 * 32 bytes per line, 72 floats on request. */
int csb_edlines_describe(csb_context* ctx, int want_float);
int csb_edlines_download_descriptors(csb_context* ctx, uint8_t* desc_out, float* desc_float_out, int32_t* n_lines_out, int64_t capacity_rows);

#ifdef __cplusplus
}
#endif
#endif /* CUBESLAM_B200_H */
