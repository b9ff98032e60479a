// capi_detect.cu -- extern "C" entry points of the proposal half + context management.
//
// Replaces, for a whole batch of frames, what detect_3d_cuboid::detect_cuboid() does per frame
// (detect_3d_cuboid/include/detect_3d_cuboid/detect_3d_cuboid.h:81-92).  No CPU fallback: if CUDA is not usable every
// computing entry point returns CSB_ERR_CUDA.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "context.h"
#include "host_plan.h"

using namespace csb;

extern "C" {

const char* csb_version(void) { return "cubeslam_b200 0.1 (sm_100a)"; }

int csb_create(csb_context** out, int device_ordinal) {
    if (!out) return CSB_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device_ordinal < 0 || device_ordinal >= n) return CSB_ERR_CUDA;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return CSB_ERR_CUDA;
    csb_context* c = new csb_context();
    c->device = device_ordinal;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { delete c; return CSB_ERR_CUDA; }
    c->num_sms = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return CSB_ERR_CUDA; }
    c->own_stream = true;
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return CSB_ERR_CUDA; }
    if (cudaHostAlloc((void**)&c->h_epoch, 64, cudaHostAllocDefault) != cudaSuccess) { delete c; return CSB_ERR_CUDA; }
    *c->h_epoch = 0;
    for (int i = 0; i < 7; i++) cudaEventCreate(&c->det.ev[i]);
    *out = c;
    return CSB_OK;
}

void csb_destroy(csb_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DetectState& d = c->det;
    DevBuf* bufs[] = {&d.d_ftab, &d.d_ttab, &d.d_order, &d.d_box_begin, &d.d_lines, &d.d_maps, &d.d_ml_seg, &d.d_ml_ang, &d.d_ml_mid, &d.d_n_merged,
                      &d.d_p_dist, &d.d_p_angle, &d.d_p_hyp, &d.d_n_valid, &d.d_keep, &d.d_norm, &d.d_n_keep, &d.d_cand_score, &d.d_cand_ok,
                      &d.d_sel_idx, &d.d_sel_flag, &d.d_sel_heap, &d.d_rank_idx, &d.d_cuboids, &d.d_n_cuboids, &d.d_counters, &d.d_dbg, &d.d_gray, &d.d_cmap, &d.d_queue, &d.d_dtmp, &d.d_flags};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 7; i++)
        if (d.ev[i]) cudaEventDestroy(d.ev[i]);
    ba_release(c->ba);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->h_epoch) cudaFreeHost(c->h_epoch);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* csb_last_error(const csb_context* c) { return c ? c->err.c_str() : "null context"; }

int csb_set_stream(csb_context* c, void* s) {
    if (!c) return CSB_ERR_INVALID;
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (s) { c->stream = (cudaStream_t)s; c->own_stream = false; }
    else { CSB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    return CSB_OK;
}

int csb_synchronize(csb_context* c) {
    if (!c) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    return CSB_OK;
}

int csb_detect_plan(const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const csb_detect_params* params, csb_task* tasks_out,
                    int max_tasks, int* n_tasks_out, int64_t* n_map_floats_out) {
    if (!frames || n_frames < 0 || (!boxes && n_boxes > 0) || !params) return CSB_ERR_INVALID;
    std::vector<csb_task> tasks;
    int64_t nm = 0;
    int rc = plan_tasks(frames, n_frames, boxes, n_boxes, *params, tasks, nullptr, &nm);
    if (rc != CSB_OK) return rc;
    if (n_tasks_out) *n_tasks_out = (int)tasks.size();
    if (n_map_floats_out) *n_map_floats_out = nm;
    if (tasks_out) {
        if ((int)tasks.size() > max_tasks) return CSB_ERR_CAPACITY;
        std::copy(tasks.begin(), tasks.end(), tasks_out);
    }
    return CSB_OK;
}

// Shared by csb_detect_upload (caller-computed distance maps) and csb_detect_upload_gray (gray frames; Canny + distance
// transform run on the device at the start of every csb_detect_run).
static int detect_upload_impl(csb_context* c, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines, int n_lines,
                              const csb_task* tasks, int n_tasks, const float* dist_maps, int64_t n_map_floats, const uint8_t* gray, int64_t n_gray_bytes,
                              const csb_detect_params* params, bool stream_maps) {
    if (!c) return CSB_ERR_INVALID;
    if (!frames || n_frames <= 0 || !params || (!boxes && n_boxes > 0) || (!lines && n_lines > 0) || (!dist_maps && !gray && n_map_floats > 0)) {
        c->err = "csb_detect_upload: null argument";
        return CSB_ERR_INVALID;
    }
    if (params->max_cuboid_num < 1) { c->err = "max_cuboid_num must be >= 1"; return CSB_ERR_INVALID; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    DetectState& d = c->det;
    d.uploaded = false; d.ran = false;
    d.params = *params;
    d.tasks.clear(); d.ttab.clear();
    int64_t nm = 0;
    int rc = plan_tasks(frames, n_frames, boxes, n_boxes, *params, d.tasks, &d.ttab, &nm);
    if (rc != CSB_OK) { c->err = "csb_detect_upload: planning failed (bad frame/box ranges or sweep tables over capacity)"; return rc; }
    if ((int)d.tasks.size() != n_tasks || nm != n_map_floats) { c->err = "csb_detect_upload: tasks / n_map_floats do not match csb_detect_plan() for these inputs"; return CSB_ERR_INVALID; }
    if (tasks)
        for (int i = 0; i < n_tasks; i++)
            if (tasks[i].map_offset != d.tasks[i].map_offset || tasks[i].box_id != d.tasks[i].box_id) { c->err = "csb_detect_upload: task list differs from plan"; return CSB_ERR_INVALID; }
    d.ftab.resize(n_frames);
    d.max_groups = 0; d.max_lines_per_frame = 0;
    for (int f = 0; f < n_frames; f++) {
        if (frames[f].line_begin < 0 || frames[f].line_end > n_lines || frames[f].line_begin > frames[f].line_end) { c->err = "bad line range"; return CSB_ERR_INVALID; }
        rc = build_frame_tab(frames[f], *params, d.ftab[f]);
        if (rc != CSB_OK) { c->err = "sweep tables over capacity"; return rc; }
        d.max_groups = std::max(d.max_groups, d.ftab[f].n_roll * d.ftab[f].n_pitch * d.ftab[f].n_yaw);
        d.max_lines_per_frame = std::max(d.max_lines_per_frame, frames[f].line_end - frames[f].line_begin);
    }
    // gray frames are packed back to back (img_height x img_width bytes each)
    int64_t gray_total = 0;
    for (int f = 0; f < n_frames; f++) { d.ftab[f].gray_offset = gray_total; gray_total += (int64_t)frames[f].img_width * frames[f].img_height; }
    d.gray_mode = gray != nullptr;
    if (d.gray_mode && n_gray_bytes != gray_total) { c->err = "csb_detect_upload_gray: n_gray_bytes != sum of img_width*img_height"; return CSB_ERR_INVALID; }
    d.n_frames = n_frames; d.n_boxes = n_boxes; d.n_lines = n_lines; d.n_tasks = n_tasks; d.n_map_floats = nm;
    d.out_total = 0; d.line_cap_total = 0; d.max_hyp_per_task = 1; d.max_roi_w = 1;
    for (const TaskTab& t : d.ttab) {
        d.out_total = std::max<int64_t>(d.out_total, t.out_offset + t.n_hyp);
        d.line_cap_total = std::max<int64_t>(d.line_cap_total, (int64_t)t.line_cap_offset + (d.ftab[t.frame_id].line_end - d.ftab[t.frame_id].line_begin));
        d.max_hyp_per_task = std::max(d.max_hyp_per_task, t.n_hyp);
        d.max_roi_w = std::max(d.max_roi_w, t.roi_w);
    }
    // task queue: biggest first; box -> task range
    std::vector<int> order(n_tasks);
    std::iota(order.begin(), order.end(), 0);
    // chunked streaming (csb_detect_batch): tasks of an earlier chunk first (their maps arrive first), largest first inside a chunk
    // (chunks are equal slices of the packed map buffer; maps are laid out in task order, so a chunk is a contiguous range)
    const int64_t nm_total = std::max<int64_t>(n_map_floats, 1);
    const int n_chunks = (stream_maps && !gray && n_map_floats >= (1 << 18)) ? 8 : 1;
    auto chunk_of = [&](int task) { return (int)std::min<int64_t>(n_chunks - 1, d.ttab[task].map_offset * n_chunks / nm_total); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        int ca = chunk_of(a), cb = chunk_of(b);
        if (ca != cb) return ca < cb;
        return d.ttab[a].n_hyp > d.ttab[b].n_hyp;
    });
    std::vector<int> box_begin(n_boxes + 1, 0);
    {
        std::vector<int> cnt(n_boxes, 0);
        for (const TaskTab& t : d.ttab) cnt[t.box_id]++;
        for (int b = 0; b < n_boxes; b++) box_begin[b + 1] = box_begin[b] + cnt[b];
    }
    const size_t OT = (size_t)std::max<int64_t>(d.out_total, 1), LT = (size_t)std::max<int64_t>(d.line_cap_total, 1), NT = (size_t)std::max(n_tasks, 1);
    const int kmax = params->max_cuboid_num;
    CSB_CUDA(c, d.d_ftab.ensure(sizeof(FrameTab) * n_frames));
    CSB_CUDA(c, d.d_ttab.ensure(sizeof(TaskTab) * NT));
    CSB_CUDA(c, d.d_order.ensure(4 * NT));
    CSB_CUDA(c, d.d_box_begin.ensure(4 * (size_t)(n_boxes + 1)));
    CSB_CUDA(c, d.d_lines.ensure(32 * (size_t)std::max(n_lines, 1)));
    CSB_CUDA(c, d.d_maps.ensure(4 * (size_t)nm + 64));
    CSB_CUDA(c, d.d_ml_seg.ensure(32 * LT));
    CSB_CUDA(c, d.d_ml_ang.ensure(8 * LT));
    CSB_CUDA(c, d.d_ml_mid.ensure(16 * LT));
    CSB_CUDA(c, d.d_n_merged.ensure(4 * NT));
    CSB_CUDA(c, d.d_p_dist.ensure(8 * OT));
    CSB_CUDA(c, d.d_p_angle.ensure(8 * OT));
    CSB_CUDA(c, d.d_p_hyp.ensure(4 * OT));
    CSB_CUDA(c, d.d_n_valid.ensure(4 * NT));
    CSB_CUDA(c, d.d_keep.ensure(4 * OT));
    CSB_CUDA(c, d.d_norm.ensure(8 * OT));
    CSB_CUDA(c, d.d_n_keep.ensure(4 * NT));
    CSB_CUDA(c, d.d_cand_score.ensure(8 * OT));
    CSB_CUDA(c, d.d_cand_ok.ensure(OT));
    CSB_CUDA(c, d.d_sel_idx.ensure(8 * OT));
    CSB_CUDA(c, d.d_sel_flag.ensure(OT));
    CSB_CUDA(c, d.d_sel_heap.ensure(16 * OT));
    CSB_CUDA(c, d.d_rank_idx.ensure(8 * OT));
    CSB_CUDA(c, d.d_cuboids.ensure(sizeof(csb_cuboid) * (size_t)std::max(n_boxes, 1) * kmax));
    CSB_CUDA(c, d.d_n_cuboids.ensure(4 * (size_t)std::max(n_boxes, 1)));
    CSB_CUDA(c, d.d_counters.ensure(64));
    if (d.gray_mode) {
        CSB_CUDA(c, d.d_gray.ensure((size_t)gray_total + 64));
        CSB_CUDA(c, d.d_cmap.ensure((size_t)nm + 64));
        CSB_CUDA(c, d.d_queue.ensure(4 * (size_t)nm + 64));
        CSB_CUDA(c, d.d_dtmp.ensure(4 * (size_t)nm + 64));
    }

    cudaStream_t st = c->stream;
    CSB_CUDA(c, cudaMemcpyAsync(d.d_ftab.p, d.ftab.data(), sizeof(FrameTab) * n_frames, cudaMemcpyHostToDevice, st));
    if (n_tasks) {
        CSB_CUDA(c, cudaMemcpyAsync(d.d_ttab.p, d.ttab.data(), sizeof(TaskTab) * n_tasks, cudaMemcpyHostToDevice, st));
        CSB_CUDA(c, cudaMemcpyAsync(d.d_order.p, order.data(), 4 * (size_t)n_tasks, cudaMemcpyHostToDevice, st));
    }
    CSB_CUDA(c, cudaMemcpyAsync(d.d_box_begin.p, box_begin.data(), 4 * (size_t)(n_boxes + 1), cudaMemcpyHostToDevice, st));
    if (n_lines) CSB_CUDA(c, cudaMemcpyAsync(d.d_lines.p, lines, 32 * (size_t)n_lines, cudaMemcpyHostToDevice, st));
    bool streaming = false;
    if (d.gray_mode) { if (gray_total) CSB_CUDA(c, cudaMemcpyAsync(d.d_gray.p, gray, (size_t)gray_total, cudaMemcpyHostToDevice, st)); }
    else if (nm && n_chunks > 1 && n_tasks > 0) {
        // distance maps go out chunk by chunk on the copy stream; k_score starts right away and waits per chunk on a flag word
        // that is copied after the chunk's data (same stream => ordered)
        streaming = true;
        CSB_CUDA(c, d.d_flags.ensure(4 * 16));
        d.epoch++;
        *c->h_epoch = d.epoch;
        std::vector<int64_t> chunk_begin(n_chunks + 1, nm);
        for (int t = 0; t < n_tasks; t++) { int k = chunk_of(t); chunk_begin[k] = std::min<int64_t>(chunk_begin[k], d.ttab[t].map_offset); }
        for (int k = n_chunks - 1; k >= 0; k--) chunk_begin[k] = std::min(chunk_begin[k], chunk_begin[k + 1]);
        chunk_begin[0] = 0;
        for (int k = 0; k < n_chunks; k++) {
            const int64_t b0 = chunk_begin[k], b1 = chunk_begin[k + 1];
            if (b1 > b0) CSB_CUDA(c, cudaMemcpyAsync(d.d_maps.as<float>() + b0, dist_maps + b0, 4 * (size_t)(b1 - b0), cudaMemcpyHostToDevice, c->copy_stream));
            CSB_CUDA(c, cudaMemcpyAsync(d.d_flags.as<unsigned>() + k, c->h_epoch, 4, cudaMemcpyHostToDevice, c->copy_stream));
        }
    } else if (nm) CSB_CUDA(c, cudaMemcpyAsync(d.d_maps.p, dist_maps, 4 * (size_t)nm, cudaMemcpyHostToDevice, st));
    // the small host vectors above are pageable and go out of scope: make sure they are consumed
    CSB_CUDA(c, cudaStreamSynchronize(st));
    d.h2d_bytes = (int64_t)(sizeof(FrameTab) * n_frames + (sizeof(TaskTab) + 4) * (size_t)n_tasks + 4 * (size_t)(n_boxes + 1) + 32 * (size_t)n_lines +
                            (d.gray_mode ? (size_t)gray_total : 4 * (size_t)nm));

    DetectBuffers& B = d.B;
    B.ftab = d.d_ftab.as<FrameTab>(); B.ttab = d.d_ttab.as<TaskTab>(); B.task_order = d.d_order.as<int>(); B.box_task_begin = d.d_box_begin.as<int>();
    B.lines = d.d_lines.as<double>(); B.maps = d.d_maps.as<float>(); B.n_tasks = n_tasks; B.pad = 0;
    B.ml_seg = d.d_ml_seg.as<double>(); B.ml_ang = d.d_ml_ang.as<double>(); B.ml_mid = d.d_ml_mid.as<double>(); B.n_merged = d.d_n_merged.as<int>();
    B.p_dist = d.d_p_dist.as<double>(); B.p_angle = d.d_p_angle.as<double>(); B.p_hyp = d.d_p_hyp.as<int>(); B.n_valid = d.d_n_valid.as<int>();
    B.keep = d.d_keep.as<int>(); B.norm_score = d.d_norm.as<double>(); B.n_keep = d.d_n_keep.as<int>();
    B.cand_score = d.d_cand_score.as<double>(); B.cand_ok = d.d_cand_ok.as<unsigned char>();
    B.sel_idx = d.d_sel_idx.as<int>(); B.sel_flag = d.d_sel_flag.as<unsigned char>(); B.sel_heap = d.d_sel_heap.as<double>();
    B.rank_idx = d.d_rank_idx.as<int>(); B.cuboids = d.d_cuboids.as<csb_cuboid>(); B.n_cuboids = d.d_n_cuboids.as<int>();
    B.counters = d.d_counters.as<int>();
    B.ready_flags = streaming ? d.d_flags.as<unsigned>() : nullptr;
    B.epoch = d.epoch; B.n_chunks = n_chunks; B.map_total = nm_total;
    B.dc.max_cuboid_num = kmax; B.dc.whether_sample_cam_roll_pitch = params->whether_sample_cam_roll_pitch;
    B.dc.nominal_skew_ratio = params->nominal_skew_ratio; B.dc.max_cut_skew = params->max_cut_skew;
    d.uploaded = true;
    return CSB_OK;
}

int csb_detect_upload(csb_context* c, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines, int n_lines,
                      const csb_task* tasks, int n_tasks, const float* dist_maps, int64_t n_map_floats, const csb_detect_params* params) {
    return detect_upload_impl(c, frames, n_frames, boxes, n_boxes, lines, n_lines, tasks, n_tasks, dist_maps, n_map_floats, nullptr, 0, params, false);
}

int csb_detect_upload_gray(csb_context* c, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines, int n_lines,
                           const csb_task* tasks, int n_tasks, const uint8_t* gray, int64_t n_gray_bytes, const csb_detect_params* params) {
    if (!c) return CSB_ERR_INVALID;
    if (!gray) { c->err = "csb_detect_upload_gray: null gray buffer"; return CSB_ERR_INVALID; }
    int nt = 0;
    int64_t nm = 0;
    int rc = csb_detect_plan(frames, n_frames, boxes, n_boxes, params, nullptr, 0, &nt, &nm);
    if (rc != CSB_OK) { c->err = "csb_detect_upload_gray: planning failed"; return rc; }
    return detect_upload_impl(c, frames, n_frames, boxes, n_boxes, lines, n_lines, tasks, n_tasks, nullptr, nm, gray, n_gray_bytes, params, false);
}

int csb_detect_run(csb_context* c, int timed) {
    if (!c) return CSB_ERR_INVALID;
    DetectState& d = c->det;
    if (!d.uploaded) { c->err = "csb_detect_run before csb_detect_upload"; return CSB_ERR_STATE; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    d.launches_last = 0;
    d.timed_last = timed != 0;
    if (d.n_tasks > 0) {
        CSB_CUDA(c, cudaMemsetAsync(d.d_counters.p, 0, 64, st));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[6], st));
        if (d.gray_mode) {
            CSB_CUDA(c, launch_distmaps(d.B, d.d_gray.as<uint8_t>(), d.d_cmap.as<uint8_t>(), d.d_queue.as<int>(), d.d_dtmp.as<unsigned>(), d.d_maps.as<float>(), d.max_roi_w, st));
        }
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[0], st));
        CSB_CUDA(c, launch_prep_lines(d.B, d.max_lines_per_frame, st));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[1], st));
        CSB_CUDA(c, launch_score(d.B, d.max_groups, d.max_hyp_per_task, c->num_sms, c->max_smem_optin, &d.map_cap_floats, st));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[2], st));
        int nsel = 0;
        CSB_CUDA(c, launch_select(d.B, d.max_hyp_per_task, c->max_smem_optin, st, &nsel));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[3], st));
        CSB_CUDA(c, launch_recover(d.B, st));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[4], st));
        d.launches_last = 3 + nsel + (d.gray_mode ? 3 : 0);
    } else if (timed) {
        CSB_CUDA(c, cudaEventRecord(d.ev[6], st));
        for (int i = 0; i < 5; i++) CSB_CUDA(c, cudaEventRecord(d.ev[i], st));
    }
    if (d.n_boxes > 0) {
        CSB_CUDA(c, launch_rank(d.B, d.n_boxes, st));
        d.launches_last++;
    }
    if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[5], st));
    d.ran = true;
    return CSB_OK;
}

int csb_detect_download(csb_context* c, csb_cuboid* cuboids_out, int32_t* n_cuboids_out, csb_detect_stats* stats) {
    if (!c) return CSB_ERR_INVALID;
    DetectState& d = c->det;
    if (!d.ran) { c->err = "csb_detect_download before csb_detect_run"; return CSB_ERR_STATE; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int kmax = d.params.max_cuboid_num;
    int64_t d2h = 0;
    if (d.n_boxes > 0) {
        if (cuboids_out) { CSB_CUDA(c, cudaMemcpyAsync(cuboids_out, d.d_cuboids.p, sizeof(csb_cuboid) * (size_t)d.n_boxes * kmax, cudaMemcpyDeviceToHost, st)); d2h += sizeof(csb_cuboid) * (size_t)d.n_boxes * kmax; }
        if (n_cuboids_out) { CSB_CUDA(c, cudaMemcpyAsync(n_cuboids_out, d.d_n_cuboids.p, 4 * (size_t)d.n_boxes, cudaMemcpyDeviceToHost, st)); d2h += 4 * (size_t)d.n_boxes; }
    }
    std::vector<int> nv, nk;
    if (stats && d.n_tasks > 0) {
        nv.resize(d.n_tasks); nk.resize(d.n_tasks);
        CSB_CUDA(c, cudaMemcpyAsync(nv.data(), d.d_n_valid.p, 4 * (size_t)d.n_tasks, cudaMemcpyDeviceToHost, st));
        CSB_CUDA(c, cudaMemcpyAsync(nk.data(), d.d_n_keep.p, 4 * (size_t)d.n_tasks, cudaMemcpyDeviceToHost, st));
        d2h += 8 * (size_t)d.n_tasks;
    }
    CSB_CUDA(c, cudaStreamSynchronize(st));
    CSB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    d.d2h_bytes = d2h;
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (int t = 0; t < d.n_tasks; t++) { stats->n_enumerated += d.ttab[t].n_enum; stats->n_scored += nv[t]; stats->n_kept += nk[t]; }
        stats->h2d_bytes = d.h2d_bytes; stats->d2h_bytes = d2h;
        stats->n_kernel_launches = d.launches_last;
        for (const TaskTab& t : d.ttab) stats->n_tasks_smem_map += (t.roi_w * t.roi_h <= d.map_cap_floats) ? 1 : 0;
        if (d.timed_last) {
            cudaEventElapsedTime(&stats->gpu_ms_prep, d.ev[0], d.ev[1]);
            cudaEventElapsedTime(&stats->gpu_ms_score, d.ev[1], d.ev[2]);
            cudaEventElapsedTime(&stats->gpu_ms_select, d.ev[2], d.ev[3]);
            cudaEventElapsedTime(&stats->gpu_ms_recover, d.ev[3], d.ev[4]);
            cudaEventElapsedTime(&stats->gpu_ms_rank, d.ev[4], d.ev[5]);
            cudaEventElapsedTime(&stats->gpu_ms_distmap, d.ev[6], d.ev[0]);
        }
    }
    return CSB_OK;
}

int csb_detect_batch(csb_context* c, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines, int n_lines,
                     const csb_task* tasks, int n_tasks, const float* dist_maps, int64_t n_map_floats, const csb_detect_params* params,
                     csb_cuboid* cuboids_out, int32_t* n_cuboids_out, csb_detect_stats* stats) {
    int rc = detect_upload_impl(c, frames, n_frames, boxes, n_boxes, lines, n_lines, tasks, n_tasks, dist_maps, n_map_floats, nullptr, 0, params, true);
    if (rc != CSB_OK) return rc;
    rc = csb_detect_run(c, stats != nullptr);
    if (rc != CSB_OK) return rc;
    return csb_detect_download(c, cuboids_out, n_cuboids_out, stats);
}

int csb_detect_batch_gray(csb_context* c, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines, int n_lines,
                          const csb_task* tasks, int n_tasks, const uint8_t* gray, int64_t n_gray_bytes, const csb_detect_params* params,
                          csb_cuboid* cuboids_out, int32_t* n_cuboids_out, csb_detect_stats* stats) {
    int rc = csb_detect_upload_gray(c, frames, n_frames, boxes, n_boxes, lines, n_lines, tasks, n_tasks, gray, n_gray_bytes, params);
    if (rc != CSB_OK) return rc;
    rc = csb_detect_run(c, stats != nullptr);
    if (rc != CSB_OK) return rc;
    return csb_detect_download(c, cuboids_out, n_cuboids_out, stats);
}

int csb_detect_debug_map(csb_context* c, int task_id, float* dist_map_out, uint8_t* edges_out, int capacity) {
    if (!c) return CSB_ERR_INVALID;
    DetectState& d = c->det;
    if (!d.ran) { c->err = "debug before run"; return CSB_ERR_STATE; }
    if (task_id < 0 || task_id >= d.n_tasks) return CSB_ERR_INVALID;
    const TaskTab& t = d.ttab[task_id];
    const int n = t.roi_w * t.roi_h;
    if (n > capacity) return CSB_ERR_CAPACITY;
    CSB_CUDA(c, cudaSetDevice(c->device));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (dist_map_out) CSB_CUDA(c, cudaMemcpy(dist_map_out, d.d_maps.as<float>() + t.map_offset, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    if (edges_out) {
        if (!d.gray_mode) { c->err = "edge maps exist only after csb_detect_upload_gray"; return CSB_ERR_STATE; }
        CSB_CUDA(c, cudaMemcpy(edges_out, d.d_cmap.as<uint8_t>() + t.map_offset, (size_t)n, cudaMemcpyDeviceToHost));
    }
    return CSB_OK;
}

int csb_detect_debug_task(csb_context* c, int task_id, int32_t* n_valid, int32_t* n_merged, int32_t* n_keep, int32_t* hyp_id, double* dist_err,
                          double* angle_err, double* corners, double* merged_lines, int32_t* keep, double* norm_score, int capacity) {
    if (!c) return CSB_ERR_INVALID;
    DetectState& d = c->det;
    if (!d.ran) { c->err = "debug before run"; return CSB_ERR_STATE; }
    if (task_id < 0 || task_id >= d.n_tasks) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    CSB_CUDA(c, cudaStreamSynchronize(st));
    const TaskTab& t = d.ttab[task_id];
    int nv = 0, nm = 0, nk = 0;
    CSB_CUDA(c, cudaMemcpy(&nv, d.d_n_valid.as<int>() + task_id, 4, cudaMemcpyDeviceToHost));
    CSB_CUDA(c, cudaMemcpy(&nm, d.d_n_merged.as<int>() + task_id, 4, cudaMemcpyDeviceToHost));
    CSB_CUDA(c, cudaMemcpy(&nk, d.d_n_keep.as<int>() + task_id, 4, cudaMemcpyDeviceToHost));
    if (n_valid) *n_valid = nv;
    if (n_merged) *n_merged = nm;
    if (n_keep) *n_keep = nk;
    if ((hyp_id || dist_err || angle_err || corners) && nv > capacity) return CSB_ERR_CAPACITY;
    if ((keep || norm_score) && nk > capacity) return CSB_ERR_CAPACITY;
    if (merged_lines && nm > capacity) return CSB_ERR_CAPACITY;
    const size_t ob = (size_t)t.out_offset;
    if (hyp_id && nv) CSB_CUDA(c, cudaMemcpy(hyp_id, d.d_p_hyp.as<int>() + ob, 4 * (size_t)nv, cudaMemcpyDeviceToHost));
    if (dist_err && nv) CSB_CUDA(c, cudaMemcpy(dist_err, d.d_p_dist.as<double>() + ob, 8 * (size_t)nv, cudaMemcpyDeviceToHost));
    if (angle_err && nv) CSB_CUDA(c, cudaMemcpy(angle_err, d.d_p_angle.as<double>() + ob, 8 * (size_t)nv, cudaMemcpyDeviceToHost));
    if (keep && nk) CSB_CUDA(c, cudaMemcpy(keep, d.d_keep.as<int>() + ob, 4 * (size_t)nk, cudaMemcpyDeviceToHost));
    if (norm_score && nk) CSB_CUDA(c, cudaMemcpy(norm_score, d.d_norm.as<double>() + ob, 8 * (size_t)nk, cudaMemcpyDeviceToHost));
    if (merged_lines && nm) CSB_CUDA(c, cudaMemcpy(merged_lines, d.d_ml_seg.as<double>() + 4 * (size_t)t.line_cap_offset, 32 * (size_t)nm, cudaMemcpyDeviceToHost));
    if (corners && nv) {
        CSB_CUDA(c, d.d_dbg.ensure(128 * (size_t)nv));
        CSB_CUDA(c, launch_debug_corners(d.B, task_id, nv, d.d_dbg.as<double>(), st));
        CSB_CUDA(c, cudaMemcpyAsync(corners, d.d_dbg.p, 128 * (size_t)nv, cudaMemcpyDeviceToHost, st));
        CSB_CUDA(c, cudaStreamSynchronize(st));
    }
    return CSB_OK;
}

}  // extern "C"
