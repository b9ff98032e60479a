// capi_detect.cu -- extern "C" entry points of the proposal half + context management.
//
// Replaces, for a whole batch of frames, what detect_3d_cuboid::detect_cuboid() does per frame
// (detect_3d_cuboid/include/detect_3d_cuboid/detect_3d_cuboid.h:81-92).  No CPU fallback: if CUDA is not usable every
// computing entry point returns CSB_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "context.h"
#include "host_plan.h"

using namespace csb;

// process-wide epoch of the chunked distance-map upload (csb_detect_batch): see detect_upload_impl
static std::atomic<unsigned> g_stream_epoch{0};

// k_score polls flags that a second stream writes while it runs; with a single hardware queue (CUDA_DEVICE_MAX_CONNECTIONS=1) the copies
// could be queued behind the kernel that waits for them, so the streamed upload is only used when the queues are independent.
static bool copy_engine_is_concurrent() {
    const char* e = std::getenv("CUDA_DEVICE_MAX_CONNECTIONS");
    return !(e && std::atoi(e) <= 1);
}

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
    DevBuf* bufs[] = {&d.d_tables, &d.d_results, &d.d_maps, &d.d_ml_seg, &d.d_ml_ang, &d.d_ml_mid, &d.d_n_merged, &d.d_vp_sup,
                      &d.d_p_dist, &d.d_p_angle, &d.d_p_hyp, &d.d_keep, &d.d_norm, &d.d_cand_score, &d.d_cand_ok,
                      &d.d_sel_idx, &d.d_sel_flag, &d.d_sel_heap, &d.d_rank_idx, &d.d_counters, &d.d_dbg, &d.d_gray, &d.d_cmap, &d.d_queue, &d.d_dtmp, &d.d_flags, &d.d_segbits};
    for (DevBuf* b : bufs) b->release();
    d.h_tables.release(); d.h_results.release();
    if (d.ev_tables) cudaEventDestroy(d.ev_tables);
    if (d.ev_order) cudaEventDestroy(d.ev_order);
    for (int i = 0; i < 7; i++)
        if (d.ev[i]) cudaEventDestroy(d.ev[i]);
    ba_release(c->ba);
    lsd_release(c->lsd);
    lbd_release(c->lbd);
    edlines_release(c->edlines);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->det.ev_fork) cudaEventDestroy(c->det.ev_fork);
    if (c->det.ev_join) cudaEventDestroy(c->det.ev_join);
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

int csb_set_blur_generation(csb_context* c, int generation) {
    if (!c || (generation != 3 && generation != 4)) return CSB_ERR_INVALID;
    c->blur_generation = generation;
    return CSB_OK;
}

int csb_set_option(csb_context* c, int option, int value) {
    if (!c) return CSB_ERR_INVALID;
    if (option == CSB_OPT_GRAY_GATHER) { c->det.gray_gather_enabled = value != 0; return CSB_OK; }
    c->err = "csb_set_option: unknown option";
    return CSB_ERR_INVALID;
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

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Shared by csb_detect_upload (caller-computed distance maps), csb_detect_batch (the same, streamed in chunks) and
// csb_detect_upload_gray (gray frames; Canny + distance transform run on the device at the start of every csb_detect_run;
// n_map_floats < 0 = "not known to the caller").
//
// Order of work: the bulk payload (maps or gray frames) is queued on the copy engine FIRST, then the host plans tasks and builds the
// sweep tables while it is in flight; the tables go out as one copy from a pinned staging arena.
static int detect_upload_impl(csb_context* c, const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const double* lines, int n_lines,
                              const csb_task* tasks, int n_tasks, const float* dist_maps, int64_t n_map_floats, const uint8_t* gray, int64_t n_gray_bytes,
                              const csb_detect_params* params, bool stream_maps) {
    if (!c) return CSB_ERR_INVALID;
    if (!frames || n_frames <= 0 || !params || (!boxes && n_boxes > 0) || (!lines && n_lines > 0) || (!dist_maps && !gray && n_map_floats > 0) || n_tasks < 0) {
        c->err = "csb_detect_upload: null argument";
        return CSB_ERR_INVALID;
    }
    if (params->max_cuboid_num < 1) { c->err = "max_cuboid_num must be >= 1"; return CSB_ERR_INVALID; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    DetectState& d = c->det;
    cudaStream_t st = c->stream;
    d.uploaded = false; d.ran = false;
    d.params = *params;
    d.gray_mode = gray != nullptr;

    // ---- 1. bulk payload on its way
    int64_t gray_total = 0;
    for (int f = 0; f < n_frames; f++) gray_total += (int64_t)frames[f].img_width * frames[f].img_height;
    bool streaming = false;
    int n_chunks = 1;
    int64_t chunk_end[CSB_MAX_CHUNKS];
    for (int k = 0; k < CSB_MAX_CHUNKS; k++) chunk_end[k] = n_map_floats;
    if (d.gray_mode) {
        // gray frames are packed back to back (img_height x img_width bytes each)
        if (n_gray_bytes != gray_total) { c->err = "csb_detect_upload_gray: n_gray_bytes != sum of img_width*img_height"; return CSB_ERR_INVALID; }
        CSB_CUDA(c, d.d_gray.ensure_zeroed((size_t)gray_total + 64, st));  // cleared once: with the ROI-segment upload k_distmap's aligned word loads touch bytes of segments that are never fetched (they reach no result)
        // Pinned (device-mapped) caller memory: the frames are not copied as a whole -- once the task table is on the device, a kernel
        // fetches the 512-byte segments the ROIs touch straight from the caller's buffer (launch_gray_gather below).  Pageable memory
        // goes through the copy engine as one block.
        d.gray_mapped = nullptr;
        if (gray_total && d.gray_gather_enabled && ((uintptr_t)gray & 15) == 0) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, gray) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) d.gray_mapped = (const uint8_t*)at.devicePointer;
            else cudaGetLastError();
        }
        if (gray_total && !d.gray_mapped) CSB_CUDA(c, cudaMemcpyAsync(d.d_gray.p, gray, (size_t)gray_total, cudaMemcpyHostToDevice, st));
    } else if (n_map_floats > 0) {
        CSB_CUDA(c, d.d_maps.ensure(4 * (size_t)n_map_floats + 64));
        if (stream_maps && n_map_floats >= (1 << 18) && copy_engine_is_concurrent()) {
            // csb_detect_batch: equal slices on the copy stream, each followed by its flag word (same stream => ordered).  k_score starts
            // right away and a task waits for the slice that holds the END of its map (slices complete in order).
            streaming = true;
            n_chunks = CSB_MAX_CHUNKS;
            if (!d.d_flags.p) {
                // a fresh (possibly recycled) block: no flag may carry a value that a later epoch could equal
                CSB_CUDA(c, d.d_flags.ensure(4 * CSB_MAX_CHUNKS));
                CSB_CUDA(c, cudaMemset(d.d_flags.p, 0, 4 * CSB_MAX_CHUNKS));
            }
            CSB_CUDA(c, cudaStreamSynchronize(c->copy_stream));  // nothing may still be reading the epoch word
            // epochs are unique per process (not per context), never 0: a flag left behind by a destroyed context cannot match
            do { d.epoch = g_stream_epoch.fetch_add(1) + 1; } while (d.epoch == 0);
            *c->h_epoch = d.epoch;
            // the copies below overwrite d_maps: earlier work on the compute stream (a csb_detect_run that was never downloaded) may still read it
            if (!d.ev_order) CSB_CUDA(c, cudaEventCreateWithFlags(&d.ev_order, cudaEventDisableTiming));
            CSB_CUDA(c, cudaEventRecord(d.ev_order, st));
            CSB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, d.ev_order, 0));
            int64_t b0 = 0;
            for (int k = 0; k < n_chunks; k++) {
                const int64_t b1 = (k == n_chunks - 1) ? n_map_floats : ((n_map_floats * (k + 1) / n_chunks) & ~(int64_t)3);
                chunk_end[k] = b1;
                if (b1 > b0) CSB_CUDA(c, cudaMemcpyAsync(d.d_maps.as<float>() + b0, dist_maps + b0, 4 * (size_t)(b1 - b0), cudaMemcpyHostToDevice, c->copy_stream));
                CSB_CUDA(c, cudaMemcpyAsync(d.d_flags.as<unsigned>() + k, c->h_epoch, 4, cudaMemcpyHostToDevice, c->copy_stream));
                b0 = b1;
            }
        } else {
            CSB_CUDA(c, cudaMemcpyAsync(d.d_maps.p, dist_maps, 4 * (size_t)n_map_floats, cudaMemcpyHostToDevice, st));
        }
    }
    // from here on an error return must not leave the copy stream reading the caller's buffer
    auto fail = [&](int code, const char* msg) { cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(st); c->err = msg; return code; };

    // ---- 2. host planning while the payload is in flight
    d.tasks.clear(); d.ttab.clear();
    int64_t nm = 0;
    int rc = plan_tasks(frames, n_frames, boxes, n_boxes, *params, d.tasks, &d.ttab, &nm);
    if (rc != CSB_OK) return fail(rc, "csb_detect_upload: planning failed (bad frame/box ranges or sweep tables over capacity)");
    if ((int)d.tasks.size() != n_tasks || (n_map_floats >= 0 && nm != n_map_floats))
        return fail(CSB_ERR_INVALID, "csb_detect_upload: tasks / n_map_floats do not match csb_detect_plan() for these inputs");
    if (tasks)
        for (int i = 0; i < n_tasks; i++)
            if (tasks[i].map_offset != d.tasks[i].map_offset || tasks[i].box_id != d.tasks[i].box_id) return fail(CSB_ERR_INVALID, "csb_detect_upload: task list differs from plan");
    d.n_frames = n_frames; d.n_boxes = n_boxes; d.n_lines = n_lines; d.n_tasks = n_tasks; d.n_map_floats = nm;

    // table arena layout (device and pinned host mirror)
    const size_t NT = (size_t)std::max(n_tasks, 1);
    const size_t off_ftab = 0;
    const size_t off_ttab = align256(off_ftab + sizeof(FrameTab) * (size_t)n_frames);
    const size_t off_order = align256(off_ttab + sizeof(TaskTab) * NT);
    const size_t off_box = align256(off_order + 4 * NT);
    const size_t off_lines = align256(off_box + 4 * (size_t)(n_boxes + 1));
    const size_t tab_bytes = align256(off_lines + 32 * (size_t)std::max(n_lines, 1));
    if (d.ev_tables) CSB_CUDA(c, cudaEventSynchronize(d.ev_tables));  // previous upload's copy out of h_tables
    CSB_CUDA(c, d.h_tables.ensure(tab_bytes));
    CSB_CUDA(c, d.d_tables.ensure(tab_bytes));
    char* hb = d.h_tables.as<char>();
    FrameTab* h_ftab = reinterpret_cast<FrameTab*>(hb + off_ftab);
    TaskTab* h_ttab = reinterpret_cast<TaskTab*>(hb + off_ttab);
    int* h_order = reinterpret_cast<int*>(hb + off_order);
    int* h_box = reinterpret_cast<int*>(hb + off_box);

    d.ftab.resize(n_frames);
    d.max_groups = 0; d.max_lines_per_frame = 0;
    int64_t goff = 0;
    for (int f = 0; f < n_frames; f++) {
        if (frames[f].line_begin < 0 || frames[f].line_end > n_lines || frames[f].line_begin > frames[f].line_end) return fail(CSB_ERR_INVALID, "bad line range");
        FrameTab& ft = h_ftab[f];
        rc = build_frame_tab(frames[f], *params, ft);
        if (rc != CSB_OK) return fail(rc, "sweep tables over capacity");
        ft.gray_offset = goff;
        goff += (int64_t)frames[f].img_width * frames[f].img_height;
        d.max_groups = std::max(d.max_groups, ft.n_roll * ft.n_pitch * ft.n_yaw);
        d.max_lines_per_frame = std::max(d.max_lines_per_frame, frames[f].line_end - frames[f].line_begin);
    }
    d.out_total = 0; d.line_cap_total = 0; d.max_hyp_per_task = 1; d.max_roi_w = 1; d.max_roi_px = 1;
    for (const TaskTab& t : d.ttab) {
        d.out_total = std::max<int64_t>(d.out_total, t.out_offset + t.n_hyp);
        d.line_cap_total = std::max<int64_t>(d.line_cap_total, (int64_t)t.line_cap_offset + (frames[t.frame_id].line_end - frames[t.frame_id].line_begin));
        d.max_hyp_per_task = std::max(d.max_hyp_per_task, t.n_hyp);
        d.max_roi_w = std::max(d.max_roi_w, t.roi_w);
        d.max_roi_px = std::max(d.max_roi_px, (((t.roi_w + 15) >> 4) + 1) * t.roi_h);  // words of the packed edge map, + 1 per row: k_distmap overlays its edge-bit rows and carries
    }
    if (n_tasks) std::memcpy(h_ttab, d.ttab.data(), sizeof(TaskTab) * (size_t)n_tasks);
    // task queue: tasks whose slice arrives first go first; inside a slice the biggest first
    auto chunk_of = [&](int task) {
        const TaskTab& t = d.ttab[task];
        const int64_t last = t.map_offset + (int64_t)t.roi_w * t.roi_h - 1;
        int k = 0;
        while (k < n_chunks - 1 && chunk_end[k] <= last) k++;
        return k;
    };
    {
        std::vector<int> ck(n_tasks);
        for (int i = 0; i < n_tasks; i++) { h_order[i] = i; ck[i] = streaming ? chunk_of(i) : 0; }
        std::stable_sort(h_order, h_order + n_tasks, [&](int a, int b) {
            if (ck[a] != ck[b]) return ck[a] < ck[b];
            return d.ttab[a].n_hyp > d.ttab[b].n_hyp;
        });
    }
    {
        std::fill(h_box, h_box + n_boxes + 1, 0);
        for (const TaskTab& t : d.ttab) h_box[t.box_id + 1]++;
        for (int b = 0; b < n_boxes; b++) h_box[b + 1] += h_box[b];
    }
    if (n_lines) std::memcpy(hb + off_lines, lines, 32 * (size_t)n_lines);
    CSB_CUDA(c, cudaMemcpyAsync(d.d_tables.p, hb, tab_bytes, cudaMemcpyHostToDevice, st));
    if (!d.ev_tables) CSB_CUDA(c, cudaEventCreateWithFlags(&d.ev_tables, cudaEventDisableTiming));
    CSB_CUDA(c, cudaEventRecord(d.ev_tables, st));

    // ---- 3. work buffers
    const size_t OT = (size_t)std::max<int64_t>(d.out_total, 1), LT = (size_t)std::max<int64_t>(d.line_cap_total, 1);
    const int kmax = params->max_cuboid_num;
    const size_t NB = (size_t)std::max(n_boxes, 1);
    d.res_off_ncub = align256(sizeof(csb_cuboid) * NB * kmax);
    d.res_off_nvalid = align256(d.res_off_ncub + 4 * NB);
    d.res_off_nkeep = align256(d.res_off_nvalid + 4 * NT);
    d.res_off_misc = align256(d.res_off_nkeep + 4 * NT);  // [0]: 128-byte segments fetched by the gray gather
    d.res_bytes = align256(d.res_off_misc + 64);
    {
        // the result arena goes to the host as one block: its padding and the unused cuboid slots are defined bytes (zero) from the start
        const size_t cap_before = d.d_results.cap;
        CSB_CUDA(c, d.d_results.ensure(d.res_bytes));
        if (d.d_results.cap != cap_before) CSB_CUDA(c, cudaMemsetAsync(d.d_results.p, 0, d.d_results.cap, st));
    }
    CSB_CUDA(c, d.h_results.ensure(d.res_bytes));
    if (d.gray_mode || n_map_floats <= 0) CSB_CUDA(c, d.d_maps.ensure(4 * (size_t)nm + 64));
    CSB_CUDA(c, d.d_ml_seg.ensure(32 * LT));
    CSB_CUDA(c, d.d_ml_ang.ensure(8 * LT));
    CSB_CUDA(c, d.d_ml_mid.ensure(16 * LT));
    CSB_CUDA(c, d.d_n_merged.ensure(4 * NT));
    CSB_CUDA(c, d.d_vp_sup.ensure(48 * NT * (size_t)std::max(d.max_groups, 1)));
    CSB_CUDA(c, d.d_p_dist.ensure(8 * OT));
    CSB_CUDA(c, d.d_p_angle.ensure(8 * OT));
    CSB_CUDA(c, d.d_p_hyp.ensure(4 * OT));
    CSB_CUDA(c, d.d_keep.ensure(4 * OT));
    CSB_CUDA(c, d.d_norm.ensure(8 * OT));
    CSB_CUDA(c, d.d_cand_score.ensure(8 * OT));
    CSB_CUDA(c, d.d_cand_ok.ensure(OT));
    CSB_CUDA(c, d.d_sel_idx.ensure(8 * OT));
    CSB_CUDA(c, d.d_sel_flag.ensure(OT));
    CSB_CUDA(c, d.d_sel_heap.ensure(16 * OT));
    CSB_CUDA(c, d.d_rank_idx.ensure(8 * OT));
    CSB_CUDA(c, d.d_counters.ensure(64));
    if (d.gray_mode) {
        CSB_CUDA(c, d.d_cmap.ensure(4 * (size_t)nm + 64));
        CSB_CUDA(c, d.d_queue.ensure(4 * (size_t)nm + 64));
        CSB_CUDA(c, d.d_dtmp.ensure(4 * (size_t)nm + 64));
    }
    d.h2d_bytes = (int64_t)(tab_bytes + (d.gray_mode ? (d.gray_mapped ? (size_t)0 : (size_t)gray_total) : 4 * (size_t)nm) + (streaming ? 4 * (size_t)n_chunks : 0));

    DetectBuffers& B = d.B;
    char* db = d.d_tables.as<char>();
    char* dr = d.d_results.as<char>();
    B.ftab = reinterpret_cast<const FrameTab*>(db + off_ftab); B.ttab = reinterpret_cast<const TaskTab*>(db + off_ttab);
    B.task_order = reinterpret_cast<const int*>(db + off_order); B.box_task_begin = reinterpret_cast<const int*>(db + off_box);
    B.lines = reinterpret_cast<const double*>(db + off_lines); B.maps = d.d_maps.as<float>(); B.n_tasks = n_tasks; B.pad = 0;
    B.ml_seg = d.d_ml_seg.as<double>(); B.ml_ang = d.d_ml_ang.as<double>(); B.ml_mid = d.d_ml_mid.as<double>(); B.n_merged = d.d_n_merged.as<int>();
    B.vp_sup = d.d_vp_sup.as<double>(); B.sup_stride = 6 * (long long)std::max(d.max_groups, 1);
    B.p_dist = d.d_p_dist.as<double>(); B.p_angle = d.d_p_angle.as<double>(); B.p_hyp = d.d_p_hyp.as<int>();
    B.n_valid = reinterpret_cast<int*>(dr + d.res_off_nvalid);
    B.keep = d.d_keep.as<int>(); B.norm_score = d.d_norm.as<double>(); B.n_keep = reinterpret_cast<int*>(dr + d.res_off_nkeep);
    B.cand_score = d.d_cand_score.as<double>(); B.cand_ok = d.d_cand_ok.as<unsigned char>();
    B.sel_idx = d.d_sel_idx.as<int>(); B.sel_flag = d.d_sel_flag.as<unsigned char>(); B.sel_heap = d.d_sel_heap.as<double>();
    B.rank_idx = d.d_rank_idx.as<int>(); B.cuboids = reinterpret_cast<csb_cuboid*>(dr); B.n_cuboids = reinterpret_cast<int*>(dr + d.res_off_ncub);
    B.counters = d.d_counters.as<int>();
    B.ready_flags = streaming ? d.d_flags.as<unsigned>() : nullptr;
    B.epoch = d.epoch; B.n_chunks = n_chunks;
    for (int k = 0; k < CSB_MAX_CHUNKS; k++) B.chunk_end[k] = chunk_end[k];
    B.dc.max_cuboid_num = kmax; B.dc.whether_sample_cam_roll_pitch = params->whether_sample_cam_roll_pitch;
    B.dc.nominal_skew_ratio = params->nominal_skew_ratio; B.dc.max_cut_skew = params->max_cut_skew;
    for (int f = 0; f < n_frames; f++) d.ftab[f] = h_ftab[f];  // host copy for the debug entries (h_tables is reused by the next upload)
    if (d.gray_mode) {
        int* misc = reinterpret_cast<int*>(dr + d.res_off_misc);
        CSB_CUDA(c, cudaMemsetAsync(misc, 0, 64, st));
        if (d.gray_mapped) {
            const long long whole = gray_total & ~(long long)15;  // the last (partial) 16-byte chunk goes by a plain copy: nothing is read past the caller's buffer
            CSB_CUDA(c, d.d_segbits.ensure(4 * (size_t)((gray_total / CSB_GRAY_SEG + 64) / 32 + 2)));
            CSB_CUDA(c, launch_gray_gather(B, d.gray_mapped, d.d_gray.as<uint8_t>(), whole, d.d_segbits.as<unsigned>(), misc, c->num_sms, st));
            if (gray_total > whole) CSB_CUDA(c, cudaMemcpyAsync(d.d_gray.as<uint8_t>() + whole, gray + whole, (size_t)(gray_total - whole), cudaMemcpyHostToDevice, st));
        }
    }
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
    return detect_upload_impl(c, frames, n_frames, boxes, n_boxes, lines, n_lines, tasks, n_tasks, nullptr, -1, gray, n_gray_bytes, params, false);
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
            // The line kernels (k_prep_lines, k_vp_support) need the tables only, not the maps: they run on the context's second stream beside
            // k_distmap and join before k_score (what a single blocking call per frame gains: 0.14 ms of latency)
            if (!d.ev_fork) { CSB_CUDA(c, cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming)); CSB_CUDA(c, cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming)); }
            CSB_CUDA(c, cudaEventRecord(d.ev_fork, st));
            CSB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, d.ev_fork, 0));
            CSB_CUDA(c, launch_prep_lines(d.B, d.max_lines_per_frame, d.max_groups, c->copy_stream));
            CSB_CUDA(c, cudaEventRecord(d.ev_join, c->copy_stream));
            CSB_CUDA(c, launch_distmaps(d.B, d.d_gray.as<uint8_t>(), d.d_cmap.as<uint8_t>(), d.d_queue.as<int>(), d.d_dtmp.as<unsigned>(), d.d_maps.as<float>(), d.max_roi_w, d.max_roi_px, st));
            if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[0], st));
            CSB_CUDA(c, cudaStreamWaitEvent(st, d.ev_join, 0));
        } else {
            if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[0], st));
            CSB_CUDA(c, launch_prep_lines(d.B, d.max_lines_per_frame, d.max_groups, st));
        }
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[1], st));
        CSB_CUDA(c, launch_score(d.B, d.max_groups, d.max_hyp_per_task, c->num_sms, c->max_smem_optin, &d.map_cap_floats, st));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[2], st));
        int nsel = 0;
        CSB_CUDA(c, launch_select(d.B, d.max_hyp_per_task, c->max_smem_optin, st, &nsel));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[3], st));
        CSB_CUDA(c, launch_recover(d.B, st));
        if (timed) CSB_CUDA(c, cudaEventRecord(d.ev[4], st));
        d.launches_last = 4 + nsel + (d.gray_mode ? 1 : 0);  // prep_lines, vp_support, score, recover + select launches (+ k_distmap)
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
    // one copy of the result arena (cuboids | n_cuboids | n_valid | n_keep) into pinned memory, then plain memcpy to the caller
    const size_t take = stats ? d.res_bytes : d.res_off_nvalid;
    CSB_CUDA(c, cudaMemcpyAsync(d.h_results.p, d.d_results.p, take, cudaMemcpyDeviceToHost, st));
    CSB_CUDA(c, cudaStreamSynchronize(st));
    CSB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    const char* hr = d.h_results.as<char>();
    if (d.n_boxes > 0) {
        if (cuboids_out) std::memcpy(cuboids_out, hr, sizeof(csb_cuboid) * (size_t)d.n_boxes * kmax);
        if (n_cuboids_out) std::memcpy(n_cuboids_out, hr + d.res_off_ncub, 4 * (size_t)d.n_boxes);
    }
    const int* nv = reinterpret_cast<const int*>(hr + d.res_off_nvalid);
    const int* nk = reinterpret_cast<const int*>(hr + d.res_off_nkeep);
    const int64_t d2h = (int64_t)take;
    d.d2h_bytes = d2h;
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (int t = 0; t < d.n_tasks; t++) { stats->n_enumerated += d.ttab[t].n_enum; stats->n_scored += nv[t]; stats->n_kept += nk[t]; }
        stats->h2d_bytes = d.h2d_bytes + (d.gray_mode ? CSB_GRAY_SEG * (int64_t)reinterpret_cast<const int*>(hr + d.res_off_misc)[0] : 0);
        stats->d2h_bytes = d2h;
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

int csb_detect_debug_atan2(csb_context* c, const double* y, const double* x, double* out, int n, int* n_fallback) {
    if (!c || !y || !x || !out || n < 0 || n % 6 != 0) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    double* d = nullptr;
    int* df = nullptr;
    CSB_CUDA(c, cudaMalloc(&d, (size_t)3 * n * sizeof(double) + 64));
    CSB_CUDA(c, cudaMalloc(&df, sizeof(int)));
    cudaMemsetAsync(df, 0, sizeof(int), c->stream);
    cudaMemcpyAsync(d, y, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d + n, x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
    cudaError_t e = launch_debug_atan2(d, d + n, d + 2 * (size_t)n, n / 6, df, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
    int nf = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&nf, df, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    cudaFree(df);
    if (n_fallback) *n_fallback = nf;
    CSB_CUDA(c, e);
    return CSB_OK;
}

int csb_detect_debug_score_phases(csb_context* c, uint64_t* cycles8, int reset) {  // 12 entries, see the header
    if (!c || !cycles8) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    unsigned long long v[12];
    CSB_CUDA(c, score_phase_cycles(v, reset != 0));
    for (int i = 0; i < 12; i++) cycles8[i] = v[i];
    return CSB_OK;
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
        // device layout: 2 bits per pixel, 16 pixels per word, row pitch ceil(w / 16) words (distmap.cu)
        const int wpr = (t.roi_w + 15) >> 4;
        std::vector<uint32_t> packed((size_t)wpr * t.roi_h);
        CSB_CUDA(c, cudaMemcpy(packed.data(), d.d_cmap.as<uint8_t>() + 4 * (size_t)t.map_offset, 4 * packed.size(), cudaMemcpyDeviceToHost));
        for (int r = 0; r < t.roi_h; r++)
            for (int col = 0; col < t.roi_w; col++) edges_out[(size_t)r * t.roi_w + col] = (uint8_t)((packed[(size_t)r * wpr + (col >> 4)] >> (2 * (col & 15))) & 3u);
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
    CSB_CUDA(c, cudaMemcpy(&nv, d.B.n_valid + task_id, 4, cudaMemcpyDeviceToHost));
    CSB_CUDA(c, cudaMemcpy(&nm, d.d_n_merged.as<int>() + task_id, 4, cudaMemcpyDeviceToHost));
    CSB_CUDA(c, cudaMemcpy(&nk, d.B.n_keep + task_id, 4, cudaMemcpyDeviceToHost));
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
