// context.h -- the opaque csb_context: device, stream, grow-only device workspaces of both halves.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "ba.h"
#include "csb_internal.h"
#include "edlines.h"
#include "lbd.h"
#include "lsd.h"
#include "proposal.h"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // like ensure(); a (re)allocated block is cleared once.  For output arrays that go to the host in strided blocks (rows beyond a frame's
    // count, padding): the host trims them, but every byte it receives is a defined one.
    cudaError_t ensure_zeroed(size_t bytes, cudaStream_t st) {
        const size_t before = cap;
        cudaError_t e = ensure(bytes);
        if (e == cudaSuccess && cap != before) e = cudaMemsetAsync(p, 0, cap, st);
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// pinned host staging memory (grows, never shrinks)
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct DetectState {
    bool uploaded = false, ran = false;
    int n_frames = 0, n_boxes = 0, n_lines = 0, n_tasks = 0;
    int max_groups = 0, max_lines_per_frame = 0, max_hyp_per_task = 0, map_cap_floats = 0, max_roi_w = 0, max_roi_px = 0;
    int64_t n_map_floats = 0, out_total = 0, line_cap_total = 0;
    csb_detect_params params{};
    std::vector<csb::TaskTab> ttab;
    std::vector<csb::FrameTab> ftab;
    std::vector<csb_task> tasks;
    csb::DetectBuffers B{};
    // d_tables: frame/task/order/box/line tables, one H2D copy from h_tables.  d_results: cuboids | n_cuboids | n_valid | n_keep, one D2H
    // copy into h_results.
    DevBuf d_tables, d_results, d_maps, d_ml_seg, d_ml_ang, d_ml_mid, d_n_merged, d_vp_sup, d_p_dist, d_p_angle, d_p_hyp, d_keep, d_norm, d_cand_score, d_cand_ok,
        d_sel_idx, d_sel_flag, d_sel_heap, d_rank_idx, d_counters, d_dbg, d_gray, d_cmap, d_queue, d_dtmp, d_flags, d_segbits;
    HostBuf h_tables, h_results;
    size_t res_off_ncub = 0, res_off_nvalid = 0, res_off_nkeep = 0, res_off_misc = 0, res_bytes = 0;
    const uint8_t* gray_mapped = nullptr;  // device alias of the caller's pinned gray buffer (csb_detect_upload_gray), or NULL: copied as a whole
    bool gray_gather_enabled = true;       // csb_set_option(CSB_OPT_GRAY_GATHER)
    cudaEvent_t ev_tables = nullptr;  // h_tables consumed by the device
    cudaEvent_t ev_order = nullptr;   // compute stream -> copy stream ordering of the streamed map upload
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // gray mode: the line kernels run on the copy stream beside k_distmap
    bool gray_mode = false;
    unsigned epoch = 0;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool timed_last = false;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int launches_last = 0;
};

struct csb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // chunked host->device streaming of csb_detect_batch
    unsigned* h_epoch = nullptr;         // pinned word copied behind every chunk
    bool own_stream = false;
    int num_sms = 0, max_smem_optin = 0;
    std::string err;
    DetectState det;
    csb::BAState ba;
    csb::LsdState* lsd = nullptr;  // created by the first csb_lsd_* call
    csb::LbdState* lbd = nullptr;  // created by the first csb_lbd_* call
    csb::EdState* edlines = nullptr;  // created by the first csb_edlines_* call
    int blur_generation = 4;          // csb_set_blur_generation: integer taps of the 8-bit 5x5 Gaussian of the LBD / EDLines front end
};

#define CSB_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                \
            return CSB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)
