// proposal.h -- device buffer table + kernel launchers of the proposal half.
#pragma once
#include <cuda_runtime.h>

#include "csb_internal.h"

namespace csb {

// All pointers are device pointers.  Per-proposal arrays are indexed by TaskTab::out_offset + i (capacity n_hyp per task),
// merged-line arrays by TaskTab::line_cap_offset + i (capacity = line count of the task's frame).
struct DetectBuffers {
    const FrameTab* ftab;
    const TaskTab* ttab;
    const int* task_order;      // tasks sorted by decreasing n_hyp (persistent-CTA queue)
    const int* box_task_begin;  // n_boxes + 1
    const double* lines;        // raw frame lines, x1 y1 x2 y2
    const float* maps;          // packed distance maps
    int n_tasks;
    int pad;
    // k_prep_lines outputs
    double* ml_seg;  // 4 per merged line
    double* ml_ang;
    double* ml_mid;  // 2 per merged line
    int* n_merged;
    double* vp_sup;      // 6 doubles per (task, group): VP-support angles low/top of vp1, vp2, vp3; task stride sup_stride
    long long sup_stride;
    // k_score outputs (compacted valid proposals in enumeration order)
    double* p_dist;
    double* p_angle;
    int* p_hyp;
    int* n_valid;
    // k_select outputs
    int* keep;
    double* norm_score;
    int* n_keep;
    // k_recover outputs (indexed by position in the kept list)
    double* cand_score;
    unsigned char* cand_ok;
    int* sel_idx;             // 2x capacity, global fallback for very large tasks
    unsigned char* sel_flag;
    double* sel_heap;         // 2 doubles per slot (value,index pairs), global fallback
    // k_rank
    int* rank_idx;            // 2x capacity
    csb_cuboid* cuboids;
    int* n_cuboids;
    int* counters;  // [0]: k_score task queue head
    // chunked host->device streaming of the distance maps (csb_detect_batch): NULL when everything is resident before launch
    const unsigned* ready_flags;
    unsigned epoch;
    int n_chunks;
    long long chunk_end[CSB_MAX_CHUNKS];  // slice k holds map floats [chunk_end[k-1], chunk_end[k]); a task waits for the slice of its last float
    DetectConst dc;
};

cudaError_t launch_prep_lines(const DetectBuffers& B, int max_lines_per_frame, int max_groups, cudaStream_t st);
cudaError_t launch_score(const DetectBuffers& B, int max_groups, int max_hyp_per_task, int num_sms, int max_smem_optin, int* map_cap_floats_out, cudaStream_t st);
cudaError_t launch_select(const DetectBuffers& B, int max_hyp_per_task, int max_smem_optin, cudaStream_t st, int* n_launches);
cudaError_t launch_recover(const DetectBuffers& B, cudaStream_t st);
cudaError_t launch_rank(const DetectBuffers& B, int n_boxes, cudaStream_t st);
cudaError_t launch_distmaps(const DetectBuffers& B, const uint8_t* gray, uint8_t* cmap, int* queue, unsigned* dtmp, float* maps, int max_roi_w, int max_pm_words, cudaStream_t st);
// granularity of the ROI-segment upload of the gray frames (bytes; a power of two, 16 .. 128): see distmap.cu
constexpr int CSB_GRAY_SEG = 32;
cudaError_t launch_gray_gather(const DetectBuffers& B, const uint8_t* gray_host_mapped, uint8_t* gray_dev, long long n_bytes, unsigned* seg_bits, int* n_segments_out,
                               int num_sms, cudaStream_t st);
cudaError_t score_phase_cycles(unsigned long long* out12, bool reset);
cudaError_t launch_debug_atan2(const double* y, const double* x, double* out, int n6, int* n_fallback, cudaStream_t st);
cudaError_t launch_debug_corners(const DetectBuffers& B, int task, int n_valid, double* out, cudaStream_t st);

}  // namespace csb
