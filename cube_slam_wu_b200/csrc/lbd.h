// lbd.h -- device state of the LBD line descriptor (SURVEY.md 8 "next" row f-2, descriptor half; csrc/lbd.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/cubeslam_b200.h"

namespace csb {

struct LbdState;  // defined in lbd.cu
void lbd_release(LbdState*& s);

// blur 5x5 + Sobel of n_frames gray frames -> {dx, dy} int16 pairs (k_lbd_grad4 / k_lbd_grad); also the first stage of EDLines
void lbd_launch_grad(const uint8_t* gray, short2* grad, int w, int h, int n_frames, cudaStream_t st, int blur_generation);

// the constant-memory band weights of k_lbd_describe (synchronises the stream)
cudaError_t lbd_upload_weights(cudaStream_t st);
// descriptors of detector-supplied key lines (EDLines: direction, numOfPixels, unclamped end points) with the warp-cooperative kernel
cudaError_t lbd_describe_keylines(const short2* grad, const float* lines, const float2* keyl_in, const int* counts, int n_frames, int stride, int w, int h,
                                  uint8_t* desc, float* descf, int* prefix, unsigned long long* ctr, int num_sms, cudaStream_t st);

}  // namespace csb
