// lsd.h -- device state of the LSD line detector (SURVEY.md 8 "next" row f-2; csrc/lsd.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/cubeslam_b200.h"

namespace csb {

struct LsdState;  // defined in lsd.cu (owns DevBuf/HostBuf members)
void lsd_release(LsdState*& s);

// what the descriptor stage (lbd.cu) needs of the last csb_lsd_run: the resident frames and the segment table, both on the device
struct LsdView {
    const uint8_t* gray;  // n_frames x h x w
    const float* lines;   // n_frames x max_lines x 4
    const int* n_lines;   // per frame (may exceed max_lines when a frame overflowed)
    int w, h, n_frames, max_lines;
};
bool lsd_view(LsdState* s, LsdView& v);  // false before the first run

}  // namespace csb
