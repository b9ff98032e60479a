// lbd.h -- device state of the LBD line descriptor (SURVEY.md 8 "next" row f-2, descriptor half; csrc/lbd.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/cubeslam_b200.h"

namespace csb {

struct LbdState;  // defined in lbd.cu
void lbd_release(LbdState*& s);

}  // namespace csb
