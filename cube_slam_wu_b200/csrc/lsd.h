// lsd.h -- device state of the LSD line detector (SURVEY.md 8 "next" row f-2; csrc/lsd.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/cubeslam_b200.h"

namespace csb {

struct LsdState;  // defined in lsd.cu (owns DevBuf/HostBuf members)
void lsd_release(LsdState*& s);

}  // namespace csb
