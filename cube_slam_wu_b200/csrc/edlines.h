// edlines.h -- device state of the EDLines line detector (SURVEY.md 8 "next" row f-2, use_LSD = false; csrc/edlines.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/cubeslam_b200.h"

namespace csb {

struct EdState;  // defined in edlines.cu
void edlines_release(EdState*& s);

}  // namespace csb
