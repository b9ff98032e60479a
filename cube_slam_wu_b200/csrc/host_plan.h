// host_plan.h -- host planner interface (proposal half).
#pragma once
#include <vector>

#include "csb_internal.h"

namespace csb {
int frame_group_count(const csb_frame& f, const csb_detect_params& p, int* n_groups);
int build_frame_tab(const csb_frame& f, const csb_detect_params& p, FrameTab& ft);
int plan_tasks(const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const csb_detect_params& p, std::vector<csb_task>& tasks,
               std::vector<TaskTab>* tabs, int64_t* n_map_floats);
}  // namespace csb
