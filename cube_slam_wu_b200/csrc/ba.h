// ba.h -- device state + launchers of the BA half.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "csb_internal.h"

namespace csb {

// Device view of the camera-object graph and of one linearisation's outputs.
struct BABuffers {
    int n_cam, n_cube, n_ec, n_ep, n_eo, pad;
    const double* cams7;   // n_cam x 7
    const double* cubes10; // n_cube x 10
    const int *cam_fixed, *cube_fixed;
    const int *ec_cam, *ec_cube; const double *ec_meas, *ec_info;
    const int *ep_cam, *ep_cube; const double *ep_meas, *ep_info, *ep_K;
    const int *eo_i, *eo_j; const double *eo_meas, *eo_info;
    // per-edge results
    double *ec_err, *ep_err, *eo_err;
    double *ec_Ji, *ec_Jj, *ep_Ji, *ep_Jj, *eo_Ji, *eo_Jj;  // optional (NULL = not materialised)
    double *ec_Hij, *ep_Hij, *eo_Hij;
    double* contrib;   // per-edge diagonal-block contributions: [Hii | bi | Hjj | bj] per edge, see ba.cu
    double* edge_chi2; // n_ec + n_ep + n_eo
    // per-vertex adjacency (CSR): entries are offsets into contrib
    const int *cam_adj_ptr, *cube_adj_ptr;
    const int64_t *cam_adj_H, *cam_adj_b, *cube_adj_H, *cube_adj_b;
    double *H_cam, *b_cam, *H_cube, *b_cube, *chi2;
};

// host copy of the topology (structure of the reduced camera system is built from it on the first csb_ba_optimize)
struct HostGraph {
    std::vector<int> cam_fixed, cube_fixed, ec_cam, ec_cube, ep_cam, ep_cube, eo_i, eo_j;
};

// a device array that grows geometrically (csb_ba_add_frame appends without reallocating every time)
struct GrowBuf { void* p = nullptr; size_t cap = 0; };
constexpr int BA_MAX_BUFS = 48;

struct BAState {
    bool has_graph = false, has_estimates = false, ran = false;
    bool have_J = false;    // the last linearisation materialised the Jacobians
    GrowBuf dev[BA_MAX_BUFS];
    bool analytic = false;  // csb_ba_set_jacobian_mode
    HostGraph host;
    void* solver = nullptr;  // SolveState of ba_solve.cu
    int n_cam = 0, n_cube = 0, n_ec = 0, n_ep = 0, n_eo = 0;
    BABuffers B{};
    std::vector<void*> allocs;  // everything cudaMalloc'ed for the current graph
    int launches_last = 0;
};

void ba_release(BAState& s);
void ba_solver_release(BAState& s);
cudaError_t ba_launch(const BABuffers& B, bool want_jacobians, cudaStream_t st, int* n_launches, bool analytic);

}  // namespace csb
