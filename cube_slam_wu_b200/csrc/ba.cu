// ba.cu -- sm_100a kernels + C ABI of the BA half: residuals, central-difference Jacobians and the block normal
// equations of the camera-cuboid graph (what g2o's BlockSolver::buildSystem() does edge by edge).
//
//   k_linearize<TYPE> : a warp per edge.  Lane c < Di+Dj evaluates the residual at +delta along tangent direction c of
//                       vertex 0 / vertex 1, lane c + 16 at -delta (the 30 evaluations of BaseBinaryEdge::linearizeOplus,
//                       base_binary_edge.hpp:130-205, delta = 1e-9: one per lane); lane c keeps Jacobian column c in
//                       registers; lane 15 evaluates the unperturbed residual.  (Closed-form Jacobians: half a warp per edge.)  The 186-double quadratic form of the edge
//                       (constructQuadraticForm, base_binary_edge.hpp:54-120) is formed with warp shuffles and stored
//                       as one contiguous record; the off-diagonal block A^T Omega B goes straight to its BSR slot.
//   k_gather          : one warp per vertex sums the records of its incident edges in g2o's edge order
//                       (deterministic: no atomics) into the diagonal blocks H_vv and b_v.
//
// Reference: object_slam/include/object_slam/g2o_Object.h:23-292, Thirdparty/g2o/g2o/types/se3quat.h,
// types_six_dof_expmap.h:59-99, core/block_solver.hpp:501-560.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "ba.h"
#include "ba_dev.cuh"
#include "context.h"
#include "csb_math.cuh"

namespace csb {

constexpr int LIN_THREADS = 128;  // 8 edges per CTA

// ANALYTIC: the Jacobian columns of EdgeSE3Cuboid / EdgeSE3Expmap in closed form instead of the reference's 30
// residual evaluations.  With Delta = (Twc M R_k)^-1 C (k = the yaw variant min_log_error picks) and xi = log(Delta) = e[0:6]:
//   camera   T <- exp(d) T      : Delta <- exp(Ad_A d) Delta,  A = (M R_k)^-1     =>  de/dd = [ Jl^-1(xi) Ad_A ; 0 ]
//   cuboid   C <- C exp(d[0:6]) : Delta <- Delta exp(d)                           =>  de/dd = blockdiag( Jl^-1(-xi), I_3 )
//   odometry e = log(M Ti Tj^-1): de/ddi = Jl^-1(xi) Ad_M,  de/ddj = -Jl^-1(-xi)
// EdgeSE3CuboidProj keeps central differences.  Validated against the numeric Jacobians (tests/test_ba_gpu.py).
template <int TYPE, bool ANALYTIC>
__global__ void __launch_bounds__(LIN_THREADS) k_linearize(BABuffers B, int n_edges, int64_t rec_base, int chi_base, double* Ji_out, double* Jj_out) {
    constexpr int D = EdgeDims<TYPE>::D, Di = EdgeDims<TYPE>::Di, Dj = EdgeDims<TYPE>::Dj, REC = EdgeDims<TYPE>::REC;
    // numeric Jacobians: a whole warp per edge -- lane c < 16 evaluates the residual at +delta along tangent c, lane c + 16 at -delta
    // (one residual evaluation per lane); closed form: half a warp per edge
    constexpr int LPE = ANALYTIC ? 16 : 32;
    const int tid = blockIdx.x * LIN_THREADS + threadIdx.x;
    const int e = tid / LPE, c = tid & 15;          // edge, column / role
    const int lane = threadIdx.x & 31, half_base = (LPE == 16) ? (lane & 16) : 0;
    const bool neg = (LPE == 32) && (lane & 16);    // the -delta half of the warp
    const unsigned FULL = 0xffffffffu;
    const bool in_range = e < n_edges;
    const bool active = in_range && !neg;           // lanes that own a row of the quadratic form / write results
    const int ee = in_range ? e : 0;

    // ---- load
    EdgeCtx x;
    int vi, vj;
    const double* info;
    bool i_free, j_free;
    if (TYPE == EDGE_CUBOID) {
        vi = B.ec_cam[ee]; vj = B.ec_cube[ee];
        x.meas_cube = cube_from_vec10(B.ec_meas + 10 * (size_t)ee);
        info = B.ec_info + 81 * (size_t)ee;
    } else if (TYPE == EDGE_PROJ) {
        vi = B.ep_cam[ee]; vj = B.ep_cube[ee];
        for (int k = 0; k < 4; k++) x.meas4[k] = B.ep_meas[4 * (size_t)ee + k];
        x.K = B.ep_K + 9 * (size_t)ee;
        info = B.ep_info + 16 * (size_t)ee;
    } else {
        vi = B.eo_i[ee]; vj = B.eo_j[ee];
        x.meas_se3 = se3_from_vec7(B.eo_meas + 7 * (size_t)ee);
        info = B.eo_info + 36 * (size_t)ee;
    }
    x.cam = se3_from_vec7(B.cams7 + 7 * (size_t)vi);
    i_free = !B.cam_fixed[vi];
    if (TYPE == EDGE_ODOM) { x.cam2 = se3_from_vec7(B.cams7 + 7 * (size_t)vj); j_free = !B.cam_fixed[vj]; x.cube = Cube{}; }
    else { x.cube = cube_from_vec10(B.cubes10 + 10 * (size_t)vj); j_free = !B.cube_fixed[vj]; x.cam2 = x.cam; }

    // ---- residuals: lane 15 -> base error; lane c -> Jacobian column c by central differences
    double J[D];
#pragma unroll
    for (int k = 0; k < D; k++) J[k] = 0;
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    if (active && ANALYTIC && TYPE != EDGE_PROJ) {
        double e0[D];
        SE3 A;  // the fixed left factor of Delta
        if (TYPE == EDGE_CUBOID) {
            Cube esti;
            esti.pose = se3_mul(se3_inverse(x.cam), x.meas_cube.pose);
            esti.scale = x.meas_cube.scale;
            SE3 best;
            cube_min_log_error_k(x.cube, esti, e0, best);  // best = Twc M R_k
            A = se3_mul(se3_inverse(best), se3_inverse(x.cam));  // (M R_k)^-1 = best^-1 Twc
        } else {
            edge_error<TYPE>(x, x.cam, x.cube, x.cam2, e0);
            A = x.meas_se3;
        }
        if (c == 15) {
#pragma unroll
            for (int k = 0; k < D; k++) J[k] = e0[k];
        } else if (c < Di) {
            if (i_free) {
                double col[6], o[6];
                se3_adjoint_col(A, c, col);
                se3_jl_inv_apply(e0, col, o);
#pragma unroll
                for (int k = 0; k < 6; k++) J[k] = o[k];
            }
        } else if (c < Di + Dj) {
            if (j_free) {
                const int dd = c - Di;
                if (dd < 6) {
                    double mxi[6], col[6], o[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) { mxi[k] = -e0[k]; col[k] = (k == dd) ? 1.0 : 0.0; }
                    se3_jl_inv_apply(mxi, col, o);
#pragma unroll
                    for (int k = 0; k < 6; k++) J[k] = (TYPE == EDGE_ODOM) ? -o[k] : o[k];
                } else if (TYPE == EDGE_CUBOID) {
                    J[dd] = 1.0;  // scale rows: e[6:9] = scale - (swapped) measured scale
                }
            }
        }
    } else if (in_range) {
        const double dl = neg ? -delta : delta;
        double ev[D];
#pragma unroll
        for (int k = 0; k < D; k++) ev[k] = 0;
        if (c == 15) {
            if (!neg) edge_error<TYPE>(x, x.cam, x.cube, x.cam2, ev);  // the error vector, on lane 15
        } else if (c < Di) {
            if (i_free) {
                double add[6];
#pragma unroll
                for (int q = 0; q < 6; q++) add[q] = (q == c) ? dl : 0.0;
                edge_error<TYPE>(x, se3_mul(se3_exp(add), x.cam), x.cube, x.cam2, ev);  // VertexSE3Expmap::oplusImpl: exp(d) * T
            }
        } else if (c < Di + Dj) {
            if (j_free) {
                const int d = c - Di;
                if (TYPE == EDGE_ODOM) {
                    double add[6];
#pragma unroll
                    for (int q = 0; q < 6; q++) add[q] = (q == d) ? dl : 0.0;
                    edge_error<TYPE>(x, x.cam, x.cube, se3_mul(se3_exp(add), x.cam2), ev);
                } else {
                    double add[9];
#pragma unroll
                    for (int q = 0; q < 9; q++) add[q] = (q == d) ? dl : 0.0;
                    edge_error<TYPE>(x, x.cam, cube_exp_update(x.cube, add), x.cam2, ev);  // VertexCuboid::oplusImpl
                }
            }
        }
        // central difference (base_binary_edge.hpp:147-160): the -delta evaluation comes from lane + 16
#pragma unroll
        for (int k = 0; k < D; k++) {
            const double em = __shfl_down_sync(FULL, ev[k], 16);
            J[k] = (c == 15) ? ev[k] : scalar * (ev[k] - em);
        }
        if (neg) {
#pragma unroll
            for (int k = 0; k < D; k++) J[k] = 0;
        }
    }

    // ---- error, chi2, omega_r = -Omega * e (every lane of the half-warp gets e from lane 15)
    double err[D], omega_r[D];
#pragma unroll
    for (int k = 0; k < D; k++) err[k] = __shfl_sync(FULL, J[k], half_base + 15);
#pragma unroll
    for (int r = 0; r < D; r++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < D; k++) s += info[r * D + k] * err[k];
        omega_r[r] = -s;
    }
    if (active && c == 15) {
        double* eo = (TYPE == EDGE_CUBOID ? B.ec_err : TYPE == EDGE_PROJ ? B.ep_err : B.eo_err) + (size_t)D * e;
        double chi = 0;
#pragma unroll
        for (int k = 0; k < D; k++) { eo[k] = err[k]; chi += err[k] * (-omega_r[k]); }
        B.edge_chi2[chi_base + e] = chi;
    }
    const bool is_col = c < Di + Dj;
    if (active && is_col) {
        if (c < Di && Ji_out) for (int k = 0; k < D; k++) Ji_out[(size_t)e * D * Di + c * D + k] = J[k];
        if (c >= Di && Jj_out) for (int k = 0; k < D; k++) Jj_out[(size_t)e * D * Dj + (c - Di) * D + k] = J[k];
    }

    // ---- quadratic form.  Lane i owns row i of [A B]^T Omega [A B]:  JtO_i[k] = sum_m J_i[m] * Omega[m][k]
    double JtO[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        double s = 0;
#pragma unroll
        for (int m = 0; m < D; m++) s += J[m] * info[m * D + k];
        JtO[k] = s;
    }
    double brow = 0;
#pragma unroll
    for (int k = 0; k < D; k++) brow += J[k] * omega_r[k];
    double* rec = B.contrib + rec_base + (size_t)REC * ee;
    double* Hij = (TYPE == EDGE_CUBOID ? B.ec_Hij : TYPE == EDGE_PROJ ? B.ep_Hij : B.eo_Hij) + (size_t)Di * Dj * ee;
    const bool row_i = c < Di, row_j = (c >= Di) && (c < Di + Dj);
#pragma unroll 1
    for (int cc = 0; cc < Di + Dj; cc++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < D; k++) s += JtO[k] * __shfl_sync(FULL, J[k], half_base + cc);
        if (!active) continue;
        if (row_i) {
            if (cc < Di) rec[cc * Di + c] = i_free ? s : 0.0;                              // Hii (col-major)
            else Hij[(cc - Di) * Di + c] = (i_free && j_free) ? s : 0.0;                    // A^T Omega B
        } else if (row_j) {
            if (cc >= Di) rec[Di * Di + Di + (cc - Di) * Dj + (c - Di)] = j_free ? s : 0.0;  // Hjj
        }
    }
    if (active) {
        if (row_i) rec[Di * Di + c] = i_free ? brow : 0.0;
        else if (row_j) rec[Di * Di + Di + Dj * Dj + (c - Di)] = j_free ? brow : 0.0;
    }
}

// one warp per vertex; dim = 6 (cameras) or 9 (cuboids)
__global__ void __launch_bounds__(128) k_gather(BABuffers B, int n_vertices, int dim, const int* adj_ptr, const int64_t* adj_H, const int64_t* adj_b, double* H, double* b) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= n_vertices) return;
    const int a0 = adj_ptr[v], a1 = adj_ptr[v + 1];
    const int nn = dim * dim;
    double acc[3] = {0, 0, 0}, accb = 0;
    // the sum stays in edge order (deterministic, same order as g2o's sequential accumulation); the loads of U records are
    // issued together so that their latency overlaps
    constexpr int U = 8;
    for (int a = a0; a < a1; a += U) {
        int64_t oh[U], obb[U];
#pragma unroll
        for (int u = 0; u < U; u++) { int aa = (a + u < a1) ? a + u : a1 - 1; oh[u] = adj_H[aa]; obb[u] = adj_b[aa]; }
        double hv[U][3], bv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int q = 0; q < 3; q++) { int idx = lane + 32 * q; hv[u][q] = (idx < nn) ? B.contrib[oh[u] + idx] : 0.0; }
            bv[u] = (lane < dim) ? B.contrib[obb[u] + lane] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (a + u < a1) {
#pragma unroll
                for (int q = 0; q < 3; q++) acc[q] += hv[u][q];
                accb += bv[u];
            }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) { int idx = lane + 32 * q; if (idx < nn) H[(size_t)v * nn + idx] = acc[q]; }
    if (lane < dim) b[(size_t)v * dim + lane] = accb;
}

// deterministic chi2 reduction: one warp, fixed order
__global__ void __launch_bounds__(32) k_chi2(const double* edge_chi2, int n, double* out) {
    const unsigned FULL = 0xffffffffu;
    double s = 0;
    for (int i = threadIdx.x; i < n; i += 32) s += edge_chi2[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(FULL, s, o);
    if (threadIdx.x == 0) *out = s;
}

cudaError_t ba_launch(const BABuffers& B, bool want_J, cudaStream_t st, int* n_launches, bool analytic) {
    int L = 0;
    auto grid = [](int n_edges, int lanes_per_edge) { return (int)(((long long)n_edges * lanes_per_edge + LIN_THREADS - 1) / LIN_THREADS); };
    if (B.n_ec) {
        if (analytic) k_linearize<EDGE_CUBOID, true><<<grid(B.n_ec, 16), LIN_THREADS, 0, st>>>(B, B.n_ec, 0, 0, want_J ? B.ec_Ji : nullptr, want_J ? B.ec_Jj : nullptr);
        else k_linearize<EDGE_CUBOID, false><<<grid(B.n_ec, 32), LIN_THREADS, 0, st>>>(B, B.n_ec, 0, 0, want_J ? B.ec_Ji : nullptr, want_J ? B.ec_Jj : nullptr);
        L++;
    }
    if (B.n_ep) { k_linearize<EDGE_PROJ, false><<<grid(B.n_ep, 32), LIN_THREADS, 0, st>>>(B, B.n_ep, (int64_t)132 * B.n_ec, B.n_ec, want_J ? B.ep_Ji : nullptr, want_J ? B.ep_Jj : nullptr); L++; }
    if (B.n_eo) {
        if (analytic) k_linearize<EDGE_ODOM, true><<<grid(B.n_eo, 16), LIN_THREADS, 0, st>>>(B, B.n_eo, (int64_t)132 * (B.n_ec + B.n_ep), B.n_ec + B.n_ep, want_J ? B.eo_Ji : nullptr, want_J ? B.eo_Jj : nullptr);
        else k_linearize<EDGE_ODOM, false><<<grid(B.n_eo, 32), LIN_THREADS, 0, st>>>(B, B.n_eo, (int64_t)132 * (B.n_ec + B.n_ep), B.n_ec + B.n_ep, want_J ? B.eo_Ji : nullptr, want_J ? B.eo_Jj : nullptr);
        L++;
    }
    if (B.n_cam) { k_gather<<<(B.n_cam * 32 + 127) / 128, 128, 0, st>>>(B, B.n_cam, 6, B.cam_adj_ptr, B.cam_adj_H, B.cam_adj_b, B.H_cam, B.b_cam); L++; }
    if (B.n_cube) { k_gather<<<(B.n_cube * 32 + 127) / 128, 128, 0, st>>>(B, B.n_cube, 9, B.cube_adj_ptr, B.cube_adj_H, B.cube_adj_b, B.H_cube, B.b_cube); L++; }
    k_chi2<<<1, 32, 0, st>>>(B.edge_chi2, B.n_ec + B.n_ep + B.n_eo, B.chi2); L++;
    if (n_launches) *n_launches = L;
    return cudaGetLastError();
}

void ba_release(BAState& s) {
    ba_solver_release(s);
    for (void* p : s.allocs) cudaFree(p);
    s.allocs.clear();
    for (GrowBuf& g : s.dev) { if (g.p) cudaFree(g.p); g.p = nullptr; g.cap = 0; }
    s.host = HostGraph();
    s.n_cam = s.n_cube = s.n_ec = s.n_ep = s.n_eo = 0;
    s.has_graph = s.has_estimates = s.ran = s.have_J = false;
}

}  // namespace csb

using namespace csb;

#define CSB_TRY(x) do { int rc__ = (x); if (rc__ != CSB_OK) return rc__; } while (0)

namespace {

enum {
    G_CAM_FIXED, G_CUBE_FIXED, G_EC_CAM, G_EC_CUBE, G_EC_MEAS, G_EC_INFO, G_EP_CAM, G_EP_CUBE, G_EP_MEAS, G_EP_INFO, G_EP_K, G_EO_I, G_EO_J, G_EO_MEAS, G_EO_INFO,
    G_CAM_ADJ_PTR, G_CAM_ADJ_H, G_CAM_ADJ_B, G_CUBE_ADJ_PTR, G_CUBE_ADJ_H, G_CUBE_ADJ_B, G_CAMS7, G_CUBES10, G_EC_ERR, G_EP_ERR, G_EO_ERR, G_EC_JI, G_EC_JJ, G_EP_JI,
    G_EP_JJ, G_EO_JI, G_EO_JJ, G_EC_HIJ, G_EP_HIJ, G_EO_HIJ, G_CONTRIB, G_EDGE_CHI2, G_H_CAM, G_B_CAM, G_H_CUBE, G_B_CUBE, G_CHI2, G_COUNT
};
static_assert(G_COUNT <= BA_MAX_BUFS, "BA_MAX_BUFS");

// capacity >= bytes; the first `keep` bytes survive a reallocation (geometric growth: appending a frame rarely reallocates)
int ensure(csb_context* c, int slot, size_t bytes, size_t keep) {
    GrowBuf& g = c->ba.dev[slot];
    if (g.p && bytes <= g.cap) return CSB_OK;
    const size_t ncap = std::max<size_t>(std::max<size_t>(bytes, 2 * g.cap), 256);
    void* np = nullptr;
    CSB_CUDA(c, cudaMalloc(&np, ncap));
    if (g.p) {
        if (keep) CSB_CUDA(c, cudaMemcpyAsync(np, g.p, keep, cudaMemcpyDeviceToDevice, c->stream));
        CSB_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(g.p);
    }
    g.p = np; g.cap = ncap;
    return CSB_OK;
}
// device copy of a host vector whose first n_old elements are already there
template <class T>
int sync_tail(csb_context* c, int slot, const std::vector<T>& h, size_t n_old) {
    CSB_TRY(ensure(c, slot, h.size() * sizeof(T), n_old * sizeof(T)));
    if (h.size() > n_old)
        CSB_CUDA(c, cudaMemcpyAsync(reinterpret_cast<T*>(c->ba.dev[slot].p) + n_old, h.data() + n_old, (h.size() - n_old) * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return CSB_OK;
}
template <class T>
int sync_all(csb_context* c, int slot, const std::vector<T>& h) { return sync_tail(c, slot, h, 0); }

struct Counts { int n_cam, n_cube, n_ec, n_ep, n_eo; };
// measurements / information matrices of the NEW edges (caller memory; the host keeps the topology only)
struct NewEdgeData { const double *ec_meas, *ec_info, *ep_meas, *ep_info, *ep_K, *eo_meas, *eo_info; };

// device array of n_new records of `rec` doubles appended behind the n_old it already holds
int append_dev(csb_context* c, int slot, const double* src, size_t n_old, size_t n_new, size_t rec) {
    CSB_TRY(ensure(c, slot, (n_old + n_new) * rec * 8, n_old * rec * 8));
    if (n_new) CSB_CUDA(c, cudaMemcpyAsync(reinterpret_cast<double*>(c->ba.dev[slot].p) + n_old * rec, src, n_new * rec * 8, cudaMemcpyHostToDevice, c->stream));
    return CSB_OK;
}

// Brings the device state in line with c->ba.host after vertices / edges were appended (old = what the device already holds):
// tails of the topology / measurement arrays, the per-vertex adjacency (rebuilt: the record offsets of the later edge types move when
// an edge is inserted), the work buffers; the estimates of the existing vertices are preserved.
int sync_device(csb_context* c, const Counts& old, const NewEdgeData& nd) {
    BAState& s = c->ba;
    const HostGraph& h = s.host;
    const int n_cam = (int)h.cam_fixed.size(), n_cube = (int)h.cube_fixed.size(), n_ec = (int)h.ec_cam.size(), n_ep = (int)h.ep_cam.size(), n_eo = (int)h.eo_i.size();
    CSB_TRY(sync_tail(c, G_CAM_FIXED, h.cam_fixed, old.n_cam)); CSB_TRY(sync_tail(c, G_CUBE_FIXED, h.cube_fixed, old.n_cube));
    CSB_TRY(sync_tail(c, G_EC_CAM, h.ec_cam, old.n_ec)); CSB_TRY(sync_tail(c, G_EC_CUBE, h.ec_cube, old.n_ec));
    CSB_TRY(append_dev(c, G_EC_MEAS, nd.ec_meas, old.n_ec, n_ec - old.n_ec, 10)); CSB_TRY(append_dev(c, G_EC_INFO, nd.ec_info, old.n_ec, n_ec - old.n_ec, 81));
    CSB_TRY(sync_tail(c, G_EP_CAM, h.ep_cam, old.n_ep)); CSB_TRY(sync_tail(c, G_EP_CUBE, h.ep_cube, old.n_ep));
    CSB_TRY(append_dev(c, G_EP_MEAS, nd.ep_meas, old.n_ep, n_ep - old.n_ep, 4)); CSB_TRY(append_dev(c, G_EP_INFO, nd.ep_info, old.n_ep, n_ep - old.n_ep, 16));
    CSB_TRY(append_dev(c, G_EP_K, nd.ep_K, old.n_ep, n_ep - old.n_ep, 9));
    CSB_TRY(sync_tail(c, G_EO_I, h.eo_i, old.n_eo)); CSB_TRY(sync_tail(c, G_EO_J, h.eo_j, old.n_eo));
    CSB_TRY(append_dev(c, G_EO_MEAS, nd.eo_meas, old.n_eo, n_eo - old.n_eo, 7)); CSB_TRY(append_dev(c, G_EO_INFO, nd.eo_info, old.n_eo, n_eo - old.n_eo, 36));

    // buildStructure(): per-vertex adjacency in edge order ec, ep, eo (block_solver.hpp:142-295 allocates the blocks;
    // sparse_optimizer.cpp:482-487 fixes the edge order) -- counting sort, flat arrays
    {
        std::vector<int> cam_ptr(n_cam + 1, 0), cube_ptr(n_cube + 1, 0);
        for (int e = 0; e < n_ec; e++) { cam_ptr[h.ec_cam[e] + 1]++; cube_ptr[h.ec_cube[e] + 1]++; }
        for (int e = 0; e < n_ep; e++) { cam_ptr[h.ep_cam[e] + 1]++; cube_ptr[h.ep_cube[e] + 1]++; }
        for (int e = 0; e < n_eo; e++) { cam_ptr[h.eo_i[e] + 1]++; cam_ptr[h.eo_j[e] + 1]++; }
        for (int v = 0; v < n_cam; v++) cam_ptr[v + 1] += cam_ptr[v];
        for (int v = 0; v < n_cube; v++) cube_ptr[v + 1] += cube_ptr[v];
        std::vector<int64_t> cam_H(cam_ptr[n_cam]), cam_b(cam_ptr[n_cam]), cube_H(cube_ptr[n_cube]), cube_b(cube_ptr[n_cube]);
        std::vector<int> cam_at(cam_ptr.begin(), cam_ptr.end() - 1), cube_at(cube_ptr.begin(), cube_ptr.end() - 1);
        auto put = [](std::vector<int>& at, std::vector<int64_t>& H, std::vector<int64_t>& b, int v, int64_t oh, int64_t ob) { const int k = at[v]++; H[k] = oh; b[k] = ob; };
        for (int e = 0; e < n_ec; e++) {
            const int64_t r = (int64_t)132 * e;
            put(cam_at, cam_H, cam_b, h.ec_cam[e], r, r + 36); put(cube_at, cube_H, cube_b, h.ec_cube[e], r + 42, r + 42 + 81);
        }
        for (int e = 0; e < n_ep; e++) {
            const int64_t r = (int64_t)132 * (n_ec + e);
            put(cam_at, cam_H, cam_b, h.ep_cam[e], r, r + 36); put(cube_at, cube_H, cube_b, h.ep_cube[e], r + 42, r + 42 + 81);
        }
        for (int e = 0; e < n_eo; e++) {
            const int64_t r = (int64_t)132 * (n_ec + n_ep) + (int64_t)84 * e;
            put(cam_at, cam_H, cam_b, h.eo_i[e], r, r + 36); put(cam_at, cam_H, cam_b, h.eo_j[e], r + 42, r + 42 + 36);
        }
        CSB_TRY(sync_all(c, G_CAM_ADJ_PTR, cam_ptr)); CSB_TRY(sync_all(c, G_CAM_ADJ_H, cam_H)); CSB_TRY(sync_all(c, G_CAM_ADJ_B, cam_b));
        CSB_TRY(sync_all(c, G_CUBE_ADJ_PTR, cube_ptr)); CSB_TRY(sync_all(c, G_CUBE_ADJ_H, cube_H)); CSB_TRY(sync_all(c, G_CUBE_ADJ_B, cube_b));
        CSB_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
    }
    CSB_TRY(ensure(c, G_CAMS7, 56 * (size_t)n_cam, 56 * (size_t)old.n_cam)); CSB_TRY(ensure(c, G_CUBES10, 80 * (size_t)n_cube, 80 * (size_t)old.n_cube));
    auto work = [&](int slot, size_t n_doubles) { return ensure(c, slot, 8 * n_doubles, 0); };
    CSB_TRY(work(G_EC_ERR, (size_t)n_ec * 9)); CSB_TRY(work(G_EP_ERR, (size_t)n_ep * 4)); CSB_TRY(work(G_EO_ERR, (size_t)n_eo * 6));
    CSB_TRY(work(G_EC_JI, (size_t)n_ec * 54)); CSB_TRY(work(G_EC_JJ, (size_t)n_ec * 81)); CSB_TRY(work(G_EP_JI, (size_t)n_ep * 24)); CSB_TRY(work(G_EP_JJ, (size_t)n_ep * 36));
    CSB_TRY(work(G_EO_JI, (size_t)n_eo * 36)); CSB_TRY(work(G_EO_JJ, (size_t)n_eo * 36));
    CSB_TRY(work(G_EC_HIJ, (size_t)n_ec * 54)); CSB_TRY(work(G_EP_HIJ, (size_t)n_ep * 54)); CSB_TRY(work(G_EO_HIJ, (size_t)n_eo * 36));
    CSB_TRY(work(G_CONTRIB, (size_t)132 * (n_ec + n_ep) + (size_t)84 * n_eo)); CSB_TRY(work(G_EDGE_CHI2, (size_t)(n_ec + n_ep + n_eo)));
    CSB_TRY(work(G_H_CAM, (size_t)n_cam * 36)); CSB_TRY(work(G_B_CAM, (size_t)n_cam * 6)); CSB_TRY(work(G_H_CUBE, (size_t)n_cube * 81)); CSB_TRY(work(G_B_CUBE, (size_t)n_cube * 9));
    CSB_TRY(work(G_CHI2, 1));

    s.n_cam = n_cam; s.n_cube = n_cube; s.n_ec = n_ec; s.n_ep = n_ep; s.n_eo = n_eo;
    BABuffers& B = s.B;
    std::memset(&B, 0, sizeof B);
    B.n_cam = n_cam; B.n_cube = n_cube; B.n_ec = n_ec; B.n_ep = n_ep; B.n_eo = n_eo;
    auto I = [&](int slot) { return reinterpret_cast<int*>(s.dev[slot].p); };
    auto D = [&](int slot) { return reinterpret_cast<double*>(s.dev[slot].p); };
    auto L = [&](int slot) { return reinterpret_cast<int64_t*>(s.dev[slot].p); };
    B.cam_fixed = I(G_CAM_FIXED); B.cube_fixed = I(G_CUBE_FIXED);
    B.ec_cam = I(G_EC_CAM); B.ec_cube = I(G_EC_CUBE); B.ec_meas = D(G_EC_MEAS); B.ec_info = D(G_EC_INFO);
    B.ep_cam = I(G_EP_CAM); B.ep_cube = I(G_EP_CUBE); B.ep_meas = D(G_EP_MEAS); B.ep_info = D(G_EP_INFO); B.ep_K = D(G_EP_K);
    B.eo_i = I(G_EO_I); B.eo_j = I(G_EO_J); B.eo_meas = D(G_EO_MEAS); B.eo_info = D(G_EO_INFO);
    B.cam_adj_ptr = I(G_CAM_ADJ_PTR); B.cam_adj_H = L(G_CAM_ADJ_H); B.cam_adj_b = L(G_CAM_ADJ_B);
    B.cube_adj_ptr = I(G_CUBE_ADJ_PTR); B.cube_adj_H = L(G_CUBE_ADJ_H); B.cube_adj_b = L(G_CUBE_ADJ_B);
    B.cams7 = D(G_CAMS7); B.cubes10 = D(G_CUBES10);
    B.ec_err = D(G_EC_ERR); B.ep_err = D(G_EP_ERR); B.eo_err = D(G_EO_ERR);
    B.ec_Ji = D(G_EC_JI); B.ec_Jj = D(G_EC_JJ); B.ep_Ji = D(G_EP_JI); B.ep_Jj = D(G_EP_JJ); B.eo_Ji = D(G_EO_JI); B.eo_Jj = D(G_EO_JJ);
    B.ec_Hij = D(G_EC_HIJ); B.ep_Hij = D(G_EP_HIJ); B.eo_Hij = D(G_EO_HIJ);
    B.contrib = D(G_CONTRIB); B.edge_chi2 = D(G_EDGE_CHI2);
    B.H_cam = D(G_H_CAM); B.b_cam = D(G_B_CAM); B.H_cube = D(G_H_CUBE); B.b_cube = D(G_B_CUBE); B.chi2 = D(G_CHI2);
    ba_solver_release(s);  // the structure of the reduced camera system (and the captured graphs) belong to the old topology
    s.ran = false; s.have_J = false;
    return CSB_OK;
}

bool in_range(const int32_t* idx, int n, int lim) {
    for (int i = 0; i < n; i++) if (idx[i] < 0 || idx[i] >= lim) return false;
    return true;
}

}  // namespace

extern "C" {

int csb_ba_set_graph(csb_context* c, const csb_ba_graph* g) {
    if (!c || !g) return CSB_ERR_INVALID;
    // everything is checked before the current graph is touched
    if (g->n_cam < 0 || g->n_cube < 0 || g->n_ec < 0 || g->n_ep < 0 || g->n_eo < 0) { c->err = "csb_ba_set_graph: negative size"; return CSB_ERR_INVALID; }
    if ((g->n_cam && !g->cam_fixed) || (g->n_cube && !g->cube_fixed) || (g->n_ec && (!g->ec_cam || !g->ec_cube || !g->ec_meas || !g->ec_info)) ||
        (g->n_ep && (!g->ep_cam || !g->ep_cube || !g->ep_meas || !g->ep_info || !g->ep_K)) || (g->n_eo && (!g->eo_cam_i || !g->eo_cam_j || !g->eo_meas || !g->eo_info))) {
        c->err = "csb_ba_set_graph: null array with a non-zero count";
        return CSB_ERR_INVALID;
    }
    if (!in_range(g->ec_cam, g->n_ec, g->n_cam) || !in_range(g->ec_cube, g->n_ec, g->n_cube) || !in_range(g->ep_cam, g->n_ep, g->n_cam) ||
        !in_range(g->ep_cube, g->n_ep, g->n_cube) || !in_range(g->eo_cam_i, g->n_eo, g->n_cam) || !in_range(g->eo_cam_j, g->n_eo, g->n_cam)) {
        c->err = "csb_ba_set_graph: edge vertex index out of range";
        return CSB_ERR_INVALID;
    }
    CSB_CUDA(c, cudaSetDevice(c->device));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    BAState& s = c->ba;
    ba_solver_release(s);
    HostGraph& h = s.host;
    h.cam_fixed.assign(g->cam_fixed, g->cam_fixed + g->n_cam); h.cube_fixed.assign(g->cube_fixed, g->cube_fixed + g->n_cube);
    h.ec_cam.assign(g->ec_cam, g->ec_cam + g->n_ec); h.ec_cube.assign(g->ec_cube, g->ec_cube + g->n_ec);
    h.ep_cam.assign(g->ep_cam, g->ep_cam + g->n_ep); h.ep_cube.assign(g->ep_cube, g->ep_cube + g->n_ep);
    h.eo_i.assign(g->eo_cam_i, g->eo_cam_i + g->n_eo); h.eo_j.assign(g->eo_cam_j, g->eo_cam_j + g->n_eo);
    s.has_graph = false; s.has_estimates = false;
    CSB_TRY(sync_device(c, Counts{0, 0, 0, 0, 0}, NewEdgeData{g->ec_meas, g->ec_info, g->ep_meas, g->ep_info, g->ep_K, g->eo_meas, g->eo_info}));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    s.has_graph = true;
    return CSB_OK;
}

int csb_ba_add_frame(csb_context* c, const csb_ba_frame* f, int32_t* cam_index_out) {
    if (!c || !f) return CSB_ERR_INVALID;
    BAState& s = c->ba;
    if (!s.has_graph) { c->err = "csb_ba_add_frame before csb_ba_set_graph (an empty graph is a valid start)"; return CSB_ERR_STATE; }
    if (!f->cam7 || f->n_new_cubes < 0 || f->n_ec < 0 || f->n_eo < 0) { c->err = "csb_ba_add_frame: bad frame"; return CSB_ERR_INVALID; }
    if ((f->n_new_cubes && (!f->new_cubes10 || !f->new_cube_fixed)) || (f->n_ec && (!f->ec_cube || !f->ec_meas || !f->ec_info)) ||
        (f->n_eo && (!f->eo_cam_i || !f->eo_meas || !f->eo_info))) { c->err = "csb_ba_add_frame: null array with a non-zero count"; return CSB_ERR_INVALID; }
    if (s.n_cam + s.n_cube > 0 && !s.has_estimates) { c->err = "csb_ba_add_frame: the existing vertices have no estimates yet"; return CSB_ERR_STATE; }
    const Counts old{s.n_cam, s.n_cube, s.n_ec, s.n_ep, s.n_eo};
    const int cam = old.n_cam, n_cube_new = old.n_cube + f->n_new_cubes;
    if (!in_range(f->ec_cube, f->n_ec, n_cube_new) || !in_range(f->eo_cam_i, f->n_eo, cam)) { c->err = "csb_ba_add_frame: edge vertex index out of range"; return CSB_ERR_INVALID; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    HostGraph& h = s.host;
    h.cam_fixed.push_back(f->cam_fixed ? 1 : 0);
    for (int i = 0; i < f->n_new_cubes; i++) h.cube_fixed.push_back(f->new_cube_fixed[i] ? 1 : 0);
    for (int e = 0; e < f->n_ec; e++) { h.ec_cam.push_back(cam); h.ec_cube.push_back(f->ec_cube[e]); }
    for (int e = 0; e < f->n_eo; e++) { h.eo_i.push_back(f->eo_cam_i[e]); h.eo_j.push_back(cam); }
    CSB_TRY(sync_device(c, old, NewEdgeData{f->ec_meas, f->ec_info, nullptr, nullptr, nullptr, f->eo_meas, f->eo_info}));
    // estimates of the new vertices behind the (possibly optimised) ones the device holds
    CSB_CUDA(c, cudaMemcpyAsync(const_cast<double*>(s.B.cams7) + 7 * (size_t)cam, f->cam7, 56, cudaMemcpyHostToDevice, c->stream));
    if (f->n_new_cubes)
        CSB_CUDA(c, cudaMemcpyAsync(const_cast<double*>(s.B.cubes10) + 10 * (size_t)old.n_cube, f->new_cubes10, 80 * (size_t)f->n_new_cubes, cudaMemcpyHostToDevice, c->stream));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    s.has_estimates = true;
    if (cam_index_out) *cam_index_out = cam;
    return CSB_OK;
}

int csb_ba_upload_estimates(csb_context* c, const double* cams7, const double* cubes10) {
    if (!c) return CSB_ERR_INVALID;
    BAState& s = c->ba;
    if (!s.has_graph) { c->err = "csb_ba_upload_estimates before csb_ba_set_graph"; return CSB_ERR_STATE; }
    if ((!cams7 && s.n_cam) || (!cubes10 && s.n_cube)) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    if (s.n_cam) CSB_CUDA(c, cudaMemcpyAsync(const_cast<double*>(s.B.cams7), cams7, 56 * (size_t)s.n_cam, cudaMemcpyHostToDevice, c->stream));
    if (s.n_cube) CSB_CUDA(c, cudaMemcpyAsync(const_cast<double*>(s.B.cubes10), cubes10, 80 * (size_t)s.n_cube, cudaMemcpyHostToDevice, c->stream));
    s.has_estimates = true;
    return CSB_OK;
}

static int ba_run_impl(csb_context* c, bool want_J) {
    BAState& s = c->ba;
    if (!s.has_graph || !s.has_estimates) { c->err = "csb_ba_run before graph/estimates"; return CSB_ERR_STATE; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    CSB_CUDA(c, ba_launch(s.B, want_J, c->stream, &s.launches_last, s.analytic));
    s.ran = true;
    s.have_J = want_J;
    return CSB_OK;
}

int csb_ba_set_jacobian_mode(csb_context* c, int mode) {
    if (!c || (mode != CSB_BA_JACOBIAN_NUMERIC && mode != CSB_BA_JACOBIAN_ANALYTIC)) return CSB_ERR_INVALID;
    c->ba.analytic = mode == CSB_BA_JACOBIAN_ANALYTIC;
    return CSB_OK;
}

int csb_ba_run(csb_context* c) {
    if (!c) return CSB_ERR_INVALID;
    return ba_run_impl(c, false);
}

int csb_ba_download(csb_context* c, const csb_ba_output* o) {
    if (!c || !o) return CSB_ERR_INVALID;
    BAState& s = c->ba;
    if (!s.ran) { c->err = "csb_ba_download before csb_ba_run"; return CSB_ERR_STATE; }
    if (!s.have_J && (o->ec_Ji || o->ec_Jj || o->ep_Ji || o->ep_Jj || o->eo_Ji || o->eo_Jj)) {
        c->err = "csb_ba_download: Jacobians requested, but the last run did not materialise them (use csb_ba_linearize with Jacobian pointers)";
        return CSB_ERR_STATE;
    }
    CSB_CUDA(c, cudaSetDevice(c->device));
    const BABuffers& B = s.B;
    auto dl = [&](double* dst, const double* src, size_t n) -> int {
        if (dst && n) CSB_CUDA(c, cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToHost, c->stream));
        return CSB_OK;
    };
    CSB_TRY(dl(o->ec_err, B.ec_err, (size_t)s.n_ec * 9)); CSB_TRY(dl(o->ep_err, B.ep_err, (size_t)s.n_ep * 4)); CSB_TRY(dl(o->eo_err, B.eo_err, (size_t)s.n_eo * 6));
    CSB_TRY(dl(o->ec_Ji, B.ec_Ji, (size_t)s.n_ec * 54)); CSB_TRY(dl(o->ec_Jj, B.ec_Jj, (size_t)s.n_ec * 81));
    CSB_TRY(dl(o->ep_Ji, B.ep_Ji, (size_t)s.n_ep * 24)); CSB_TRY(dl(o->ep_Jj, B.ep_Jj, (size_t)s.n_ep * 36));
    CSB_TRY(dl(o->eo_Ji, B.eo_Ji, (size_t)s.n_eo * 36)); CSB_TRY(dl(o->eo_Jj, B.eo_Jj, (size_t)s.n_eo * 36));
    CSB_TRY(dl(o->H_cam, B.H_cam, (size_t)s.n_cam * 36)); CSB_TRY(dl(o->b_cam, B.b_cam, (size_t)s.n_cam * 6));
    CSB_TRY(dl(o->H_cube, B.H_cube, (size_t)s.n_cube * 81)); CSB_TRY(dl(o->b_cube, B.b_cube, (size_t)s.n_cube * 9));
    CSB_TRY(dl(o->ec_Hij, B.ec_Hij, (size_t)s.n_ec * 54)); CSB_TRY(dl(o->ep_Hij, B.ep_Hij, (size_t)s.n_ep * 54)); CSB_TRY(dl(o->eo_Hij, B.eo_Hij, (size_t)s.n_eo * 36));
    CSB_TRY(dl(o->chi2, B.chi2, 1));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    return CSB_OK;
}

int csb_ba_linearize(csb_context* c, const double* cams7, const double* cubes10, const csb_ba_output* out) {
    if (!c || !out) return CSB_ERR_INVALID;
    CSB_TRY(csb_ba_upload_estimates(c, cams7, cubes10));
    bool want_J = out->ec_Ji || out->ec_Jj || out->ep_Ji || out->ep_Jj || out->eo_Ji || out->eo_Jj;
    CSB_TRY(ba_run_impl(c, want_J));
    return csb_ba_download(c, out);
}

}  // extern "C"
