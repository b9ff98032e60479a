// ba_solve.cu -- Levenberg-Marquardt on the device for the camera-cuboid graph.
//
// What the reference does per outer iteration (Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189,
// block_solver.hpp:353-486, linear_solver_dense.h:65-113): computeActiveErrors, buildSystem, then up to 10 trials of
// "(H + lambda I) x = b -> oplus -> computeActiveErrors -> rho test".  Here the whole trial runs on the device:
//
//   k_cube_inv     (H_ll + lambda I)^-1 of every free cuboid block and t_l = that * b_l            (one thread per cuboid)
//   k_edge_w       W_e = H_pl(e) (H_ll + lambda I)^-1 for every camera-cuboid edge                  (54 threads per edge)
//   k_schur_blocks S_IJ = [H_pp + lambda I]_IJ - sum_e1,e2 W_e1 H_pl(e2)^T (+ odometry H_ij), lower block triangle,
//                  one warp per 6x6 block, contributions gathered in a fixed order (no atomics)
//   k_schur_rhs    r_I = b_I - sum_e W_e b_l(e)
//   k_chol_solve   blocked right-looking Cholesky of S (32x32 tiles) + L y = r, L^T x = y: ONE cooperative kernel with grid barriers
//   k_cube_back    x_l = (H_ll + lambda I)^-1 (b_l - sum_e H_pl(e)^T x_cam(e))
//   k_apply        trial estimates = oplus(estimates, x): cameras exp(dx) * T, cuboids pose * exp(dx), scale + ds
//   k_edge_chi2    chi2 of every edge at the trial estimates (residual only), reduced in a fixed order
//   k_scale        x^T (lambda x + b)
// The reference solves the full (cuboids + cameras) dense system with Eigen's LDLT; eliminating the block-diagonal cuboid part
// first is the same linear system, so the iterates agree to round-off.  The accept / reject logic (rho, lambda, ni) is scalar
// and stays on the host, reading three doubles per trial.
//
// These kernels are not on the bit-exact path (the reference's own solve is only reproducible to round-off), so they use fma().
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "ba.h"
#include "ba_dev.cuh"
#include "context.h"

namespace csb {

struct SolveView {
    int n_fc, n_fl;       // free cameras / free cuboids
    int n, ld;            // Schur system size (6 * n_fc) and leading dimension (multiple of 32)
    int n_pl;             // camera-cuboid edges with both ends free (ec first, then ep)
    const int *cam_col, *cube_col;      // vertex -> free index or -1
    const int *fc_cam, *fl_cube;        // free index -> vertex
    const int *pl_cam, *pl_cube;        // per pl edge: free camera / free cuboid index
    const int *pl_src;                  // per pl edge: index into ec_Hij (>= 0) or -1 - index into ep_Hij
    const int *blk_ptr, *blk_I, *blk_J; // S block list (lower block triangle incl. diagonal)
    const int2* blk_ent;                // entries: (e1, e2) Schur pair | (-1, eo) odometry block | (-2, eo) transposed
    int n_blk;
    const int *cam_pl_ptr, *cam_pl_e;   // pl edges of every free camera, edge order
    const int *cube_pl_ptr, *cube_pl_e; // pl edges of every free cuboid, edge order
    double *Ainv, *tl, *W, *S, *rhs, *x_cam, *x_cube, *scal;  // scal: [0] trial chi2, [1] scale, [2] max diag, [3] cholesky ok (1/0), [4] lambda of the trial
    double *trial_cams7, *trial_cubes10, *edge_chi2;
    double* linv;      // inverses of the diagonal Cholesky tiles (ld / 32 tiles of 32 x 32)
    unsigned* bar;     // grid barrier of k_chol_solve: [0] arrivals, [1] generation
};

__device__ __forceinline__ const double* pl_hij(const BABuffers& B, const SolveView& V, int e) {
    const int s = V.pl_src[e];
    return s >= 0 ? B.ec_Hij + 54 * (size_t)s : B.ep_Hij + 54 * (size_t)(-1 - s);
}

// ---- (H_ll + lambda I)^-1 by Cholesky, one thread per free cuboid ---------------------------------------------------
__global__ void k_cube_inv(BABuffers B, SolveView V) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= V.n_fl) return;
    const double lambda = V.scal[4];
    const int v = V.fl_cube[l];
    double A[81], Li[81];
    for (int i = 0; i < 81; i++) A[i] = B.H_cube[81 * (size_t)v + i];  // symmetric: storage order irrelevant
    for (int i = 0; i < 9; i++) A[i * 9 + i] += lambda;
    // Cholesky A = L L^T (lower, row-major in A)
    bool ok = true;
    for (int j = 0; j < 9; j++) {
        double s = A[j * 9 + j];
        for (int k = 0; k < j; k++) s = fma(-A[j * 9 + k], A[j * 9 + k], s);
        if (!(s > 0)) { ok = false; s = 1; }
        const double d = sqrt(s);
        A[j * 9 + j] = d;
        for (int i = j + 1; i < 9; i++) {
            double t = A[i * 9 + j];
            for (int k = 0; k < j; k++) t = fma(-A[i * 9 + k], A[j * 9 + k], t);
            A[i * 9 + j] = t / d;
        }
    }
    // Li = L^-1 (lower), then Ainv = Li^T Li
    for (int c = 0; c < 9; c++)
        for (int r = 0; r < 9; r++) {
            if (r < c) { Li[r * 9 + c] = 0; continue; }
            double t = (r == c) ? 1.0 : 0.0;
            for (int k = c; k < r; k++) t = fma(-A[r * 9 + k], Li[k * 9 + c], t);
            Li[r * 9 + c] = t / A[r * 9 + r];
        }
    double* out = V.Ainv + 81 * (size_t)l;
    for (int r = 0; r < 9; r++)
        for (int c = 0; c < 9; c++) {
            double t = 0;
            for (int k = (r > c ? r : c); k < 9; k++) t = fma(Li[k * 9 + r], Li[k * 9 + c], t);
            out[r * 9 + c] = ok ? t : 0.0;
        }
    for (int r = 0; r < 9; r++) {
        double t = 0;
        for (int c = 0; c < 9; c++) t = fma(out[r * 9 + c], B.b_cube[9 * (size_t)v + c], t);
        V.tl[9 * (size_t)l + r] = t;
    }
    if (!ok) V.scal[3] = 0.0;
}

// W_e (6x9, column-major like H_ij) = H_pl(e) * Ainv(l(e))
__global__ void k_edge_w(BABuffers B, SolveView V) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = t / 54, idx = t - e * 54;
    if (e >= V.n_pl) return;
    const int c = idx / 6, r = idx - c * 6;
    const double* H = pl_hij(B, V, e);
    const double* Ai = V.Ainv + 81 * (size_t)V.pl_cube[e];
    double s = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) s = fma(H[k * 6 + r], Ai[k * 9 + c], s);
    V.W[54 * (size_t)e + idx] = s;
}

// one warp per listed 6x6 block of the lower block triangle of S (row-major, leading dimension ld)
__global__ void __launch_bounds__(128) k_schur_blocks(BABuffers B, SolveView V) {
    const int blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (blk >= V.n_blk) return;
    const double lambda = V.scal[4];
    const int I = V.blk_I[blk], J = V.blk_J[blk];
    for (int el = lane; el < 36; el += 32) {
        const int r = el / 6, c = el - r * 6;
        double acc = 0;
        if (I == J) {
            acc = B.H_cam[36 * (size_t)V.fc_cam[I] + c * 6 + r];
            if (r == c) acc += lambda;
        }
        for (int q = V.blk_ptr[blk]; q < V.blk_ptr[blk + 1]; q++) {
            const int2 en = V.blk_ent[q];
            if (en.x >= 0) {
                const double* W = V.W + 54 * (size_t)en.x;
                const double* H = pl_hij(B, V, en.y);
#pragma unroll
                for (int k = 0; k < 9; k++) acc = fma(-W[k * 6 + r], H[k * 6 + c], acc);
            } else if (en.x == -1) acc += B.eo_Hij[36 * (size_t)en.y + c * 6 + r];   // rows = vertex i of the edge = block row
            else acc += B.eo_Hij[36 * (size_t)en.y + r * 6 + c];                       // transposed
        }
        V.S[(size_t)(6 * I + r) * V.ld + 6 * J + c] = acc;
    }
}

// r_I = b_I - sum_e W_e b_l(e); the threads past the system fill the padding rows / columns of S (identity on the diagonal, zero
// right-hand side) so that the factorisation runs over full tiles.  Launched over ld threads, after S has been cleared.
__global__ void k_schur_rhs(BABuffers B, SolveView V) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = t / 6, r = t - I * 6;
    if (t >= V.n) {
        if (t < V.ld) { V.S[(size_t)t * V.ld + t] = 1.0; V.rhs[t] = 0.0; }  // (row ld of S is already zero there)
        return;
    }
    double acc = B.b_cam[6 * (size_t)V.fc_cam[I] + r];
    for (int q = V.cam_pl_ptr[I]; q < V.cam_pl_ptr[I + 1]; q++) {
        const int e = V.cam_pl_e[q];
        const double* W = V.W + 54 * (size_t)e;
        const double* bl = B.b_cube + 9 * (size_t)V.fl_cube[V.pl_cube[e]];
#pragma unroll
        for (int k = 0; k < 9; k++) acc = fma(-W[k * 6 + r], bl[k], acc);
    }
    V.rhs[6 * I + r] = acc;
    V.S[(size_t)V.ld * V.ld + 6 * I + r] = acc;  // the right-hand side rides through the factorisation as row ld of S (k_chol_solve)
}

// ---- blocked Cholesky, lower, in place, row-major, 32x32 tiles ------------------------------------------------------
constexpr int NB = 32;

// ---- the whole dense solve S x = r in ONE cooperative kernel --------------------------------------------------------------------
// Blocked right-looking Cholesky (32 x 32 tiles, lower, in place) followed by the two triangular solves, with grid-wide barriers
// between the steps instead of one kernel launch per step (config #4: 38 tile columns = 114 launches of the k_chol_* kernels above
// plus a serial k_trsv became one launch).  Per tile column k:
//   (1) CTA 0: L_kk = chol(S_kk) in shared memory, and its inverse (a 32 x 32 triangular inverse, one column per lane) -> `linv`
//   (2) every CTA, rows below the tile: L_ik = S_ik L_kk^-T as a product with the explicit inverse (no dependent chain per row)
//   (3) every CTA, tiles (i, j) of the trailing lower triangle: S_ij -= L_ik L_jk^T
// Then L y = r (tile by tile: y_k = L_kk^-1 r_k by CTA 0 with the stored inverses, then every CTA subtracts L_ik y_k from its rows)
// and L^T x = y backwards.  `linv` keeps the inverse of every diagonal tile (tiles x 32 x 32).
// The grid barrier is a generation counter in global memory: the kernel is launched with cudaLaunchCooperativeKernel, so all CTAs
// are resident.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned n_ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        volatile unsigned* gen = bar + 1;
        const unsigned g = *gen;
        if (atomicAdd(bar, 1u) == n_ctas - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

constexpr int CF_THREADS = 256;
#ifdef CSB_CHOL_DEBUG
#define CHOL_T(i) do { if (cta == 0 && tid == 0) { const long long n__ = clock64(); tacc[i] += n__ - tprev; tprev = n__; } } while (0)
#else
#define CHOL_T(i) do { } while (0)
#endif

// inverse of the lower-triangular tile in shared memory `a` into `b` (row r, column c): one warp, lane = column c of the inverse, forward
// substitution on e_c with the solution in registers (every index static) and the tile read by broadcast
__device__ __forceinline__ void tri_inverse_tile(double (*a)[NB + 1], double (*b)[NB + 1], const double* dinv, int lane) {
    double x[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) {
        double t0 = (r == lane) ? 1.0 : 0.0, t1 = 0.0;
#pragma unroll
        for (int q = 0; q + 1 < r; q += 2) { t0 = fma(-a[r][q], x[q], t0); t1 = fma(-a[r][q + 1], x[q + 1], t1); }
        if (r & 1) t0 = fma(-a[r][r - 1], x[r - 1], t0);
        x[r] = (r >= lane) ? (t0 + t1) * dinv[r] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < NB; r++) b[r][lane] = x[r];
}

__global__ void __launch_bounds__(CF_THREADS) k_chol_solve(double* S, int ld, double* v, double* linv, double* ok_flag, unsigned* bar) {
    __shared__ double a[NB][NB + 1], b[NB][NB + 1];
    __shared__ double xs[NB], dinv[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = ld / NB;
    const unsigned G = gridDim.x;
    const int cta = blockIdx.x;
#ifdef CSB_CHOL_DEBUG
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    // CTA 0: factor the diagonal tile k0 (already updated) and store L_kk and its inverse
    // CTA 0: L_kk = chol(S_kk) of the (already updated) diagonal tile in shared memory, all threads on the trailing part of every column,
    // then its inverse; both go back to global memory
    auto factor_diag = [&](int k) {
        const int k0 = k * NB;
        double* Lk = linv + (size_t)k * NB * NB;
        __syncthreads();  // the CTA's update of this tile is complete
        CHOL_T(3);
        for (int i = tid; i < NB * NB; i += CF_THREADS) a[i / NB][i % NB] = S[(size_t)(k0 + i / NB) * ld + k0 + i % NB];
        __syncthreads();
        CHOL_T(5);
        // Two-level blocking of the 32 x 32 tile: panels of 8 columns.  Inside a panel one warp (lane = row) factors column after column
        // -- pivot by rsqrt, the at most seven panel columns to its right updated from registers, __syncwarp only -- then all 256 threads
        // apply the panel to the rest of the tile as one rank-8 update (two block barriers per panel instead of two per column).
        {
            bool ok = true;
            constexpr int PW = 8;
            for (int p0 = 0; p0 < NB; p0 += PW) {
                if (warp == 0) {
                    for (int j = p0; j < p0 + PW; j++) {
                        double d = a[j][j];
                        if (!(d > 0)) { ok = false; d = 1.0; }
                        const double inv = rsqrt(d);
                        const double lj = (lane == j) ? d * inv : a[lane][j] * inv;  // column j of L (rows >= j)
                        double v[PW - 1], l[PW - 1];
#pragma unroll
                        for (int u = 0; u < PW - 1; u++) {
                            const int cc = j + 1 + u;
                            const bool in = cc < p0 + PW && cc <= lane;
                            v[u] = in ? a[lane][cc] : 0.0;
                            l[u] = in ? a[cc][j] : 0.0;
                        }
                        __syncwarp();
                        if (lane >= j) a[lane][j] = lj;
                        if (lane == j) dinv[j] = inv;
#pragma unroll
                        for (int u = 0; u < PW - 1; u++) {
                            const int cc = j + 1 + u;
                            // l[u] = a[cc][j] was read BEFORE it was scaled by 1 / sqrt(d_jj): scale it here (cc > j)
                            if (cc < p0 + PW && cc <= lane) a[lane][cc] = fma(-lj, l[u] * inv, v[u]);
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                const int r0 = p0 + PW, rem = NB - r0;
                for (int idx = tid; idx < rem * rem; idx += CF_THREADS) {
                    const int r = r0 + idx / rem, c = r0 + idx % rem;
                    if (c <= r) {
                        double acc = 0;
#pragma unroll
                        for (int q = 0; q < PW; q++) acc = fma(a[r][p0 + q], a[c][p0 + q], acc);
                        a[r][c] -= acc;
                    }
                }
                __syncthreads();
            }
            if (!ok && tid == 0) *ok_flag = 0.0;
        }
        if (warp == 0) {
            CHOL_T(6);
            tri_inverse_tile(a, b, dinv, lane);
            CHOL_T(7);
        }
        __syncthreads();
        for (int i = tid; i < NB * NB; i += CF_THREADS) {
            const int r = i / NB, c = i % NB;
            S[(size_t)(k0 + r) * ld + k0 + c] = (c <= r) ? a[r][c] : 0.0;
            Lk[i] = b[r][c];
        }
        __threadfence();
    };
    if (cta == 0) factor_diag(0);
    CHOL_T(0);
    grid_barrier(bar, G);
    CHOL_T(1);
    for (int k = 0; k < tiles - 1; k++) {
        const int k0 = k * NB;
        const double* Lk = linv + (size_t)k * NB * NB;
        // (2) L_ik = S_ik Linv^T : out[r][c] = sum_q S_ik[r][q] Linv[c][q]   (Linv lower: q <= c); a warp per row, lane = output column
        {
            __syncthreads();
            for (int i = tid; i < NB * NB; i += CF_THREADS) b[i / NB][i % NB] = Lk[i];
            __syncthreads();
            const int n_rows = ld - k0 - NB + 1;  // + the right-hand side, row `ld` of S: L y = r comes out of the factorisation
            for (int r = cta * (CF_THREADS / 32) + warp; r < n_rows; r += G * (CF_THREADS / 32)) {
                double* p = S + (size_t)(k0 + NB + r) * ld + k0;
                const double mine = p[lane];
                double acc = 0;
#pragma unroll
                for (int q = 0; q < NB; q++) {
                    const double sv = __shfl_sync(0xffffffffu, mine, q);
                    acc = fma(sv, (q <= lane) ? b[lane][q] : 0.0, acc);
                }
                p[lane] = acc;
            }
        }
        CHOL_T(2);
        grid_barrier(bar, G);
        CHOL_T(1);
        // (3) trailing update, tiles (ti >= tj) behind the panel, strided over the CTAs; CTA 0 takes tile (0, 0) first -- the next diagonal
        // tile -- and factors it at once, while the other CTAs are still updating
        {
            const int tt = tiles - k - 1;
            const int n_t = tt * (tt + 1) / 2;
            const int tx = tid & 15, ty = tid >> 4;
            // the right-hand-side row against tile column tj: one warp each, taken from the far end of the CTA list (CTA 0 has the diagonal tile)
            for (int tj = (int)(G - 1 - cta) * (CF_THREADS / 32) + warp; tj < tt; tj += (int)G * (CF_THREADS / 32)) {
                const int c0 = k0 + NB + tj * NB;
                const double yv = S[(size_t)ld * ld + k0 + lane];
                const double* Lj = S + (size_t)(c0 + lane) * ld + k0;  // lane = column of the tile = row c0 + lane of L
                double acc = 0;
#pragma unroll
                for (int q = 0; q < NB; q++) acc = fma(__shfl_sync(0xffffffffu, yv, q), Lj[q], acc);
                S[(size_t)ld * ld + c0 + lane] -= acc;
            }
            for (int t = cta; t < n_t; t += G) {
                int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
                while (ti * (ti + 1) / 2 > t) ti--;
                const int tj = t - ti * (ti + 1) / 2;
                const int r0 = k0 + NB + ti * NB, c0 = k0 + NB + tj * NB;
                __syncthreads();
                for (int i = tid; i < NB * NB; i += CF_THREADS) {
                    a[i / NB][i % NB] = S[(size_t)(r0 + i / NB) * ld + k0 + i % NB];
                    b[i / NB][i % NB] = S[(size_t)(c0 + i / NB) * ld + k0 + i % NB];
                }
                __syncthreads();
                double acc[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
                for (int q = 0; q < NB; q++) {
                    const double a0 = a[ty][q], a1 = a[ty + 16][q], b0 = b[tx][q], b1 = b[tx + 16][q];
                    acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
                    acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
                }
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j2 = 0; j2 < 2; j2++) {
                        const int r = r0 + ty + 16 * i, c = c0 + tx + 16 * j2;
                        if (c <= r) S[(size_t)r * ld + c] -= acc[i][j2];
                    }
                if (t == 0) { CHOL_T(3); factor_diag(k + 1); CHOL_T(0); }  // (cta == 0)
            }
        }
        CHOL_T(3);
        grid_barrier(bar, G);
        CHOL_T(1);
    }
    if (cta != 0) return;
    // ---- L y = r is row `ld` of the factorised matrix (the right-hand side went through the panel / update steps as one more row);
    // CTA 0 alone does the backward solve (0.7 M multiply-adds; the lower triangle streams through once)
    {
        // last tile column: only the right-hand-side row is left below it
        const int k0 = (tiles - 1) * NB;
        const double* Lk = linv + (size_t)(tiles - 1) * NB * NB;
        __syncthreads();
        if (warp == 0) {
            double* p = S + (size_t)ld * ld + k0;
            const double mine = p[lane];
            double acc = 0;
#pragma unroll
            for (int q = 0; q < NB; q++) acc = fma(__shfl_sync(0xffffffffu, mine, q), (q <= lane) ? Lk[lane * NB + q] : 0.0, acc);
            p[lane] = acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < ld; i += CF_THREADS) v[i] = S[(size_t)ld * ld + i];
    __syncthreads();
    CHOL_T(4);
    // L^T x = y
    for (int k = tiles - 1; k >= 0; k--) {
        const int k0 = k * NB;
        const double* Lk = linv + (size_t)k * NB * NB;
        __syncthreads();
        if (warp == 0) {
            const double yv = v[k0 + lane];  // x_k = Linv^T y_k: lane = row r, sum over q >= r of Linv[q][r] y[q]
            double acc = 0;
#pragma unroll
            for (int q = 0; q < NB; q++) acc = fma((q >= lane) ? Lk[q * NB + lane] : 0.0, __shfl_sync(0xffffffffu, yv, q), acc);
            v[k0 + lane] = acc;
            xs[lane] = acc;
        }
        __syncthreads();
        for (int row = tid; row < k0; row += CF_THREADS) {
            double t = v[row];
#pragma unroll
            for (int c = 0; c < NB; c++) t = fma(-S[(size_t)(k0 + c) * ld + row], xs[c], t);
            v[row] = t;
        }
    }
    CHOL_T(5);
#ifdef CSB_CHOL_DEBUG
    if (tid == 0) printf("k_chol_solve cycles: diag-rest %lld barrier %lld panel %lld update %lld fwd %lld bwd+tile-load %lld factor %lld inverse %lld\n", tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7]);
#endif
}

// x_l = Ainv (b_l - sum_e H_pl(e)^T x_cam(e)), one thread per (cuboid, row) after a per-cuboid gather
__global__ void __launch_bounds__(32) k_cube_back(BABuffers B, SolveView V) {
    const int l = blockIdx.x, lane = threadIdx.x;
    __shared__ double g[9];
    if (lane < 9) {
        double acc = B.b_cube[9 * (size_t)V.fl_cube[l] + lane];
        for (int q = V.cube_pl_ptr[l]; q < V.cube_pl_ptr[l + 1]; q++) {
            const int e = V.cube_pl_e[q];
            const double* H = pl_hij(B, V, e);
            const double* xc = V.rhs + 6 * (size_t)V.pl_cam[e];
#pragma unroll
            for (int r = 0; r < 6; r++) acc = fma(-H[lane * 6 + r], xc[r], acc);
        }
        g[lane] = acc;
    }
    __syncwarp();
    if (lane < 9) {
        const double* Ai = V.Ainv + 81 * (size_t)l;
        double t = 0;
#pragma unroll
        for (int c = 0; c < 9; c++) t = fma(Ai[lane * 9 + c], g[c], t);
        V.x_cube[9 * (size_t)l + lane] = t;
    }
}

__device__ __forceinline__ void se3_store7(const SE3& s, double* v) { v[0] = s.t.x; v[1] = s.t.y; v[2] = s.t.z; v[3] = s.r.x; v[4] = s.r.y; v[5] = s.r.z; v[6] = s.r.w; }

// trial = oplus(current, x); solved == 0 -> x = 0 (the reference's behaviour when the linear solve fails)
__global__ void k_apply(BABuffers B, SolveView V) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool solved = V.scal[3] != 0.0;
    if (t < B.n_cam) {
        SE3 T = se3_from_vec7(B.cams7 + 7 * (size_t)t);
        const int I = V.cam_col[t];
        if (I >= 0) {
            double u[6];
            for (int k = 0; k < 6; k++) u[k] = solved ? V.rhs[6 * I + k] : 0.0;
            T = se3_mul(se3_exp(u), T);  // VertexSE3Expmap::oplusImpl
        }
        se3_store7(T, V.trial_cams7 + 7 * (size_t)t);
    } else if (t < B.n_cam + B.n_cube) {
        const int v = t - B.n_cam;
        Cube c = cube_from_vec10(B.cubes10 + 10 * (size_t)v);
        const int l = V.cube_col[v];
        if (l >= 0) {
            double u[9];
            for (int k = 0; k < 9; k++) u[k] = solved ? V.x_cube[9 * (size_t)l + k] : 0.0;
            c = cube_exp_update(c, u);  // VertexCuboid::oplusImpl
        }
        double* o = V.trial_cubes10 + 10 * (size_t)v;
        se3_store7(c.pose, o);
        o[7] = c.scale.x; o[8] = c.scale.y; o[9] = c.scale.z;
    }
}

// chi2 of every edge at the given estimates (computeActiveErrors + Edge::chi2), one thread per edge
__global__ void __launch_bounds__(128) k_edge_chi2(BABuffers B, const double* cams7, const double* cubes10, double* out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_all = B.n_ec + B.n_ep + B.n_eo;
    if (t >= n_all) return;
    EdgeCtx x;
    double e[9], chi = 0;
    if (t < B.n_ec) {
        const int ee = t;
        x.meas_cube = cube_from_vec10(B.ec_meas + 10 * (size_t)ee);
        const SE3 cam = se3_from_vec7(cams7 + 7 * (size_t)B.ec_cam[ee]);
        const Cube cube = cube_from_vec10(cubes10 + 10 * (size_t)B.ec_cube[ee]);
        edge_error<EDGE_CUBOID>(x, cam, cube, cam, e);
        const double* info = B.ec_info + 81 * (size_t)ee;
        for (int r = 0; r < 9; r++) { double s = 0; for (int k = 0; k < 9; k++) s += info[r * 9 + k] * e[k]; chi += e[r] * s; }
    } else if (t < B.n_ec + B.n_ep) {
        const int ee = t - B.n_ec;
        for (int k = 0; k < 4; k++) x.meas4[k] = B.ep_meas[4 * (size_t)ee + k];
        x.K = B.ep_K + 9 * (size_t)ee;
        const SE3 cam = se3_from_vec7(cams7 + 7 * (size_t)B.ep_cam[ee]);
        const Cube cube = cube_from_vec10(cubes10 + 10 * (size_t)B.ep_cube[ee]);
        edge_error<EDGE_PROJ>(x, cam, cube, cam, e);
        const double* info = B.ep_info + 16 * (size_t)ee;
        for (int r = 0; r < 4; r++) { double s = 0; for (int k = 0; k < 4; k++) s += info[r * 4 + k] * e[k]; chi += e[r] * s; }
    } else {
        const int ee = t - B.n_ec - B.n_ep;
        x.meas_se3 = se3_from_vec7(B.eo_meas + 7 * (size_t)ee);
        const SE3 ci = se3_from_vec7(cams7 + 7 * (size_t)B.eo_i[ee]), cj = se3_from_vec7(cams7 + 7 * (size_t)B.eo_j[ee]);
        edge_error<EDGE_ODOM>(x, ci, Cube{}, cj, e);
        const double* info = B.eo_info + 36 * (size_t)ee;
        for (int r = 0; r < 6; r++) { double s = 0; for (int k = 0; k < 6; k++) s += info[r * 6 + k] * e[k]; chi += e[r] * s; }
    }
    out[t] = chi;
}

// fixed-order reductions by one warp: out[0] = sum(edge chi2), out[1] = x^T (lambda x + b) + 1e-3
__global__ void __launch_bounds__(32) k_trial_scalars(BABuffers B, SolveView V) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const double lambda = V.scal[4];
    double s = 0;
    const int n_all = B.n_ec + B.n_ep + B.n_eo;
    for (int i = lane; i < n_all; i += 32) s += V.edge_chi2[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(FULL, s, o);
    double q = 0;
    const bool solved = V.scal[3] != 0.0;
    if (solved) {
        for (int i = lane; i < 9 * V.n_fl; i += 32) { const double x = V.x_cube[i]; q += x * (lambda * x + B.b_cube[9 * (size_t)V.fl_cube[i / 9] + i % 9]); }
        for (int i = lane; i < 6 * V.n_fc; i += 32) { const double x = V.rhs[i]; q += x * (lambda * x + B.b_cam[6 * (size_t)V.fc_cam[i / 6] + i % 6]); }
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_down_sync(FULL, q, o);
    if (lane == 0) { V.scal[0] = s; V.scal[1] = q + 1e-3; }
}

__global__ void __launch_bounds__(32) k_max_diag(BABuffers B, SolveView V) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    double m = 0;
    for (int i = lane; i < 9 * V.n_fl; i += 32) m = fmax(m, fabs(B.H_cube[81 * (size_t)V.fl_cube[i / 9] + (i % 9) * 10]));
    for (int i = lane; i < 6 * V.n_fc; i += 32) m = fmax(m, fabs(B.H_cam[36 * (size_t)V.fc_cam[i / 6] + (i % 6) * 7]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(FULL, m, o));
    if (lane == 0) V.scal[2] = m;
}


// ---- host side ------------------------------------------------------------------------------------------------------
struct SolveState {
    bool built = false;
    int chol_ctas = 1;  // grid of the cooperative k_chol_solve: one CTA per SM, all resident
    SolveView V{};
    std::vector<void*> allocs;
    // one LM trial / one linearisation as CUDA graphs (captured on first use; a failed capture leaves the direct launches)
    cudaGraphExec_t trial_graph = nullptr, lin_graph[2] = {nullptr, nullptr};
    bool trial_graph_failed = false, lin_graph_failed[2] = {false, false};
    int trial_nodes = 0, lin_nodes[2] = {0, 0};
    double* h_pin = nullptr;  // pinned: [0..1] = {cholesky-ok flag, lambda} going up, [2..6] = scalars coming back
};

}  // namespace csb

using namespace csb;

namespace {

#define CSB_TRY(x) do { int rc__ = (x); if (rc__ != CSB_OK) return rc__; } while (0)

template <class T>
int up(csb_context* c, std::vector<void*>& allocs, const T** out, const std::vector<T>& h) {
    void* p = nullptr;
    CSB_CUDA(c, cudaMalloc(&p, std::max<size_t>(h.size(), 1) * sizeof(T)));
    allocs.push_back(p);
    if (!h.empty()) CSB_CUDA(c, cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T*>(p);
    return CSB_OK;
}
template <class T>
int al(csb_context* c, std::vector<void*>& allocs, T** out, size_t n) {
    void* p = nullptr;
    CSB_CUDA(c, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    allocs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return CSB_OK;
}

// Structure of the reduced camera system (what BlockSolver::buildStructure does for Hschur), built from the host copy of the graph.
int build_solver(csb_context* c) {
    BAState& s = c->ba;
    SolveState* st = new SolveState();
    s.solver = st;
    SolveView& V = st->V;
    const HostGraph& g = s.host;
    std::vector<int> cam_col(s.n_cam, -1), cube_col(s.n_cube, -1), fc_cam, fl_cube;
    for (int i = 0; i < s.n_cube; i++) if (!g.cube_fixed[i]) { cube_col[i] = (int)fl_cube.size(); fl_cube.push_back(i); }
    for (int i = 0; i < s.n_cam; i++) if (!g.cam_fixed[i]) { cam_col[i] = (int)fc_cam.size(); fc_cam.push_back(i); }
    V.n_fc = (int)fc_cam.size(); V.n_fl = (int)fl_cube.size();
    V.n = 6 * V.n_fc; V.ld = ((V.n + NB - 1) / NB) * NB;
    if (V.ld == 0) V.ld = NB;
    std::vector<int> pl_cam, pl_cube, pl_src;
    for (int e = 0; e < s.n_ec; e++) if (cam_col[g.ec_cam[e]] >= 0 && cube_col[g.ec_cube[e]] >= 0) { pl_cam.push_back(cam_col[g.ec_cam[e]]); pl_cube.push_back(cube_col[g.ec_cube[e]]); pl_src.push_back(e); }
    for (int e = 0; e < s.n_ep; e++) if (cam_col[g.ep_cam[e]] >= 0 && cube_col[g.ep_cube[e]] >= 0) { pl_cam.push_back(cam_col[g.ep_cam[e]]); pl_cube.push_back(cube_col[g.ep_cube[e]]); pl_src.push_back(-1 - e); }
    V.n_pl = (int)pl_cam.size();
    // adjacency of pl edges
    std::vector<std::vector<int>> cam_pl(V.n_fc), cube_pl(V.n_fl);
    for (int e = 0; e < V.n_pl; e++) { cam_pl[pl_cam[e]].push_back(e); cube_pl[pl_cube[e]].push_back(e); }
    auto csr = [](const std::vector<std::vector<int>>& a, std::vector<int>& ptr, std::vector<int>& idx) {
        ptr.assign(a.size() + 1, 0);
        for (size_t i = 0; i < a.size(); i++) { ptr[i + 1] = ptr[i] + (int)a[i].size(); idx.insert(idx.end(), a[i].begin(), a[i].end()); }
    };
    std::vector<int> cam_pl_ptr, cam_pl_e, cube_pl_ptr, cube_pl_e;
    csr(cam_pl, cam_pl_ptr, cam_pl_e); csr(cube_pl, cube_pl_ptr, cube_pl_e);
    // S block entries, keyed by (I, J) with I >= J; order inside a block: cuboid index, then e1, then e2; odometry edges last, in edge order
    const size_t nb = (size_t)V.n_fc;
    std::vector<std::vector<int2>> ent(nb * (nb + 1) / 2);
    auto key = [&](int I, int J) { return (size_t)I * (I + 1) / 2 + J; };
    for (int l = 0; l < V.n_fl; l++)
        for (int e1 : cube_pl[l])
            for (int e2 : cube_pl[l]) {
                const int I = pl_cam[e1], J = pl_cam[e2];
                if (I >= J) ent[key(I, J)].push_back(make_int2(e1, e2));
            }
    for (int e = 0; e < s.n_eo; e++) {
        const int ci = cam_col[g.eo_i[e]], cj = cam_col[g.eo_j[e]];
        if (ci < 0 || cj < 0) continue;
        if (ci == cj) continue;  // self edges do not occur
        if (ci > cj) ent[key(ci, cj)].push_back(make_int2(-1, e));
        else ent[key(cj, ci)].push_back(make_int2(-2, e));
    }
    std::vector<int> blk_ptr{0}, blk_I, blk_J;
    std::vector<int2> blk_ent;
    for (int I = 0; I < V.n_fc; I++)
        for (int J = 0; J <= I; J++) {
            const auto& v = ent[key(I, J)];
            if (v.empty() && I != J) continue;  // untouched off-diagonal blocks stay zero
            blk_I.push_back(I); blk_J.push_back(J);
            blk_ent.insert(blk_ent.end(), v.begin(), v.end());
            blk_ptr.push_back((int)blk_ent.size());
        }
    V.n_blk = (int)blk_I.size();
    auto& A = st->allocs;
    CSB_TRY(up(c, A, &V.cam_col, cam_col)); CSB_TRY(up(c, A, &V.cube_col, cube_col)); CSB_TRY(up(c, A, &V.fc_cam, fc_cam)); CSB_TRY(up(c, A, &V.fl_cube, fl_cube));
    CSB_TRY(up(c, A, &V.pl_cam, pl_cam)); CSB_TRY(up(c, A, &V.pl_cube, pl_cube)); CSB_TRY(up(c, A, &V.pl_src, pl_src));
    CSB_TRY(up(c, A, &V.blk_ptr, blk_ptr)); CSB_TRY(up(c, A, &V.blk_I, blk_I)); CSB_TRY(up(c, A, &V.blk_J, blk_J)); CSB_TRY(up(c, A, &V.blk_ent, blk_ent));
    CSB_TRY(up(c, A, &V.cam_pl_ptr, cam_pl_ptr)); CSB_TRY(up(c, A, &V.cam_pl_e, cam_pl_e)); CSB_TRY(up(c, A, &V.cube_pl_ptr, cube_pl_ptr)); CSB_TRY(up(c, A, &V.cube_pl_e, cube_pl_e));
    CSB_TRY(al(c, A, &V.Ainv, 81 * (size_t)V.n_fl)); CSB_TRY(al(c, A, &V.tl, 9 * (size_t)V.n_fl)); CSB_TRY(al(c, A, &V.W, 54 * (size_t)V.n_pl));
    CSB_TRY(al(c, A, &V.S, (size_t)(V.ld + 1) * V.ld)); /* + the right-hand side as row ld */ CSB_TRY(al(c, A, &V.rhs, (size_t)V.ld)); CSB_TRY(al(c, A, &V.x_cube, 9 * (size_t)V.n_fl));
    CSB_TRY(al(c, A, &V.scal, 8)); CSB_CUDA(c, cudaMallocHost(&st->h_pin, 8 * sizeof(double))); CSB_TRY(al(c, A, &V.trial_cams7, 7 * (size_t)s.n_cam)); CSB_TRY(al(c, A, &V.trial_cubes10, 10 * (size_t)s.n_cube));
    CSB_TRY(al(c, A, &V.edge_chi2, (size_t)(s.n_ec + s.n_ep + s.n_eo)));
    CSB_TRY(al(c, A, &V.linv, (size_t)V.ld * NB)); CSB_TRY(al(c, A, &V.bar, 4));
    CSB_CUDA(c, cudaMemset(V.bar, 0, 16));
    V.x_cam = V.rhs;
    {
        int per_sm = 0;
        CSB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chol_solve, CF_THREADS, 0));
        if (per_sm < 1) { c->err = "k_chol_solve does not fit an SM"; return CSB_ERR_CUDA; }
        st->chol_ctas = c->num_sms;
    }
    st->built = true;
    return CSB_OK;
}

}  // namespace

namespace csb {
void ba_solver_release(BAState& s) {
    SolveState* st = reinterpret_cast<SolveState*>(s.solver);
    if (!st) return;
    for (void* p : st->allocs) cudaFree(p);
    if (st->trial_graph) cudaGraphExecDestroy(st->trial_graph);
    for (int i = 0; i < 2; i++) if (st->lin_graph[i]) cudaGraphExecDestroy(st->lin_graph[i]);
    if (st->h_pin) cudaFreeHost(st->h_pin);
    delete st;
    s.solver = nullptr;
}
}  // namespace csb

namespace {

// the kernels of one LM trial, in stream order (10 kernels + 1 memset); returns the number of kernels
int launch_trial(const BABuffers& B, const SolveView& V, const SolveState* st, int n_cam, int n_cube, int n_edges, cudaStream_t sm, cudaError_t* err) {
    int n = 0;
    if (V.n_fl) { k_cube_inv<<<(V.n_fl + 31) / 32, 32, 0, sm>>>(B, V); n++; }
    if (V.n_pl) { k_edge_w<<<(V.n_pl * 54 + 127) / 128, 128, 0, sm>>>(B, V); n++; }
    *err = cudaMemsetAsync(V.S, 0, sizeof(double) * (size_t)(V.ld + 1) * V.ld, sm);
    if (*err != cudaSuccess) return n;
    if (V.n_blk) { k_schur_blocks<<<(V.n_blk * 32 + 127) / 128, 128, 0, sm>>>(B, V); n++; }
    k_schur_rhs<<<(V.ld + 127) / 128, 128, 0, sm>>>(B, V); n++;
    {
        double* a_S = V.S; int a_ld = V.ld; double* a_v = V.rhs; double* a_linv = V.linv; double* a_ok = V.scal + 3; unsigned* a_bar = V.bar;
        void* args[] = {&a_S, &a_ld, &a_v, &a_linv, &a_ok, &a_bar};
        *err = cudaLaunchCooperativeKernel((const void*)k_chol_solve, dim3(st->chol_ctas), dim3(CF_THREADS), args, 0, sm);
        if (*err != cudaSuccess) return n;
        n++;
    }
    if (V.n_fl) { k_cube_back<<<V.n_fl, 32, 0, sm>>>(B, V); n++; }
    k_apply<<<(n_cam + n_cube + 127) / 128, 128, 0, sm>>>(B, V); n++;
    if (n_edges) { k_edge_chi2<<<(n_edges + 127) / 128, 128, 0, sm>>>(B, V.trial_cams7, V.trial_cubes10, V.edge_chi2); n++; }
    k_trial_scalars<<<1, 32, 0, sm>>>(B, V); n++;
    *err = cudaGetLastError();
    return n;
}

// capture `body` (which issues work on `sm`) into an executable graph; on any failure the stream is left usable and *out stays NULL
template <class F>
bool capture_graph(cudaStream_t sm, cudaGraphExec_t* out, F body) {
    if (cudaStreamBeginCapture(sm, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    const bool ok_body = body();
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(sm, &g);
    if (e != cudaSuccess || !ok_body || !g) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return false; }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { cudaGetLastError(); return false; }
    *out = ex;
    return true;
}

}  // namespace

extern "C" int csb_ba_optimize(csb_context* c, int iterations, double* cams7_out, double* cubes10_out, csb_ba_optimize_stats* stats) {
    if (!c || iterations < 0) return CSB_ERR_INVALID;
    BAState& s = c->ba;
    if (!s.has_graph || !s.has_estimates) { c->err = "csb_ba_optimize before graph/estimates"; return CSB_ERR_STATE; }
    CSB_CUDA(c, cudaSetDevice(c->device));
    if (!s.solver) CSB_TRY(build_solver(c));
    SolveState* st = reinterpret_cast<SolveState*>(s.solver);
    SolveView& V = st->V;
    BABuffers& B = s.B;
    cudaStream_t sm = c->stream;
    const int n_edges = s.n_ec + s.n_ep + s.n_eo;
    // OptimizationAlgorithmLevenberg state (levenberg.cpp:45-58)
    const double tau = 1e-5, goodUp = 2. / 3., goodLow = 1. / 3.;
    const int maxTrials = 10;
    double lambda = -1, ni = 2;
    int nBad = 0, it_done = 0, trials_total = 0, kernels = 0, launches = 0;
    double chi_final = 0;
    double* hp = st->h_pin;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (stats) { CSB_CUDA(c, cudaEventCreate(&ev0)); CSB_CUDA(c, cudaEventCreate(&ev1)); CSB_CUDA(c, cudaEventRecord(ev0, sm)); }
    const int am = c->ba.analytic ? 1 : 0;
    for (int it = 0; it < iterations; it++) {
        // computeActiveErrors + buildSystem at the current estimates: one graph launch
        if (!st->lin_graph[am] && !st->lin_graph_failed[am]) {
            int nl = 0;
            const bool analytic = c->ba.analytic;
            if (!capture_graph(sm, &st->lin_graph[am], [&]() { return ba_launch(B, false, sm, &nl, analytic) == cudaSuccess; })) st->lin_graph_failed[am] = true;
            st->lin_nodes[am] = nl;
        }
        if (st->lin_graph[am]) { CSB_CUDA(c, cudaGraphLaunch(st->lin_graph[am], sm)); kernels += st->lin_nodes[am]; launches++; }
        else { int nl = 0; CSB_CUDA(c, ba_launch(B, false, sm, &nl, c->ba.analytic)); kernels += nl; launches += nl; }
        if (it == 0) { k_max_diag<<<1, 32, 0, sm>>>(B, V); kernels++; launches++; }
        CSB_CUDA(c, cudaMemcpyAsync(&hp[2], B.chi2, 8, cudaMemcpyDeviceToHost, sm));
        CSB_CUDA(c, cudaMemcpyAsync(&hp[3], V.scal + 2, 8, cudaMemcpyDeviceToHost, sm));
        CSB_CUDA(c, cudaStreamSynchronize(sm));
        double currentChi = hp[2], tempChi = currentChi;
        const double iniChi = currentChi;
        if (it == 0) { lambda = tau * hp[3]; ni = 2; nBad = 0; }
        double rho = 0;
        int qmax = 0;
        do {
            hp[0] = 1.0; hp[1] = lambda;  // cholesky-ok flag, lambda of this trial -> scal[3..4]
            CSB_CUDA(c, cudaMemcpyAsync(V.scal + 3, hp, 16, cudaMemcpyHostToDevice, sm));
            if (!st->trial_graph && !st->trial_graph_failed) {
                cudaError_t ce = cudaSuccess;
                int nk = 0;
                if (!capture_graph(sm, &st->trial_graph, [&]() { nk = launch_trial(B, V, st, s.n_cam, s.n_cube, n_edges, sm, &ce); return ce == cudaSuccess; })) st->trial_graph_failed = true;
                st->trial_nodes = nk;
            }
            if (st->trial_graph) { CSB_CUDA(c, cudaGraphLaunch(st->trial_graph, sm)); kernels += st->trial_nodes; launches++; }
            else {
                cudaError_t ce = cudaSuccess;
                const int nk = launch_trial(B, V, st, s.n_cam, s.n_cube, n_edges, sm, &ce);
                CSB_CUDA(c, ce);
                kernels += nk; launches += nk;
            }
            CSB_CUDA(c, cudaMemcpyAsync(&hp[4], V.scal, 32, cudaMemcpyDeviceToHost, sm));
            CSB_CUDA(c, cudaStreamSynchronize(sm));
            const double* h_scal = &hp[4];
            const bool ok2 = h_scal[3] != 0.0;
            tempChi = ok2 ? h_scal[0] : std::numeric_limits<double>::max();
            rho = (currentChi - tempChi) / h_scal[1];
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, goodUp);
                const double scaleFactor = std::max(goodLow, alpha);
                lambda *= scaleFactor; ni = 2; currentChi = tempChi;
                // accept: the trial estimates become the estimates (discardTop)
                CSB_CUDA(c, cudaMemcpyAsync(const_cast<double*>(B.cams7), V.trial_cams7, 56 * (size_t)s.n_cam, cudaMemcpyDeviceToDevice, sm));
                CSB_CUDA(c, cudaMemcpyAsync(const_cast<double*>(B.cubes10), V.trial_cubes10, 80 * (size_t)s.n_cube, cudaMemcpyDeviceToDevice, sm));
            } else {
                lambda *= ni; ni *= 2;  // reject: estimates untouched (pop)
            }
            qmax++; trials_total++;
        } while (rho < 0 && qmax < maxTrials);
        it_done++;
        chi_final = currentChi;
        if (qmax == maxTrials || rho == 0) break;  // Terminate
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) break;
    }
    if (cams7_out && s.n_cam) CSB_CUDA(c, cudaMemcpyAsync(cams7_out, B.cams7, 56 * (size_t)s.n_cam, cudaMemcpyDeviceToHost, sm));
    if (cubes10_out && s.n_cube) CSB_CUDA(c, cudaMemcpyAsync(cubes10_out, B.cubes10, 80 * (size_t)s.n_cube, cudaMemcpyDeviceToHost, sm));
    if (stats) CSB_CUDA(c, cudaEventRecord(ev1, sm));
    CSB_CUDA(c, cudaStreamSynchronize(sm));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->iterations = it_done; stats->trials = trials_total; stats->n_kernel_launches = kernels; stats->n_launches = launches;
        stats->chi2 = chi_final; stats->lambda = lambda; stats->schur_dim = V.n;
        cudaEventElapsedTime(&stats->gpu_ms, ev0, ev1);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
    return CSB_OK;
}
