// edlines.cu -- EDLines line detection on the GPU: the use_LSD = false branch of line_lbd_detect::detect_filter_lines
// (line_lbd/class/line_lbd_allclass.cpp:130-149, 200-235; what object_slam selects, main_obj.cpp:503-505).  SURVEY.md 8 "next" row f-2.
//
// The algorithm is stated in edlines_dev.cuh as one-thread-per-item functions (see there for the reference lines): that source is executed on
// the host by tests/test_edlines_emul.py, bit for bit against oracle/oracle_edlines.cpp.  The streaming stages use those functions directly;
// the two sequential stages have warp-cooperative production kernels here that make the same decisions and produce the same numbers
// (tests/test_edlines_gpu.py: segment lists and descriptors bit-identical to the oracle on hardware).
//
//   lbd_launch_grad  (lbd.cu)  blur 5x5 + Sobel -> {dx, dy}                                 HBM / issue bound, shared with the LBD descriptor
//   k_ed_pixel       thread per pixel: packed gradient + direction map                     HBM bound (4 B read, 2 B written per pixel)
//   k_ed_anchor      thread per 32 anchor candidates: one word of the column-major bitmap  L2 bound
//   k_ed_draw_warp   one warp per frame: smart routing with the edge bitmap in shared memory, edge chains
//                                                                                          latency bound (the walk is sequential by construction)
//   k_ed_fit_warp    one warp per chain: least-squares segments, extension 32 pixels per step, validation with one det_atan2 per lane
//   k_ed_emit        one thread per frame: ordered compaction, end-point order, length filter
//   csb_edlines_describe -> k_lbd_describe (lbd.cu): the warp-cooperative LBD kernel on the detector's own key-line fields
//
// k_ed_draw (one thread per frame) remains as the fallback for frames whose edge bitmap does not fit shared memory.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "context.h"
#include "edlines.h"
#include "edlines_dev.cuh"
#include "lbd.h"

namespace csb {

__global__ void __launch_bounds__(256) k_ed_pixel(EdBuffers B, EdDims d, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ed_pixel(B, d, i);
}

__global__ void __launch_bounds__(128) k_ed_anchor(EdBuffers B, EdDims d) {
    const int word = blockIdx.x * blockDim.x + threadIdx.x, frame = blockIdx.y;
    if (word >= d.anchor_words) return;
    const int n_cand = d.nxc * d.nyc;
    uint32_t bits = 0;
    for (int b = 0; b < 32; b++) {
        const int item = word * 32 + b;
        if (item < n_cand && ed_anchor(B, d, frame, item)) bits |= 1u << b;
    }
    B.anchors[(size_t)frame * d.anchor_words + word] = bits;
}

__global__ void __launch_bounds__(32) k_ed_draw(EdBuffers B, EdDims d) {
    if (threadIdx.x == 0) ed_draw(B, d, blockIdx.x);
}

// ---- smart routing, one WARP per frame (the production kernel; k_ed_draw above is the one-thread transcription the host emulation shares) -----
// The walk is sequential -- every step reads the packed gradient of the pixel and of three neighbours and the `edge` bit, all dependent on
// the previous step -- so its speed is the latency of those reads.  With the bitmap in global memory every step's read-modify-write
// invalidates its L1 line and the next step's test goes to L2; here the frame's edge bitmap lives in shared memory (the gradient map is
// read-only and stays L1-resident around the walk; a shared-memory window for it was measured slower: its reloads cost more than the L1
// misses they save).  The 32 lanes execute the routing redundantly (identical registers, broadcast reads), lane 0 writes the bitmap,
// walk pixels are staged in registers (lane n % 32 keeps pixel n) and leave as one coalesced store per 32 steps, and the chain assembly
// (first part reversed + second part without the anchor) is a coalesced copy by all lanes.  Same results as ed_draw().
struct EdWindow {
    const uint16_t* gd;   // the frame's packed gradient map (read-only path: consecutive steps stay inside the same few L1 lines)
    int W, H;
};
__device__ __forceinline__ void edw_ensure(EdWindow&, int, int, int) {}
__device__ __forceinline__ uint16_t edw_at(const EdWindow& w, int x, int y) { return __ldg(w.gd + (size_t)y * w.W + x); }

// one walk, warp-uniform; returns the pixels written to `out` (global, capacity cap).  Same decisions as ed_walk() (edlines_dev.cuh), with the
// dependent chain of a step cut to one level of loads: the three candidate pixels' packed gradients AND their edge bits are fetched
// together, and the chosen candidate's values become the next step's `cur` / `visited` (ed_walk reloads both after the move; marking the
// current pixel cannot change the bit of a different pixel, and the gradient map is read-only).
__device__ __forceinline__ int edw_walk(EdWindow& w, uint32_t* edge, unsigned x, unsigned y, int lastDirection, ushort2* out, int cap, EdWalkState& st, int lane) {
    const int W = w.W, H = w.H;
    const uint16_t* gd = w.gd;
    int n = 0;
    ushort2 mine = make_ushort2(0, 0);
    int idx = (int)(y * (unsigned)W + x);
    uint16_t cur = __ldg(gd + idx);
    bool visited = (edge[idx >> 5] >> (idx & 31)) & 1u;
    // The lanes run the walk redundantly, so they can do something useful on the side: whenever the walk has left the middle of the window
    // fetched last, lane l prefetches (to L1) the lines of row y - 16 + l left and right of x.  A vertical edge otherwise meets a new cache
    // line -- an L2 round trip -- at every step.
    int pfx = -100000, pfy = -100000;
    while (ed_g(cur) > 0 && !visited) {
        if (abs((int)x - pfx) > 24 || abs((int)y - pfy) > 6) {
            pfx = (int)x; pfy = (int)y;
            const int ry = min(max((int)y - 16 + lane, 0), H - 1);
            const uint16_t* row = gd + (size_t)ry * W;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(row + max((int)x - 40, 0)));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(row + min((int)x + 40, W - 1)));
        }
        __syncwarp();
        if (lane == 0) edge[idx >> 5] |= 1u << (idx & 31);
        if ((n & 31) == lane) mine = make_ushort2((unsigned short)x, (unsigned short)y);
        n++;
        if ((n & 31) == 0) { const int o = n - 32 + lane; if (o < cap) out[o] = mine; }
        __syncwarp();
        // direction of this step (:1716-1850 and its three repetitions)
        int dir;
        if (ed_horizontal(cur)) {
            if (lastDirection == ED_UP || lastDirection == ED_DOWN) dir = (x > st.lastX) ? ED_RIGHT : ED_LEFT;
            else dir = lastDirection;
        } else {
            if (lastDirection == ED_RIGHT || lastDirection == ED_LEFT) dir = (y > st.lastY) ? ED_DOWN : ED_UP;
            else dir = lastDirection;
        }
        st.lastX = x; st.lastY = y;
        int o1, o2, o3, mx1, my1, mx2, my2, mx3, my3;
        if (dir == ED_RIGHT || dir == ED_LEFT) {
            const int sx = (dir == ED_RIGHT) ? 1 : -1;
            if (x == (dir == ED_RIGHT ? (unsigned)W - 1 : 0u) || y == 0 || y == (unsigned)H - 1) break;
            o2 = sx; o1 = sx - W; o3 = sx + W;
            mx1 = sx; my1 = -1; mx2 = sx; my2 = 0; mx3 = sx; my3 = 1;
        } else {
            const int sy = (dir == ED_DOWN) ? 1 : -1;
            if (x == 0 || x == (unsigned)W - 1 || y == (dir == ED_DOWN ? (unsigned)H - 1 : 0u)) break;
            o2 = sy * W; o1 = o2 + 1; o3 = o2 - 1;
            mx1 = 1; my1 = sy; mx2 = 0; my2 = sy; mx3 = -1; my3 = sy;
        }
        const int i1 = idx + o1, i2 = idx + o2, i3 = idx + o3;
        const uint16_t r1 = __ldg(gd + i1), r2 = __ldg(gd + i2), r3 = __ldg(gd + i3);
        const uint32_t e1 = edge[i1 >> 5], e2 = edge[i2 >> 5], e3 = edge[i3 >> 5];
        // neighbours are compared through an unsigned-char cast in the reference (:1746-1748)
        const unsigned char g1 = (unsigned char)ed_g(r1), g2 = (unsigned char)ed_g(r2), g3 = (unsigned char)ed_g(r3);
        if (g1 >= g2 && g1 >= g3) { x += mx1; y += my1; idx = i1; cur = r1; visited = (e1 >> (i1 & 31)) & 1u; }
        else if (g3 >= g2 && g3 >= g1) { x += mx3; y += my3; idx = i3; cur = r3; visited = (e3 >> (i3 & 31)) & 1u; }
        else { x += mx2; y += my2; idx = i2; cur = r2; visited = (e2 >> (i2 & 31)) & 1u; }
        lastDirection = dir;
    }
    // the staged tail
    if (n & 31) { const int o = (n & ~31) + lane; if (lane < (n & 31) && o < cap) out[o] = mine; }
    __syncwarp();
    return n;
}

__global__ void __launch_bounds__(32) k_ed_draw_warp(EdBuffers B, EdDims d) {
    extern __shared__ __align__(16) unsigned char ed_smem[];
    uint32_t* edge = reinterpret_cast<uint32_t*>(ed_smem);
    const int frame = blockIdx.x, lane = threadIdx.x;
    const int W = d.w, H = d.h;
    const uint32_t* anc = B.anchors + (size_t)frame * d.anchor_words;
    ushort2* p1 = B.part1 + (size_t)frame * d.part_cap;
    ushort2* p2 = B.part2 + (size_t)frame * d.part_cap;
    ushort2* chain = B.chain_px + (size_t)frame * d.chain_cap;
    int* sid = B.chain_sid + (size_t)frame * (d.max_edges + 2);
    for (int i = lane; i < d.edge_words; i += 32) edge[i] = 0;
    EdWindow w{B.gd + (size_t)frame * W * H, W, H};
    __syncwarp();
    EdWalkState st{0u, 0u};
    int n_chains = 0, n_px = 0;
    long long kept1 = 0, kept2 = 0;
    bool overflow = false;
    unsigned long long n_anchor = 0;
    const int n_cand = d.nxc * d.nyc;
    for (int w0 = 0; w0 < d.anchor_words; w0 += 32) {
        // 32 words of the anchor bitmap per load; skip empty stretches at once
        const uint32_t my_bits = (w0 + lane < d.anchor_words) ? anc[w0 + lane] : 0u;
        unsigned nonzero = __ballot_sync(0xffffffffu, my_bits != 0);
        while (nonzero) {
            const int wl = __ffs(nonzero) - 1;
            nonzero &= nonzero - 1;
            uint32_t bits = __shfl_sync(0xffffffffu, my_bits, wl);
            const int wi = w0 + wl;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const int item = wi * 32 + b;
                if (item >= n_cand) break;
                n_anchor++;
                const int i = item / d.nyc, j = item - i * d.nyc;
                const unsigned x = 1 + 2 * i, y = 1 + 2 * j;
                const int idx = (int)(y * (unsigned)W + x);
                if ((edge[idx >> 5] >> (idx & 31)) & 1u) continue;
                edw_ensure(w, (int)x, (int)y, lane);
                const bool horiz = ed_horizontal(edw_at(w, (int)x, (int)y));
                const int len1 = edw_walk(w, edge, x, y, horiz ? ED_RIGHT : ED_DOWN, p1, d.part_cap, st, lane);
                __syncwarp();
                if (lane == 0) edge[idx >> 5] &= ~(1u << (idx & 31));  // the anchor starts the second part too
                __syncwarp();
                const int len2 = edw_walk(w, edge, x, y, horiz ? ED_LEFT : ED_UP, p2, d.part_cap, st, lane);
                if (len1 + len2 < ED_MIN_LINE_LEN + 1) continue;  // short edge: pixels stay marked, chain dropped
                kept1 += len1; kept2 += len2;
                if (len1 > d.part_cap || len2 > d.part_cap || n_chains >= d.max_edges + 1 || n_px + len1 + len2 - 1 > d.chain_cap) { overflow = true; continue; }
                if (lane == 0) sid[n_chains] = n_px;
                n_chains++;
                __syncwarp();  // the walk's global stores are visible to the whole warp (same CTA: ordered after the barrier)
                __threadfence_block();
                for (int t = lane; t < len1; t += 32) chain[n_px + t] = p1[len1 - 1 - t];
                for (int t = 1 + lane; t < len2; t += 32) chain[n_px + len1 + t - 1] = p2[t];
                n_px += len1 + len2 - 1;
                __syncwarp();
            }
        }
    }
    if (lane != 0) return;
    // EdgeDrawing's capacity errors (:2329-2341) abort the detection of the frame ("Line Detection not finished")
    if (overflow || n_chains > d.max_edges || kept1 > d.part_cap || kept2 > d.part_cap) { B.n_chains[frame] = -1; return; }
    sid[n_chains] = n_px;
    B.n_chains[frame] = n_chains;
    atomicAdd(B.stats + 0, n_anchor);
    atomicAdd(B.stats + 1, (unsigned long long)n_px);
    atomicAdd(B.stats + 2, (unsigned long long)n_chains);
}

// ---- line fitting, one WARP per chain (the production kernel; ed_fit() is the one-thread transcription the host emulation shares) ------------
// Same decisions and the same numbers as ed_fit(), with the per-pixel work spread over the lanes:
//  * the least-squares sums are sums of integer products (exact in double, < 2^53), so a lane-strided accumulation + butterfly gives the very
//    value the sequential loop gives; they are rounded to float once, like there;
//  * the fit error of the 15 initial pixels is an inexact double sum: the squares come from 15 lanes, the additions run in pixel order;
//  * extension: 32 chain pixels per step -- every lane tests one pixel against the line, a ballot gives the outlier mask, and the
//    reference's "stop at the 4th consecutive outlier" is the first run of four ones in that mask (with the run carried in from the
//    previous step prepended), found with shifts and ffs;
//  * validation: gradient sums (integers) and the per-pixel direction test (one det_atan2 per pixel) lane-strided, counts added up;
//  * everything scalar (the 2 x 2 solve, the NFA) is evaluated by all lanes on identical operands.
__device__ __forceinline__ double warp_sum_exact(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void ed_sums_warp(const ushort2* p, int n, bool horiz, int lane, double& sa2, double& sa, double& sab, double& sb) {
    double a2 = 0, a1 = 0, ab = 0, b1 = 0;
    for (int i = lane; i < n; i += 32) {
        const ushort2 q = p[i];
        const double a = (double)(horiz ? q.x : q.y), b = (double)(horiz ? q.y : q.x);
        a2 += a * a; a1 += a; ab += a * b; b1 += b;
    }
    sa2 = warp_sum_exact(a2); sa = warp_sum_exact(a1); sab = warp_sum_exact(ab); sb = warp_sum_exact(b1);
}

__device__ void ed_fit_warp(const EdBuffers& B, const EdDims& d, int frame, int chain_id, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int W = d.w, H = d.h;
    const uint16_t* gd = B.gd + (size_t)frame * W * H;
    const short2* grad = B.grad + (size_t)frame * W * H;
    const int* sid = B.chain_sid + (size_t)frame * (d.max_edges + 2);
    const ushort2* ch = B.chain_px + (size_t)frame * d.chain_cap;
    ushort2* ln = B.line_px + (size_t)frame * d.chain_cap;
    EdLine* stage = B.stage + (size_t)frame * d.stage_cap;
    int S = sid[chain_id];
    const int E = sid[chain_id + 1];
    const int slot_base = S / ED_MIN_LINE_LEN, slot_cap = (E - S) / ED_MIN_LINE_LEN;   // see ed_fit
    for (int k = lane; k < slot_cap; k += 32) stage[slot_base + k].n_px = 0;
    int n_lines = 0;
    const double logNT = 2.0 * (log10((double)(unsigned)W) + log10((double)(unsigned)H));
    int off = S, newOffsetS = off;
    double lineFitErr = 0;
    double eq[2] = {0, 0};
    EdFit fit{};
    unsigned long long n_staged = 0;
    while (E > S + ED_MIN_LINE_LEN) {
        bool horiz = false;
        while (E > S + ED_MIN_LINE_LEN) {
            const ushort2 first = ch[S];
            horiz = ed_horizontal(gd[first.y * W + first.x]);
            double sa2, sa, sab, sb;
            ed_sums_warp(ch + S, ED_MIN_LINE_LEN, horiz, lane, sa2, sa, sab, sb);
            fit.ATA[0] = (float)sa2; fit.ATA[1] = (float)sa; fit.ATA[2] = (float)sa; fit.ATA[3] = (float)ED_MIN_LINE_LEN;
            fit.ATV[0] = (float)sab; fit.ATV[1] = (float)sb;
            ed_solve(fit, eq);
            double c2 = 0;
            if (lane < ED_MIN_LINE_LEN) {
                const ushort2 q = ch[S + lane];
                const double xx = (double)q.x, yy = (double)q.y;
                const double c = horiz ? yy - xx * eq[0] - eq[1] : xx - yy * eq[0] - eq[1];
                c2 = c * c;
            }
            double fe = 0;
#pragma unroll
            for (int i = 0; i < ED_MIN_LINE_LEN; i++) fe += __shfl_sync(FULL, c2, i);
            lineFitErr = sqrt(fe);
            if (lineFitErr <= ED_FIT_ERR) break;
            S += ED_SKIP;
        }
        if (lineFitErr > ED_FIT_ERR) break;
        const int lineStart = off;
        {
            const ushort2 first = ch[S];
            horiz = ed_horizontal(gd[first.y * W + first.x]);
        }
        double coef1 = 0;
        bool bExtended = true, bFirstTry = true;
        int tryTimes = 0;
        while (bExtended) {
            tryTimes++;
            if (bFirstTry) {
                bFirstTry = false;
                if (lane < ED_MIN_LINE_LEN) ln[off + lane] = ch[S + lane];
                off += ED_MIN_LINE_LEN; S += ED_MIN_LINE_LEN;
                __syncwarp();
            } else {
                const int length = off - lineStart, newLength = off - newOffsetS;
                if (length > 0 && newLength > 0) {
                    const ushort2 first = ln[lineStart];
                    const bool hz = ed_horizontal(gd[first.y * W + first.x]);
                    double sa2, sa, sab, sb;
                    ed_sums_warp(ln + newOffsetS, newLength, hz, lane, sa2, sa, sab, sb);
                    fit.ATA[0] = fit.ATA[0] + (float)sa2; fit.ATA[1] = fit.ATA[1] + (float)sa; fit.ATA[2] = fit.ATA[2] + (float)sa;
                    fit.ATA[3] = fit.ATA[3] + (float)newLength;
                    fit.ATV[0] = fit.ATV[0] + (float)sab; fit.ATV[1] = fit.ATV[1] + (float)sb;
                    ed_solve(fit, eq);
                }
            }
            coef1 = horiz ? 1 / sqrt(eq[0] * eq[0] + 1) : 1 / sqrt(1 + eq[0] * eq[0]);
            int numOfOutlier = 0;
            newOffsetS = off;
            while (E > S) {
                const int n = min(32, E - S);
                ushort2 q = make_ushort2(0, 0);
                bool out = false;
                if (lane < n) {
                    q = ch[S + lane];
                    const double xx = (double)q.x, yy = (double)q.y;
                    const double dist = horiz ? fabs(eq[0] * xx - yy + eq[1]) * coef1 : fabs(xx - eq[0] * yy - eq[1]) * coef1;
                    out = dist > ED_FIT_ERR;
                }
                const unsigned m = __ballot_sync(FULL, out);
                // the run of outliers that entered this step (<= 3) goes in front of the mask; a run of four ends the extension
                const unsigned long long M = ((unsigned long long)m << numOfOutlier) | ((1ull << numOfOutlier) - 1ull);
                const unsigned long long Q = M & (M >> 1) & (M >> 2) & (M >> 3);
                int consumed = n;
                bool stop = false;
                if (Q) {
                    const int p = (__ffsll((long long)Q) - 1) + 3 - numOfOutlier;  // position (in this step) of the 4th consecutive outlier
                    if (p < n) { consumed = p + 1; stop = true; }
                }
                if (lane < consumed) ln[off + lane] = q;
                off += consumed; S += consumed;
                if (stop) { numOfOutlier = 4; break; }
                // the run of outliers at the end of this step (all of it an outlier: the run that came in grows)
                const unsigned inv = ~m & ((n == 32) ? 0xffffffffu : ((1u << n) - 1u));
                numOfOutlier = inv ? (n - 1 - (31 - __clz(inv))) : numOfOutlier + n;
            }
            __syncwarp();
            off -= numOfOutlier;
            S -= numOfOutlier;
            if (!(off - newOffsetS > 0 && tryTimes < ED_TRY_TIME)) bExtended = false;
        }
        double le[3];
        if (horiz) { le[0] = eq[0] * coef1; le[1] = -1 * coef1; le[2] = eq[1] * coef1; }
        else { le[0] = 1 * coef1; le[1] = -eq[0] * coef1; le[2] = -eq[1] * coef1; }
        // LineValidation_ (:2793-2873)
        bool ok = true;
        float direction = 0.f;
        {
            const int n = off - lineStart;
            int mgx = 0, mgy = 0;
            for (int i = lane; i < n; i += 32) {
                const ushort2 q = ln[lineStart + i];
                const short2 g = grad[q.y * W + q.x];
                mgx += g.x; mgy += g.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { mgx += __shfl_xor_sync(FULL, mgx, o); mgy += __shfl_xor_sync(FULL, mgy, o); }
            const double dx = fabs(le[1]), dy = fabs(le[0]);
            if (mgx == 0 && mgy == 0) ok = false;
            if (ok) {
                if (mgx > 0 && mgy >= 0) direction = (float)det_atan2(-dy, dx);
                if (mgx <= 0 && mgy > 0) direction = (float)det_atan2(dy, dx);
                if (mgx < 0 && mgy <= 0) direction = (float)det_atan2(dy, -dx);
                if (mgx >= 0 && mgy < 0) direction = (float)det_atan2(-dy, -dx);
                const double ad = fabs((double)direction);
                if (ad < 0.15 || ED_PI - ad < 0.15) {
                    if (fabs(le[2]) < 10 || fabs((double)(unsigned)H - fabs(le[2])) < 10) ok = false;
                }
                if (ok && fabs(ad - ED_PI * 0.5) < 0.15) {
                    if (fabs(le[2]) < 10 || fabs((double)(unsigned)W - fabs(le[2])) < 10) ok = false;
                }
            }
            if (ok) {
                int k = 0;
                for (int i = lane; i < n; i += 32) {
                    const ushort2 q = ln[lineStart + i];
                    const short2 g = grad[q.y * W + q.x];
                    const double pd = det_atan2(-(double)g.x, (double)g.y);
                    const double dis = fabs((double)direction - pd);
                    if (fabs(2 * ED_PI - dis) < 0.392699 || dis < 0.392699) k++;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(FULL, k, o);
                ok = ed_nfa(n, k, 0.125, logNT) > 0;
            }
        }
        if (ok) {
            if (lane == 0) {
                const double a1 = le[1] * le[1], a2 = le[0] * le[0], a3 = le[0] * le[1], a4 = le[2] * le[0], a5 = le[2] * le[1];
                EdLine L;
                double Px = (double)ln[lineStart].x, Py = (double)ln[lineStart].y;
                L.ep[0] = (float)(a1 * Px - a3 * Py - a4);
                L.ep[1] = (float)(a2 * Py - a3 * Px - a5);
                Px = (double)ln[off - 1].x; Py = (double)ln[off - 1].y;
                L.ep[2] = (float)(a1 * Px - a3 * Py - a4);
                L.ep[3] = (float)(a2 * Py - a3 * Px - a5);
                L.direction = direction;
                L.n_px = off - lineStart;
                if (n_lines < slot_cap) stage[slot_base + n_lines] = L;
            }
            n_lines++;
            n_staged++;
        } else {
            off = lineStart;
        }
    }
    if (lane == 0 && n_staged) atomicAdd(B.stats + 3, n_staged);
}

constexpr int EDF_WARPS = 4;
constexpr int EDF_CTAS_PER_FRAME = 16;  // 64 warps per frame stride over its chains (a frame has tens of chains, the capacity is thousands)
__global__ void __launch_bounds__(32 * EDF_WARPS) k_ed_fit_warp(EdBuffers B, EdDims d) {
    const int frame = blockIdx.y, n = B.n_chains[frame];
    for (int chain = blockIdx.x * EDF_WARPS + (threadIdx.x >> 5); chain < n; chain += gridDim.x * EDF_WARPS) ed_fit_warp(B, d, frame, chain, threadIdx.x & 31);
}

__global__ void __launch_bounds__(32) k_ed_emit(EdBuffers B, EdDims d, int filter, float length_thres, int max_lines) {
    if (threadIdx.x == 0) ed_emit(B, d, blockIdx.x, filter, length_thres, max_lines);
}

struct EdState {
    bool uploaded = false, ran = false, timed_last = false;
    EdDims d{};
    csb_lsd_params params{};
    DevBuf d_gray, d_grad, d_gd, d_anchors, d_edge, d_part1, d_part2, d_chain, d_line, d_sid, d_nch, d_stage, d_lines, d_nlines, d_stats, d_keyl, d_weights,
        d_desc, d_descf;
    bool weights_set = false, described = false, desc_float = false;
    HostBuf h_gray, h_out;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int launches_last = 0;
};

void edlines_release(EdState*& s) {
    if (!s) return;
    DevBuf* bufs[] = {&s->d_gray, &s->d_grad, &s->d_gd, &s->d_anchors, &s->d_edge, &s->d_part1, &s->d_part2, &s->d_chain, &s->d_line,
                      &s->d_sid, &s->d_nch, &s->d_stage, &s->d_lines, &s->d_nlines, &s->d_stats, &s->d_keyl, &s->d_weights, &s->d_desc, &s->d_descf};
    for (DevBuf* b : bufs) b->release();
    s->h_gray.release();
    s->h_out.release();
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    delete s;
    s = nullptr;
}

static EdBuffers ed_buffers(EdState& s) {
    EdBuffers B{};
    B.grad = s.d_grad.as<short2>();
    B.gd = s.d_gd.as<uint16_t>();
    B.anchors = s.d_anchors.as<uint32_t>();
    B.edge = s.d_edge.as<uint32_t>();
    B.part1 = s.d_part1.as<ushort2>();
    B.part2 = s.d_part2.as<ushort2>();
    B.chain_px = s.d_chain.as<ushort2>();
    B.line_px = s.d_line.as<ushort2>();
    B.chain_sid = s.d_sid.as<int>();
    B.n_chains = s.d_nch.as<int>();
    B.stage = s.d_stage.as<EdLine>();
    B.lines = s.d_lines.as<float>();
    B.keyl = s.d_keyl.as<float2>();
    B.n_lines = s.d_nlines.as<int>();
    B.stats = s.d_stats.as<unsigned long long>();
    return B;
}

}  // namespace csb

using namespace csb;

extern "C" {

int csb_edlines_upload(csb_context* c, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params) {
    if (!c || !gray || !params || n_frames <= 0 || width < 8 || height < 8 || width > 32767 || height > 32767 || (int64_t)width * height > (1 << 24) ||
        params->max_lines <= 0)
        return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    if (!c->edlines) {
        c->edlines = new EdState();
        for (auto& e : c->edlines->ev) CSB_CUDA(c, cudaEventCreate(&e));
    }
    EdState& s = *c->edlines;
    s.uploaded = false;
    s.ran = false;
    s.described = false;
    s.params = *params;
    s.d = ed_make_dims(width, height, n_frames);
    const EdDims& d = s.d;
    const size_t npx = (size_t)width * height * n_frames, nf = (size_t)n_frames;
    CSB_CUDA(c, s.d_gray.ensure(npx));
    CSB_CUDA(c, s.d_grad.ensure(npx * 4));
    CSB_CUDA(c, s.d_gd.ensure(npx * 2));
    CSB_CUDA(c, s.d_anchors.ensure(nf * d.anchor_words * 4));
    CSB_CUDA(c, s.d_edge.ensure(nf * d.edge_words * 4));
    CSB_CUDA(c, s.d_part1.ensure(nf * std::max(d.part_cap, 1) * 4));
    CSB_CUDA(c, s.d_part2.ensure(nf * std::max(d.part_cap, 1) * 4));
    CSB_CUDA(c, s.d_chain.ensure(nf * std::max(d.chain_cap, 1) * 4));
    CSB_CUDA(c, s.d_line.ensure(nf * std::max(d.chain_cap, 1) * 4));
    CSB_CUDA(c, s.d_sid.ensure(nf * (d.max_edges + 2) * 4));
    CSB_CUDA(c, s.d_nch.ensure(nf * 4));
    CSB_CUDA(c, s.d_stage.ensure(nf * d.stage_cap * sizeof(EdLine)));
    CSB_CUDA(c, s.d_lines.ensure_zeroed(nf * params->max_lines * 16, c->stream));
    CSB_CUDA(c, s.d_keyl.ensure_zeroed(nf * params->max_lines * 8, c->stream));
    CSB_CUDA(c, s.d_nlines.ensure(nf * 4));
    CSB_CUDA(c, s.d_stats.ensure(64));
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, gray) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        CSB_CUDA(c, cudaMemcpyAsync(s.d_gray.p, gray, npx, cudaMemcpyHostToDevice, c->stream));
    } else {
        CSB_CUDA(c, cudaStreamSynchronize(c->stream));  // the staging buffer may still feed an earlier upload
        CSB_CUDA(c, s.h_gray.ensure(npx));
        std::memcpy(s.h_gray.p, gray, npx);
        CSB_CUDA(c, cudaMemcpyAsync(s.d_gray.p, s.h_gray.p, npx, cudaMemcpyHostToDevice, c->stream));
    }
    s.h2d_bytes = (int64_t)npx;
    s.uploaded = true;
    return CSB_OK;
}

int csb_edlines_run(csb_context* c, int timed) {
    if (!c || !c->edlines || !c->edlines->uploaded) {
        if (c) c->err = "csb_edlines_run before csb_edlines_upload";
        return CSB_ERR_STATE;
    }
    EdState& s = *c->edlines;
    CSB_CUDA(c, cudaSetDevice(c->device));
    const EdDims& d = s.d;
    EdBuffers B = ed_buffers(s);
    cudaStream_t st = c->stream;
    const size_t npx = (size_t)d.w * d.h * d.n_frames;
    CSB_CUDA(c, cudaMemsetAsync(s.d_stats.p, 0, 64, st));
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[0], st));
    lbd_launch_grad(s.d_gray.as<uint8_t>(), s.d_grad.as<short2>(), d.w, d.h, d.n_frames, st, c->blur_generation);
    k_ed_pixel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(B, d, npx);
    k_ed_anchor<<<dim3((d.anchor_words + 127) / 128, d.n_frames), 128, 0, st>>>(B, d);
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[1], st));
    {
        // the warp-per-frame kernel keeps the frame's edge bitmap in shared memory; a frame whose bitmap does not fit (> ~1.8 Mpx) takes
        // the one-thread transcription with the bitmap in global memory
        const size_t smem = 4 * (size_t)((d.edge_words + 3) & ~3);
        if (smem <= (size_t)c->max_smem_optin - 1024) {
            CSB_CUDA(c, cudaFuncSetAttribute(k_ed_draw_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_ed_draw_warp<<<d.n_frames, 32, smem, st>>>(B, d);
        } else {
            CSB_CUDA(c, cudaMemsetAsync(s.d_edge.p, 0, (size_t)d.n_frames * d.edge_words * 4, st));
            k_ed_draw<<<d.n_frames, 32, 0, st>>>(B, d);
        }
    }
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[2], st));
    k_ed_fit_warp<<<dim3(EDF_CTAS_PER_FRAME, d.n_frames), 32 * EDF_WARPS, 0, st>>>(B, d);
    k_ed_emit<<<d.n_frames, 32, 0, st>>>(B, d, s.params.filter, s.params.line_length_thres, s.params.max_lines);
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[3], st));
    CSB_CUDA(c, cudaGetLastError());
    s.launches_last = 6;
    s.timed_last = timed != 0;
    s.ran = true;
    s.described = false;
    return CSB_OK;
}

int csb_edlines_download(csb_context* c, float* lines_out, int32_t* n_lines_out, csb_edlines_stats* stats) {
    if (!c || !c->edlines || !c->edlines->ran) {
        if (c) c->err = "csb_edlines_download before csb_edlines_run";
        return CSB_ERR_STATE;
    }
    EdState& s = *c->edlines;
    CSB_CUDA(c, cudaSetDevice(c->device));
    const int nf = s.d.n_frames, cap = s.params.max_lines;
    const size_t lb = (size_t)nf * cap * 16, nb = (size_t)nf * 4;
    CSB_CUDA(c, s.h_out.ensure(lb + 2 * nb + 64));
    char* h = s.h_out.as<char>();
    CSB_CUDA(c, cudaMemcpyAsync(h + lb, s.d_nlines.p, nb, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaMemcpyAsync(h + lb + nb, s.d_nch.p, nb, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaMemcpyAsync(h + lb + 2 * nb, s.d_stats.p, 64, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    const int32_t* nl = reinterpret_cast<const int32_t*>(h + lb);
    const int32_t* nch = reinterpret_cast<const int32_t*>(h + lb + nb);
    int rows = 0;
    for (int f = 0; f < nf; f++) rows = std::max(rows, std::min(nl[f], cap));
    const size_t pitch = (size_t)cap * 16;
    if (rows > 0) CSB_CUDA(c, cudaMemcpy2DAsync(h, pitch, s.d_lines.p, pitch, (size_t)rows * 16, nf, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    s.d2h_bytes = (int64_t)((size_t)rows * 16 * nf + 2 * nb + 64);
    bool overflow = false;
    int64_t total = 0;
    int failed = 0;
    for (int f = 0; f < nf; f++) {
        overflow |= nl[f] > cap;
        total += std::min(nl[f], cap);
        failed += nch[f] < 0;
    }
    if (lines_out)
        for (int f = 0; f < nf; f++) std::memcpy(reinterpret_cast<char*>(lines_out) + f * pitch, h + f * pitch, (size_t)std::min(nl[f], cap) * 16);
    if (n_lines_out) std::memcpy(n_lines_out, nl, nb);
    if (stats) {
        const unsigned long long* st = reinterpret_cast<const unsigned long long*>(h + lb + 2 * nb);
        std::memset(stats, 0, sizeof(*stats));
        stats->n_lines = total;
        stats->n_anchors = (int64_t)st[0];
        stats->n_chain_px = (int64_t)st[1];
        stats->n_chains = (int64_t)st[2];
        stats->h2d_bytes = s.h2d_bytes;
        stats->d2h_bytes = s.d2h_bytes;
        stats->n_kernel_launches = s.launches_last;
        stats->n_frames_failed = failed;
        if (s.timed_last) {
            cudaEventElapsedTime(&stats->gpu_ms_maps, s.ev[0], s.ev[1]);
            cudaEventElapsedTime(&stats->gpu_ms_draw, s.ev[1], s.ev[2]);
            cudaEventElapsedTime(&stats->gpu_ms_fit, s.ev[2], s.ev[3]);
        }
    }
    if (overflow) {
        c->err = "csb_edlines: more segments than max_lines in at least one frame";
        return CSB_ERR_CAPACITY;
    }
    return CSB_OK;
}

int csb_edlines_describe(csb_context* c, int want_float) {
    if (!c || !c->edlines || !c->edlines->ran) {
        if (c) c->err = "csb_edlines_describe before csb_edlines_run";
        return CSB_ERR_STATE;
    }
    EdState& s = *c->edlines;
    CSB_CUDA(c, cudaSetDevice(c->device));
    const size_t rows = (size_t)s.d.n_frames * s.params.max_lines;
    if (!s.weights_set) {
        CSB_CUDA(c, lbd_upload_weights(c->stream));
        s.weights_set = true;
    }
    CSB_CUDA(c, s.d_desc.ensure_zeroed(rows * 32, c->stream));
    if (want_float) CSB_CUDA(c, s.d_descf.ensure_zeroed(rows * 72 * 4, c->stream));
    CSB_CUDA(c, s.d_weights.ensure((size_t)(s.d.n_frames + 1) * 4 + 64));  // scratch of the describe stage: line prefix + work counters
    EdBuffers B = ed_buffers(s);
    // the warp-cooperative descriptor kernel of lbd.cu on the detector's own key-line fields (k_lbdk_line is the one-thread-per-line
    // transcription the host emulation shares)
    int* prefix = s.d_weights.as<int>();
    unsigned long long* ctr = reinterpret_cast<unsigned long long*>(s.d_weights.as<char>() + (((size_t)(s.d.n_frames + 1) * 4 + 15) & ~(size_t)15));
    CSB_CUDA(c, lbd_describe_keylines(B.grad, B.lines, B.keyl, B.n_lines, s.d.n_frames, s.params.max_lines, s.d.w, s.d.h, s.d_desc.as<uint8_t>(),
                                      want_float ? s.d_descf.as<float>() : nullptr, prefix, ctr, c->num_sms, c->stream));
    CSB_CUDA(c, cudaGetLastError());
    s.described = true;
    s.desc_float = want_float != 0;
    return CSB_OK;
}

int csb_edlines_download_descriptors(csb_context* c, uint8_t* desc_out, float* desc_float_out, int32_t* n_lines_out, int64_t capacity_rows) {
    if (!c || !c->edlines || !c->edlines->described) {
        if (c) c->err = "csb_edlines_download_descriptors before csb_edlines_describe";
        return CSB_ERR_STATE;
    }
    EdState& s = *c->edlines;
    if (desc_float_out && !s.desc_float) {
        c->err = "csb_edlines_download_descriptors: float descriptors were not requested";
        return CSB_ERR_STATE;
    }
    CSB_CUDA(c, cudaSetDevice(c->device));
    const int nf = s.d.n_frames, cap = s.params.max_lines;
    std::vector<int32_t> nl(nf);
    CSB_CUDA(c, cudaMemcpyAsync(nl.data(), s.d_nlines.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    int64_t total = 0;
    int rows = 0;
    for (int f = 0; f < nf; f++) { nl[f] = std::min(nl[f], cap); total += nl[f]; rows = std::max(rows, nl[f]); }
    if (n_lines_out) std::memcpy(n_lines_out, nl.data(), (size_t)nf * 4);
    if (total > capacity_rows) {
        c->err = "csb_edlines_download_descriptors: more lines than capacity_rows";
        return CSB_ERR_CAPACITY;
    }
    if (rows == 0) return CSB_OK;
    auto pull = [&](const void* dev, size_t rb, void* out) -> cudaError_t {
        if (!out) return cudaSuccess;
        cudaError_t e = s.h_out.ensure((size_t)rows * nf * rb);
        if (e != cudaSuccess) return e;
        char* h = s.h_out.as<char>();
        e = cudaMemcpy2DAsync(h, (size_t)rows * rb, dev, (size_t)cap * rb, (size_t)rows * rb, nf, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return e;
        size_t o = 0;
        for (int f = 0; f < nf; f++) {
            std::memcpy(reinterpret_cast<char*>(out) + o, h + (size_t)f * rows * rb, (size_t)nl[f] * rb);
            o += (size_t)nl[f] * rb;
        }
        return cudaSuccess;
    };
    CSB_CUDA(c, pull(s.d_desc.p, 32, desc_out));
    CSB_CUDA(c, pull(s.d_descf.p, 288, desc_float_out));
    return CSB_OK;
}

int csb_edlines_detect_batch(csb_context* c, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params, float* lines_out,
                             int32_t* n_lines_out, csb_edlines_stats* stats) {
    int rc = csb_edlines_upload(c, gray, n_frames, width, height, params);
    if (rc != CSB_OK) return rc;
    rc = csb_edlines_run(c, stats != nullptr);
    if (rc != CSB_OK) return rc;
    return csb_edlines_download(c, lines_out, n_lines_out, stats);
}

}  // extern "C"
