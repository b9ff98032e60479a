// edlines_dev.cuh -- EDLines (the use_LSD = false branch of line_lbd_detect) as one-thread-per-item device functions.
//
// Reference: line_lbd/libs/binary_descriptor.cpp:796-1148 (OctaveKeyLines, one octave), :1583-2380 (EdgeDrawing), :2383-2630 (EDline),
// :2632-2790 (LeastSquaresLineFit_), :2793-2873 (LineValidation_), line_descriptor/descriptor.hpp:655-848 (nfa); the restatement the
// results are compared with is oracle/oracle_edlines.cpp.
//
// Every stage is a plain function of an item index with no shared memory and no synchronisation, so the SAME source runs under nvcc
// (kernels in edlines.cu are one-line wrappers) and under g++ (tests/emul/edlines_emul.cpp loops over the items): the CPU test-suite
// executes this very code against the oracle, which is how it could be validated before a GPU was available for it.
//
//   ed_pixel   : item = pixel.      packed gradient map: bits 0..14 = thresholded (|dx| + |dy|) / 4 (MatExpr rounding), bit 15 = direction
//                                   (1: |dx| < |dy|, a horizontal edge).
//   ed_anchor  : item = candidate.  anchors are the odd-x / odd-y pixels whose gradient beats both neighbours across the edge by 8
//                                   (:1643-1670); one bit per candidate, COLUMN-major, so that the drawing stage meets the set bits in the
//                                   reference's scan order.
//   ed_draw    : item = frame.      smart routing from every anchor (two walks), the `edge` bitmap, chain assembly (first part reversed +
//                                   second part without the anchor), the reference's capacity checks.  Inherently sequential per frame.
//   ed_fit     : item = chain.      initial 15-pixel least-squares segment, extension with up to 3 trailing outliers, up to 6 refits,
//                                   validation (direction from the mean gradient, border lines, aligned-pixel count, NFA), end points.
//                                   Chains are independent; a chain's lines go to its own slots (chain start / 15) of the frame's staging.
//   ed_emit    : item = frame.      chains in order, lines in order: length, end-point order by direction (OctaveKeyLines :1075-1140),
//                                   length filter (line_lbd_allclass.cpp:200-208), compact float4 rows.
//
// Arithmetic: integer maps; the fits keep their sums in float exactly as the reference's cv::Mat_<float> members do (each sum of integer
// products is exact in double and rounded to float once; ATA / ATV grow by float additions); everything else is double, no contraction.
// atan2 is det_atan2 on both sides; log / exp / pow / sinh / log10 of the NFA come from the platform's libm and only feed `> 0`.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#include "csb_math.cuh"

namespace csb {

constexpr int ED_MIN_LINE_LEN = 15, ED_TRY_TIME = 6, ED_SKIP = 2, ED_ANCHOR_THR = 8, ED_GRAD_THR = 80;
constexpr double ED_FIT_ERR = 1.6, ED_MLN10 = 2.30258509299404568402, ED_PI = 3.14159265358979323846;
constexpr int ED_UP = 1, ED_RIGHT = 2, ED_DOWN = 3, ED_LEFT = 4;

struct EdDims {
    int w, h, n_frames;
    int nxc, nyc;        // anchor candidates per row / column: x = 1 + 2 i < w - 1, y = 1 + 2 j < h - 1
    int anchor_words;    // 32-bit words of the column-major candidate bitmap (per frame)
    int edge_words;      // 32-bit words of the edge bitmap (per frame)
    int part_cap;        // edgePixelArraySize = w h / 5: capacity of each walk scratch, also the reference's pixel limit
    int max_edges;       // maxNumOfEdge = part_cap / 20
    int chain_cap;       // chain pixels per frame (2 part_cap)
    int stage_cap;       // staged lines per frame (chain_cap / 15 + 2)
};

inline EdDims ed_make_dims(int w, int h, int n_frames) {
    EdDims d{};
    d.w = w; d.h = h; d.n_frames = n_frames;
    d.nxc = (w - 1) / 2;
    d.nyc = (h - 1) / 2;
    d.anchor_words = (d.nxc * d.nyc + 31) / 32;
    d.edge_words = (w * h + 31) / 32;
    d.part_cap = (w * h) / 5;            // edgePixelArraySize (:1589)
    d.max_edges = d.part_cap / 20;       // maxNumOfEdge (:1590)
    d.chain_cap = 2 * d.part_cap;
    d.stage_cap = d.chain_cap / ED_MIN_LINE_LEN + 2;
    return d;
}

struct EdLine {          // one staged line
    float ep[4];
    float direction;
    int n_px;            // 0: empty slot
};

struct EdBuffers {
    const short2* grad;       // n_frames x h x w {dx, dy} (k_lbd_grad / k_lbd_grad4)
    uint16_t* gd;             // n_frames x h x w packed gradient + direction
    uint32_t* anchors;        // n_frames x anchor_words
    uint32_t* edge;           // n_frames x edge_words
    ushort2* part1;           // n_frames x part_cap   walk scratch
    ushort2* part2;           // n_frames x part_cap
    ushort2* chain_px;        // n_frames x chain_cap
    ushort2* line_px;         // n_frames x chain_cap  (EDline's line arrays, chain-local)
    int* chain_sid;           // n_frames x (max_edges + 2)
    int* n_chains;            // n_frames (-1: the reference's capacity error)
    EdLine* stage;            // n_frames x stage_cap
    float* lines;             // n_frames x max_lines x 4
    float2* keyl;             // n_frames x max_lines {lineDirection_, numOfPixels} of the emitted lines (optional: the descriptor stage reads it)
    int* n_lines;             // n_frames
    unsigned long long* stats;  // [0] anchors, [1] chain pixels, [2] chains, [3] staged lines
};

// ---------------------------------------------------------------------------------------------------------------------------------
CSB_HD void ed_pixel(const EdBuffers& B, const EdDims& d, size_t item) {
    const short2 g = B.grad[item];
    const int ax = g.x < 0 ? -(int)g.x : (int)g.x, ay = g.y < 0 ? -(int)g.y : (int)g.y;
    const int sum = ax + ay;
    const int t = sum > ED_GRAD_THR + 1 ? sum : 0;       // cv::threshold(..., gradienThreshold_ + 1, 255, THRESH_TOZERO)
    // `gImg_ / 4` on CV_16S = saturate_cast<short>(t * 0.25): round half to even.  t = 4 q + r: r = 0, 1 -> q; r = 3 -> q + 1; r = 2 -> nearest even
    const int q = t >> 2, r = t & 3;
    const int v = q + ((r == 3) || (r == 2 && (q & 1)) ? 1 : 0);
    B.gd[item] = (uint16_t)(v | (ax < ay ? 0x8000 : 0));
}

CSB_HD int ed_g(uint16_t v) { return v & 0x7fff; }
CSB_HD bool ed_horizontal(uint16_t v) { return (v & 0x8000) != 0; }

// item = candidate index inside a frame (column-major: i * nyc + j); returns true if (1 + 2 i, 1 + 2 j) is an anchor
CSB_HD bool ed_anchor(const EdBuffers& B, const EdDims& d, int frame, int item) {
    const int i = item / d.nyc, j = item - i * d.nyc;
    const int x = 1 + 2 * i, y = 1 + 2 * j;
    const uint16_t* gd = B.gd + (size_t)frame * d.w * d.h;
    const int idx = y * d.w + x;
    const uint16_t c = gd[idx];
    const int g = ed_g(c);
    if (ed_horizontal(c)) return g >= ed_g(gd[idx - d.w]) + ED_ANCHOR_THR && g >= ed_g(gd[idx + d.w]) + ED_ANCHOR_THR;
    return g >= ed_g(gd[idx - 1]) + ED_ANCHOR_THR && g >= ed_g(gd[idx + 1]) + ED_ANCHOR_THR;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// smart routing
// ---------------------------------------------------------------------------------------------------------------------------------
struct EdWalkState {
    unsigned lastX, lastY;  // carried from walk to walk like the reference's locals
};

// one walk (the reference repeats this body four times: :1716-1850, :1858-1990, :2000-2160, :2170-2312); returns the pixels written
CSB_HD int ed_walk(const uint16_t* gd, uint32_t* edge, int W, int H, unsigned x, unsigned y, int lastDirection, ushort2* out, int cap, EdWalkState& st) {
    int n = 0;
    int idx = (int)(y * (unsigned)W + x);
    while (ed_g(gd[idx]) > 0 && !((edge[idx >> 5] >> (idx & 31)) & 1u)) {
        edge[idx >> 5] |= 1u << (idx & 31);
        if (n < cap) out[n] = make_ushort2((unsigned short)x, (unsigned short)y);
        n++;
        int shouldGo = 0;
        // neighbours are compared through an unsigned-char cast in the reference (:1746-1748)
#define ED_GV(off) ((unsigned char)ed_g(gd[idx + (off)]))
        if (ed_horizontal(gd[idx])) {
            if (lastDirection == ED_UP || lastDirection == ED_DOWN) shouldGo = (x > st.lastX) ? ED_RIGHT : ED_LEFT;
            st.lastX = x; st.lastY = y;
            if (lastDirection == ED_RIGHT || shouldGo == ED_RIGHT) {
                if (x == (unsigned)W - 1 || y == 0 || y == (unsigned)H - 1) break;
                const unsigned char g1 = ED_GV(-W + 1), g2 = ED_GV(1), g3 = ED_GV(W + 1);
                if (g1 >= g2 && g1 >= g3) { x = x + 1; y = y - 1; }
                else if (g3 >= g2 && g3 >= g1) { x = x + 1; y = y + 1; }
                else { x = x + 1; }
                lastDirection = ED_RIGHT;
            } else if (lastDirection == ED_LEFT || shouldGo == ED_LEFT) {
                if (x == 0 || y == 0 || y == (unsigned)H - 1) break;
                const unsigned char g1 = ED_GV(-W - 1), g2 = ED_GV(-1), g3 = ED_GV(W - 1);
                if (g1 >= g2 && g1 >= g3) { x = x - 1; y = y - 1; }
                else if (g3 >= g2 && g3 >= g1) { x = x - 1; y = y + 1; }
                else { x = x - 1; }
                lastDirection = ED_LEFT;
            }
        } else {
            if (lastDirection == ED_RIGHT || lastDirection == ED_LEFT) shouldGo = (y > st.lastY) ? ED_DOWN : ED_UP;
            st.lastX = x; st.lastY = y;
            if (lastDirection == ED_DOWN || shouldGo == ED_DOWN) {
                if (x == 0 || x == (unsigned)W - 1 || y == (unsigned)H - 1) break;
                const unsigned char g1 = ED_GV(W + 1), g2 = ED_GV(W), g3 = ED_GV(W - 1);
                if (g1 >= g2 && g1 >= g3) { x = x + 1; y = y + 1; }
                else if (g3 >= g2 && g3 >= g1) { x = x - 1; y = y + 1; }
                else { y = y + 1; }
                lastDirection = ED_DOWN;
            } else if (lastDirection == ED_UP || shouldGo == ED_UP) {
                if (x == 0 || x == (unsigned)W - 1 || y == 0) break;
                const unsigned char g1 = ED_GV(-W + 1), g2 = ED_GV(-W), g3 = ED_GV(-W - 1);
                if (g1 >= g2 && g1 >= g3) { x = x + 1; y = y - 1; }
                else if (g3 >= g2 && g3 >= g1) { x = x - 1; y = y - 1; }
                else { y = y - 1; }
                lastDirection = ED_UP;
            }
        }
#undef ED_GV
        idx = (int)(y * (unsigned)W + x);
    }
    return n;
}

// item = frame.  The edge bitmap must be zero on entry.
CSB_HD void ed_draw(const EdBuffers& B, const EdDims& d, int frame) {
    const int W = d.w, H = d.h;
    const uint16_t* gd = B.gd + (size_t)frame * W * H;
    const uint32_t* anc = B.anchors + (size_t)frame * d.anchor_words;
    uint32_t* edge = B.edge + (size_t)frame * d.edge_words;
    ushort2* p1 = B.part1 + (size_t)frame * d.part_cap;
    ushort2* p2 = B.part2 + (size_t)frame * d.part_cap;
    ushort2* chain = B.chain_px + (size_t)frame * d.chain_cap;
    int* sid = B.chain_sid + (size_t)frame * (d.max_edges + 2);
    EdWalkState st{0u, 0u};
    int n_chains = 0, n_px = 0;
    long long kept1 = 0, kept2 = 0;   // the reference's offsetPFirst / offsetPSecond (pixels of kept chains)
    bool overflow = false;
    unsigned long long n_anchor = 0;
    const int n_cand = d.nxc * d.nyc;
    for (int wi = 0; wi < d.anchor_words; wi++) {
        uint32_t bits = anc[wi];
        while (bits) {
            int b = 0;
            while (!((bits >> b) & 1u)) b++;
            bits &= bits - 1;
            const int item = wi * 32 + b;
            if (item >= n_cand) break;
            n_anchor++;
            const int i = item / d.nyc, j = item - i * d.nyc;
            const unsigned x = 1 + 2 * i, y = 1 + 2 * j;
            const int idx = (int)(y * (unsigned)W + x);
            if ((edge[idx >> 5] >> (idx & 31)) & 1u) continue;
            const bool horiz = ed_horizontal(gd[idx]);
            const int len1 = ed_walk(gd, edge, W, H, x, y, horiz ? ED_RIGHT : ED_DOWN, p1, d.part_cap, st);
            edge[idx >> 5] &= ~(1u << (idx & 31));  // the anchor starts the second part too
            const int len2 = ed_walk(gd, edge, W, H, x, y, horiz ? ED_LEFT : ED_UP, p2, d.part_cap, st);
            if (len1 + len2 < ED_MIN_LINE_LEN + 1) continue;  // short edge: pixels stay marked, chain dropped
            kept1 += len1; kept2 += len2;
            if (len1 > d.part_cap || len2 > d.part_cap || n_chains >= d.max_edges + 1 || n_px + len1 + len2 - 1 > d.chain_cap) { overflow = true; continue; }
            sid[n_chains++] = n_px;
            for (int t = len1 - 1; t >= 0; t--) chain[n_px++] = p1[t];
            for (int t = 1; t < len2; t++) chain[n_px++] = p2[t];
        }
    }
    // EdgeDrawing's capacity errors (:2329-2341) abort the detection of the frame ("Line Detection not finished")
    if (overflow || n_chains > d.max_edges || kept1 > d.part_cap || kept2 > d.part_cap) {
        B.n_chains[frame] = -1;
        return;
    }
    sid[n_chains] = n_px;
    B.n_chains[frame] = n_chains;
#if defined(__CUDA_ARCH__)
    atomicAdd(B.stats + 0, n_anchor);
    atomicAdd(B.stats + 1, (unsigned long long)n_px);
    atomicAdd(B.stats + 2, (unsigned long long)n_chains);
#else
    B.stats[0] += n_anchor; B.stats[1] += (unsigned long long)n_px; B.stats[2] += (unsigned long long)n_chains;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------------------
// line fitting
// ---------------------------------------------------------------------------------------------------------------------------------
CSB_HD bool ed_double_equal(double a, double b) {
    if (a == b) return true;
    const double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
    double abs_max = aa > bb ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
CSB_HD double ed_log_gamma(double x) {
    if (x > 15.0) return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
    double b = 0.0;
    for (int n = 0; n < 7; n++) {
        a -= log(x + (double)n);
        b += q[n] * pow(x, (double)n);
    }
    return a + log(b);
}
CSB_HD double ed_nfa(int n, int k, double p, double logNT) {
    const double tolerance = 0.1;
    if (n == 0 || k == 0) return -logNT;
    if (n == k) return -logNT - (double)n * log10(p);
    const double p_term = p / (1.0 - p);
    const double log1term = ed_log_gamma((double)n + 1.0) - ed_log_gamma((double)k + 1.0) - ed_log_gamma((double)(n - k) + 1.0) + (double)k * log(p) +
                            (double)(n - k) * log(1.0 - p);
    double term = exp(log1term);
    if (ed_double_equal(term, 0.0)) {
        if ((double)k > (double)n * p) return -log1term / ED_MLN10 - logNT;
        return -logNT;
    }
    double bin_tail = term;
    for (int i = k + 1; i <= n; i++) {
        const double bin_term = (double)(n - i + 1) / (double)i;
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1.0) {
            const double err = term * ((1.0 - pow(mult_term, (double)(n - i + 1))) / (1.0 - mult_term) - 1.0);
            if (err < tolerance * fabs(-log10(bin_tail) - logNT) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - logNT;
}

struct EdFit {
    float ATA[4], ATV[2];  // the reference's cv::Mat_<float> members
};

CSB_HD void ed_solve(const EdFit& f, double* eq) {
    const double coef = 1.0 / (double(f.ATA[0]) * double(f.ATA[3]) - double(f.ATA[1]) * double(f.ATA[2]));
    eq[0] = coef * (double(f.ATA[3]) * double(f.ATV[0]) - double(f.ATA[1]) * double(f.ATV[1]));
    eq[1] = coef * (double(f.ATA[0]) * double(f.ATV[1]) - double(f.ATA[2]) * double(f.ATV[0]));
}

// sums over n pixels starting at p: a = the running coordinate (x for a horizontal fit), b = the other one
CSB_HD void ed_sums(const ushort2* p, int n, bool horiz, double& sa2, double& sa, double& sab, double& sb) {
    sa2 = 0; sa = 0; sab = 0; sb = 0;
    for (int i = 0; i < n; i++) {
        const double a = (double)(horiz ? p[i].x : p[i].y), b = (double)(horiz ? p[i].y : p[i].x);
        sa2 += a * a; sa += a; sab += a * b; sb += b;
    }
}

// item = chain of a frame
CSB_HD void ed_fit(const EdBuffers& B, const EdDims& d, int frame, int chain_id) {
    const int W = d.w, H = d.h;
    const uint16_t* gd = B.gd + (size_t)frame * W * H;
    const short2* grad = B.grad + (size_t)frame * W * H;
    const int* sid = B.chain_sid + (size_t)frame * (d.max_edges + 2);
    const ushort2* ch = B.chain_px + (size_t)frame * d.chain_cap;
    ushort2* ln = B.line_px + (size_t)frame * d.chain_cap;
    EdLine* stage = B.stage + (size_t)frame * d.stage_cap;
    int S = sid[chain_id];
    const int E = sid[chain_id + 1];
    // Staging slots of this chain: [S / 15, S / 15 + (E - S) / 15).  Every line -- accepted or not -- consumes at least its 15 initial chain
    // pixels, so a chain yields at most (E - S) / 15 lines, and floor((S + len) / 15) >= floor(S / 15) + floor(len / 15) keeps the ranges
    // of consecutive chains disjoint.
    const int slot_base = S / ED_MIN_LINE_LEN, slot_cap = (E - S) / ED_MIN_LINE_LEN;
    for (int k = 0; k < slot_cap; k++) stage[slot_base + k].n_px = 0;
    int n_lines = 0;
    const double logNT = 2.0 * (log10((double)(unsigned)W) + log10((double)(unsigned)H));
    int off = S;                                 // offsetInLineArray, chain-local: line pixels never outnumber the chain pixels consumed
    int newOffsetS = off;
    double lineFitErr = 0;
    double eq[2] = {0, 0};
    EdFit fit{};
    unsigned long long n_staged = 0;
    while (E > S + ED_MIN_LINE_LEN) {
        bool horiz = false;
        while (E > S + ED_MIN_LINE_LEN) {
            // LeastSquaresLineFit_ (:2632-2710) through the first 15 pixels
            horiz = ed_horizontal(gd[ch[S].y * W + ch[S].x]);
            double sa2, sa, sab, sb;
            ed_sums(ch + S, ED_MIN_LINE_LEN, horiz, sa2, sa, sab, sb);
            fit.ATA[0] = (float)sa2; fit.ATA[1] = (float)sa; fit.ATA[2] = (float)sa; fit.ATA[3] = (float)ED_MIN_LINE_LEN;
            fit.ATV[0] = (float)sab; fit.ATV[1] = (float)sb;
            ed_solve(fit, eq);
            double fe = 0;
            for (int i = 0; i < ED_MIN_LINE_LEN; i++) {
                const double xx = (double)ch[S + i].x, yy = (double)ch[S + i].y;
                const double c = horiz ? yy - xx * eq[0] - eq[1] : xx - yy * eq[0] - eq[1];
                fe += c * c;
            }
            lineFitErr = sqrt(fe);
            if (lineFitErr <= ED_FIT_ERR) break;
            S += ED_SKIP;
        }
        if (lineFitErr > ED_FIT_ERR) break;
        const int lineStart = off;
        horiz = ed_horizontal(gd[ch[S].y * W + ch[S].x]);
        double coef1 = 0;
        bool bExtended = true, bFirstTry = true;
        int tryTimes = 0;
        while (bExtended) {
            tryTimes++;
            if (bFirstTry) {
                bFirstTry = false;
                for (int i = 0; i < ED_MIN_LINE_LEN; i++) ln[off++] = ch[S++];
            } else {
                // LeastSquaresLineFit_ (:2712-2790): add the pixels [newOffsetS, off) to the running float sums
                const int length = off - lineStart, newLength = off - newOffsetS;
                if (length > 0 && newLength > 0) {
                    const bool hz = ed_horizontal(gd[ln[lineStart].y * W + ln[lineStart].x]);
                    double sa2, sa, sab, sb;
                    ed_sums(ln + newOffsetS, newLength, hz, sa2, sa, sab, sb);
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
                const double xx = (double)ch[S].x, yy = (double)ch[S].y;
                const double dist = horiz ? fabs(eq[0] * xx - yy + eq[1]) * coef1 : fabs(xx - eq[0] * yy - eq[1]) * coef1;
                ln[off++] = ch[S++];
                if (dist > ED_FIT_ERR) {
                    numOfOutlier++;
                    if (numOfOutlier > 3) break;
                } else {
                    numOfOutlier = 0;
                }
            }
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
            for (int i = 0; i < n; i++) {
                const short2 g = grad[ln[lineStart + i].y * W + ln[lineStart + i].x];
                mgx += g.x; mgy += g.y;
            }
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
                for (int i = 0; i < n; i++) {
                    const short2 g = grad[ln[lineStart + i].y * W + ln[lineStart + i].x];
                    const double pd = det_atan2(-(double)g.x, (double)g.y);
                    const double dis = fabs((double)direction - pd);
                    if (fabs(2 * ED_PI - dis) < 0.392699 || dis < 0.392699) k++;
                }
                ok = ed_nfa(n, k, 0.125, logNT) > 0;
            }
        }
        if (ok) {
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
            n_lines++;
            n_staged++;
        } else {
            off = lineStart;
        }
    }
#if defined(__CUDA_ARCH__)
    if (n_staged) atomicAdd(B.stats + 3, n_staged);
#else
    B.stats[3] += n_staged;
#endif
}

// item = frame
CSB_HD void ed_emit(const EdBuffers& B, const EdDims& d, int frame, int filter, float length_thres, int max_lines) {
    const int nch = B.n_chains[frame];
    const int* sid = B.chain_sid + (size_t)frame * (d.max_edges + 2);
    const EdLine* stage = B.stage + (size_t)frame * d.stage_cap;
    float* out = B.lines + (size_t)frame * max_lines * 4;
    int n_out = 0;
    for (int c = 0; c < nch; c++) {
        const int slot_base = sid[c] / ED_MIN_LINE_LEN, slot_cap = (sid[c + 1] - sid[c]) / ED_MIN_LINE_LEN;   // see ed_fit
        for (int k = 0; k < slot_cap; k++) {
            const EdLine& L = stage[slot_base + k];
            if (L.n_px <= 0) break;
            // OctaveKeyLines :836-851 (length), :1075-1140 (end-point order); scale[0] = 1
            const float dx = fabsf(L.ep[0] - L.ep[2]), dy = fabsf(L.ep[1] - L.ep[3]);
            const float len = sqrtf(dx * dx + dy * dy);
            const float sx = L.ep[0], sy = L.ep[1], ex = L.ep[2], ey = L.ep[3];
            const float ddx = ex - sx, ddy = ey - sy;
            const double dir = (double)L.direction;
            bool change = false;
            if (dir >= -0.75 * ED_PI && dir < -0.25 * ED_PI && ddy > 0) change = true;
            if (dir >= -0.25 * ED_PI && dir < 0.25 * ED_PI && ddx < 0) change = true;
            if (dir >= 0.25 * ED_PI && dir < 0.75 * ED_PI && ddy < 0) change = true;
            if (((dir >= 0.75 * ED_PI && dir < ED_PI) || (dir >= -ED_PI && dir < -0.75 * ED_PI)) && ddx > 0) change = true;
            if (filter && !(len > length_thres)) continue;
            if (n_out < max_lines) {
                float* o = out + 4 * (size_t)n_out;
                if (change) { o[0] = ex; o[1] = ey; o[2] = sx; o[3] = sy; }
                else { o[0] = sx; o[1] = sy; o[2] = ex; o[3] = ey; }
                if (B.keyl) B.keyl[(size_t)frame * max_lines + n_out] = make_float2(L.direction, (float)L.n_px);
            }
            n_out++;
        }
    }
    B.n_lines[frame] = n_out;
}

}  // namespace csb
