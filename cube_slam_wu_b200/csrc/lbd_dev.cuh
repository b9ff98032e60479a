// lbd_dev.cuh -- LBD descriptors for key lines whose fields come from the detector (EDLines: direction = lineDirection_, numOfPixels =
// pixels of the fitted chain segment, end points as projected), as one-thread-per-item device functions.
//
// Reference: BinaryDescriptor::computeLBD (line_lbd/libs/binary_descriptor.cpp:1150-1512) as computeImpl (:607-794) reaches it from
// line_lbd_detect::detect_descrip_lines with use_LSD = false; binaryConversion (:405-417, 766-773).  Oracle: orc_lbd_describe_keylines
// (oracle/oracle_lbd.cpp).  Same arithmetic as k_lbd_describe (lbd.cu), which takes LSD-style key lines and is warp-cooperative; this
// variant trades speed for being executable on the host (tests/emul/edlines_emul.cpp), like edlines_dev.cuh.
//
//   lbdk_row    : item = (line, row of the 63-row support region): the row's four weighted sums -> rs[line][4][64]
//   lbdk_finish : item = line: band sums, mean / std, the two normalisations, 72 floats + 32 bytes
//   lbdk_line   : item = line: the 63 rows one after the other, then lbdk_finish (what the kernel wrapper calls)
#pragma once
#include <cmath>
#include <cstdint>

#include "csb_math.cuh"

namespace csb {

constexpr int LBDK_BANDS = 9, LBDK_BAND_W = 7, LBDK_ROWS = LBDK_BANDS * LBDK_BAND_W;

// same sequence as det_sincos in oracle_lsd.cpp / lsd.cu / lbd.cu
CSB_HD void lbdk_sincos(double x, double& s, double& c) {
    const double two_over_pi = 6.36619772367581382433e-01, pio2_hi = 1.57079632673412561417e+00, pio2_lo = 6.07710050650619224932e-11;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double kf = floor(x * two_over_pi + 0.5);
    const int k = (int)kf;
    double r = x - kf * pio2_hi;
    r = r - kf * pio2_lo;
    const double z = r * r;
    const double sn = r + (z * r) * (S1 + z * (S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)))));
    const double cs = 1.0 - (0.5 * z - z * (z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))))));
    switch (k & 3) {
        case 0: s = sn; c = cs; break;
        case 1: s = cs; c = -sn; break;
        case 2: s = -sn; c = -cs; break;
        default: s = -cs; c = sn; break;
    }
}

// BinaryDescriptor ctor (binary_descriptor.cpp:232-258), integer divisions as written there; host only (glibc exp on both sides)
inline void lbdk_weights(float* G63, float* L21) {
    double u = (LBDK_BAND_W * 3 - 1) / 2;
    double sigma = (LBDK_BAND_W * 2 + 1) / 2;
    double invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < LBDK_BAND_W * 3; i++) {
        const double dis = i - u;
        L21[i] = (float)std::exp(dis * dis * invsigma2);
    }
    u = (LBDK_BANDS * LBDK_BAND_W - 1) / 2;
    sigma = u;
    invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < LBDK_ROWS; i++) {
        const double dis = i - u;
        G63[i] = (float)std::exp(dis * dis * invsigma2);
    }
}

// one row of one line; line = {sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY}; out4 = rs[0..3][row] with stride 64
CSB_HD void lbdk_row(const short2* g, int w, int h, const float* line, float direction, int num_px, int row, float coefG, float* rs) {
    float dL0, dL1;
    {
        double sn, cs;
        lbdk_sincos((double)direction, sn, cs);
        dL0 = (float)cs;
        dL1 = (float)sn;
    }
    const float dO0 = -dL1, dO1 = dL0;
    const int len = (int)(short)num_px;
    const int halfWidth = (len - 1) / 2, halfHeight = (LBDK_ROWS - 1) / 2;
    const float midX = (float)(0.5 * (double)(line[0] + line[2])), midY = (float)(0.5 * (double)(line[1] + line[3]));
    float x = (-dL0 * (float)halfWidth + dL1 * (float)halfHeight) + midX;
    float y = (-dL1 * (float)halfWidth - dL0 * (float)halfHeight) + midY;
    for (int r = 0; r < row; r++) {  // the row origin moves by (-dL1, +dL0) per row, accumulated in float (:1322-1323)
        x = x - dL1;
        y = y + dL0;
    }
    float pL = 0.f, nL = 0.f, pO = 0.f, nO = 0.f;
    const int wm = w - 1, hm = h - 1;
    for (int s = 0; s < len; s++) {
        int xi = (int)(short)(int)roundf(x), yi = (int)(short)(int)roundf(y);
        xi = xi < 0 ? 0 : (xi > wm ? wm : xi);
        yi = yi < 0 ? 0 : (yi > hm ? hm : yi);
        const short2 gg = g[yi * w + xi];
        const float dx = (float)gg.x, dy = (float)gg.y;
        const float gDL = dx * dL0 + dy * dL1;
        const float gDO = dx * dO0 + dy * dO1;
        if (gDL > 0) pL = pL + gDL; else nL = nL - gDL;
        if (gDO > 0) pO = pO + gDO; else nO = nO - gDO;
        x = x + dL0;
        y = y + dL1;
    }
    rs[0 * 64 + row] = coefG * pL;
    rs[1 * 64 + row] = coefG * nL;
    rs[2 * 64 + row] = coefG * pO;
    rs[3 * 64 + row] = coefG * nO;
}

// rs = this line's [4][64] row sums; L = gaussCoefL_ (21 floats); des72 (optional) and desc32 outputs
CSB_HD void lbdk_finish(const float* rs, const float* L, float* des72_out, uint8_t* desc32_out) {
    float des[72];
    for (int a = 0; a < LBDK_BANDS * 4; a++) {
        const int b = a >> 2, q = a & 3;
        const int r_lo = (b - 1) * LBDK_BAND_W < 0 ? 0 : (b - 1) * LBDK_BAND_W;
        const int r_hi = (b + 2) * LBDK_BAND_W > LBDK_ROWS ? LBDK_ROWS : (b + 2) * LBDK_BAND_W;
        float S = 0.f, S2 = 0.f;
        for (int r = r_lo; r < r_hi; r++) {
            const int rb = r / LBDK_BAND_W, j = r - rb * LBDK_BAND_W;
            const float c = (rb == b) ? L[j + LBDK_BAND_W] : (rb == b + 1) ? L[j + 2 * LBDK_BAND_W] : L[j];
            const float v = rs[q * 64 + r];
            S = S + c * v;
            S2 = S2 + (c * c) * (v * v);
        }
        const float invN = (b == 0 || b == LBDK_BANDS - 1) ? (float)(1.0 / (LBDK_BAND_W * 2.0)) : (float)(1.0 / (LBDK_BAND_W * 3.0));
        const float mean = S * invN;
        des[b * 8 + q] = mean;
        des[b * 8 + 4 + q] = sqrtf(S2 * invN - mean * mean);
    }
    float tM = 0.f, tS = 0.f;
    for (int b = 0; b < LBDK_BANDS; b++) {
        for (int i = 0; i < 4; i++) tM = tM + des[b * 8 + i] * des[b * 8 + i];
        for (int i = 4; i < 8; i++) tS = tS + des[b * 8 + i] * des[b * 8 + i];
    }
    tM = 1.f / sqrtf(tM);
    tS = 1.f / sqrtf(tS);
    for (int i = 0; i < 72; i++) {
        float v = des[i] * (((i & 7) < 4) ? tM : tS);
        if (v > 0.4f) v = 0.4f;
        des[i] = v;
    }
    float t2 = 0.f;
    for (int i = 0; i < 72; i++) t2 = t2 + des[i] * des[i];
    t2 = 1.f / sqrtf(t2);
    for (int i = 0; i < 72; i++) {
        des[i] = des[i] * t2;
        if (des72_out) des72_out[i] = des[i];
    }
    const unsigned char comb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
                                       {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};
    for (int c = 0; c < 32; c++) {
        const float* f1 = des + 8 * comb[c][0];
        const float* f2 = des + 8 * comb[c][1];
        unsigned r = 0;
        for (int i = 0; i < 8; i++) r |= (f1[i] > f2[i]) ? (1u << i) : 0u;
        desc32_out[c] = (uint8_t)r;
    }
}

// item = line: all 63 rows, then the descriptor (the kernel wrapper's unit of work; no global row-sum buffer)
CSB_HD void lbdk_line(const short2* g, int w, int h, const float* line, float direction, int num_px, const float* G, const float* L, float* des72_out,
                      uint8_t* desc32_out) {
    float rs[4 * 64];
    for (int r = 0; r < LBDK_ROWS; r++) lbdk_row(g, w, h, line, direction, num_px, r, G[r], rs);
    lbdk_finish(rs, L, des72_out, desc32_out);
}

}  // namespace csb
