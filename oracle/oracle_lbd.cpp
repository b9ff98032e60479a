// oracle_lbd.cpp -- TEST INFRASTRUCTURE ONLY (never linked or called by the product).
//
// CPU restatement of the LBD line descriptor of the reference's line_lbd package, as line_lbd_detect::detect_descrip_lines
// (line_lbd/class/line_lbd_allclass.cpp:239-281) reaches it with use_LSD = true, one octave:
//   keyline fields             LSDDetector::detectImpl            line_lbd/libs/LSDDetector.cpp:80-101 (clamp), :219-250 (fields)
//   BinaryDescriptor::compute  -> computeImpl                     line_lbd/libs/binary_descriptor.cpp:607-794
//   gradient images            computeGaussianPyramid / computeSobel  :347-402   (GaussianBlur 5x5 sigma 1 on the u8 frame, Sobel 3x3 -> int16)
//   band weights               BinaryDescriptor ctor              :217-259   (gaussCoefL_: 21 taps, sigma 7; gaussCoefG_: 63 taps, sigma 31)
//   descriptor                 computeLBD                         :1150-1512 (63 rows x numOfPixels samples, 9 bands x 8 floats, two normalisations, 0.4 clip)
//   binary form                binaryConversion + combinations    :74-107, :405-417, :766-773 (32 bytes)
//
// PARITY UNPINNED: the reference has no golden descriptor vectors, does not compile here (OpenCV, ROS), and python cv2 4.13 ships no
// line_descriptor module.  What IS pinned (tests/test_lbd_oracle.py): the OpenCV arithmetic the path relies on -- cv::GaussianBlur on
// CV_8U (cv2 4.13's bit-exact fixed-point path: taps {14, 62, 104, 62, 14} / 256, BORDER_REFLECT_101, one rounding) and cv::Sobel
// -- against cv2 4.13 fixtures; the descriptor arithmetic itself is checked against an independent numpy restatement.
//
// Specified arithmetic (shared with csrc/lbd.cu): the reference takes `atan2` / `cos` / `sin` of float operands through libm and stores
// float results; here they are det_atan2 / det_sincos evaluated in double and rounded once to float (libm_trig = 1 selects libm).
// Everything else is float arithmetic in the reference's order, no FMA contraction (-ffp-contract=off).
//
// Undefined in the reference and not restated: line_lbd_detect::get_line_descriptors (mat_to_keylines, line_lbd_allclass.cpp:70-112,
// reads KeyLine fields before they are set).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle_math.h"

extern "C" void orc_lsd_det_sincos(double x, double* s, double* c);  // oracle_lsd.cpp

namespace {

constexpr int NUM_OF_BANDS = 9, WIDTH_OF_BAND = 7, HEIGHT_OF_LSP = NUM_OF_BANDS * WIDTH_OF_BAND;

// binary_descriptor.cpp:74-107
const int COMB[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
                         {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

inline int reflect101(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    if (p < 0) p = 0;
    if (p >= len) p = len - 1;
    return p;
}

// cv::GaussianBlur(src, dst, Size(5, 5), 1) on CV_8UC1 (binary_descriptor.cpp:356, :815-816).  OpenCV is not part of the reference tree and
// its version is not pinned by the reference's CMake files; the 8-bit Gaussian exists in two generations that differ in the integer taps:
//   generation 4 (default here; cv2 4.13's bit-exact fixed-point path, pinned by tests/golden/lbd_cv2.npz): {14, 62, 104, 62, 14} / 256
//   generation 3 (OpenCV <= 3.4.0: sepFilter2D with 8 fractional bits, every tap rounded on its own):   {14, 63, 103, 63, 14} / 256
// Both: exact integer row sums, exact column sums, one rounding (v + 2^15) >> 16, saturated to 255 (generation 3's taps sum to 257).
// Generation 3 is what the reference's author ran: with it -- and only with it -- the replay of object_slam's online mode reproduces ALL 58
// rows of the committed output_obj_poses.txt to the printed digits (tests/test_reference_replay.py); with generation 4 the landmark history
// leaves the committed one at frame 28.
int g_blur_generation = 4;
void blur5(const uint8_t* g, int w, int h, std::vector<uint8_t>& out) {
    static const int K4[5] = {14, 62, 104, 62, 14}, K3[5] = {14, 63, 103, 63, 14};
    const int* K = g_blur_generation == 3 ? K3 : K4;
    std::vector<int> hz((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int t = 0; t < 5; t++) s += K[t] * g[(size_t)y * w + reflect101(x + t - 2, w)];
            hz[(size_t)y * w + x] = s;
        }
    out.resize((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int t = 0; t < 5; t++) s += K[t] * hz[(size_t)reflect101(y + t - 2, h) * w + x];
            const int v = (s + 32768) >> 16;
            out[(size_t)y * w + x] = (uint8_t)(v > 255 ? 255 : v);
        }
}

// cv::Sobel(img, d, CV_16SC1, 1, 0, 3) / (0, 1, 3), BORDER_REFLECT_101 (binary_descriptor.cpp:395-396)
void sobel3(const std::vector<uint8_t>& b, int w, int h, int16_t* dx, int16_t* dy) {
    for (int y = 0; y < h; y++) {
        const int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
        for (int x = 0; x < w; x++) {
            const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            const int a = b[(size_t)ym * w + xm], bb = b[(size_t)ym * w + x], c = b[(size_t)ym * w + xp];
            const int d = b[(size_t)y * w + xm], f = b[(size_t)y * w + xp];
            const int q = b[(size_t)yp * w + xm], hh = b[(size_t)yp * w + x], i = b[(size_t)yp * w + xp];
            dx[(size_t)y * w + x] = (int16_t)((c + 2 * f + i) - (a + 2 * d + q));
            dy[(size_t)y * w + x] = (int16_t)((q + 2 * hh + i) - (a + 2 * bb + c));
        }
    }
}

struct KeyLine {
    float sx, sy, ex, ey, angle, length;
    int num_px;
};

// LSDDetector.cpp:80-101 (checkLineExtremes) + :219-250 with octaveScale = 1
KeyLine make_keyline(const float* l, int w, int h, int libm_trig) {
    float e[4] = {l[0], l[1], l[2], l[3]};
    if (e[0] < 0) e[0] = 0;
    if (e[0] >= w) e[0] = (float)w - 1.0f;
    if (e[2] < 0) e[2] = 0;
    if (e[2] >= w) e[2] = (float)w - 1.0f;
    if (e[1] < 0) e[1] = 0;
    if (e[1] >= h) e[1] = (float)h - 1.0f;
    if (e[3] < 0) e[3] = 0;
    if (e[3] >= h) e[3] = (float)h - 1.0f;
    KeyLine k;
    k.sx = e[0]; k.sy = e[1]; k.ex = e[2]; k.ey = e[3];
    const double ddx = double(e[0] - e[2]), ddy = double(e[1] - e[3]);
    k.length = (float)std::sqrt(ddx * ddx + ddy * ddy);  // :239
    // cv::LineIterator(img, Point2f, Point2f): Point2f -> Point by cvRound (round half to even), 8-connected: max(|dx|, |dy|) + 1 pixels;
    // the clamped end points are inside the image, so clipLine changes nothing
    const int x1 = (int)std::nearbyintf(e[0]), y1 = (int)std::nearbyintf(e[1]), x2 = (int)std::nearbyintf(e[2]), y2 = (int)std::nearbyintf(e[3]);
    const int adx = x2 > x1 ? x2 - x1 : x1 - x2, ady = y2 > y1 ? y2 - y1 : y1 - y2;
    k.num_px = (adx > ady ? adx : ady) + 1;
    const float fy = e[3] - e[1], fx = e[2] - e[0];  // :245: atan2(endPointY - startPointY, endPointX - startPointX), float operands
    k.angle = libm_trig ? (float)std::atan2((double)fy, (double)fx) : (float)orc::det_atan2((double)fy, (double)fx);
    return k;
}

struct Weights {
    float G[HEIGHT_OF_LSP], L[3 * WIDTH_OF_BAND];
    Weights() {  // binary_descriptor.cpp:232-258 (integer divisions as written there)
        double u = (WIDTH_OF_BAND * 3 - 1) / 2;
        double sigma = (WIDTH_OF_BAND * 2 + 1) / 2;
        double invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < WIDTH_OF_BAND * 3; i++) {
            const double dis = i - u;
            L[i] = (float)std::exp(dis * dis * invsigma2);
        }
        u = (NUM_OF_BANDS * WIDTH_OF_BAND - 1) / 2;
        sigma = u;
        invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < HEIGHT_OF_LSP; i++) {
            const double dis = i - u;
            G[i] = (float)std::exp(dis * dis * invsigma2);
        }
    }
};

// computeLBD (binary_descriptor.cpp:1150-1512) for one line; des = 72 floats
void compute_lbd(const KeyLine& kl, const int16_t* pdx, const int16_t* pdy, int w, int h, int libm_trig, const Weights& wt, float* des) {
    const short heightOfLSP = HEIGHT_OF_LSP;
    float pL[NUM_OF_BANDS] = {0}, nL[NUM_OF_BANDS] = {0}, pL2[NUM_OF_BANDS] = {0}, nL2[NUM_OF_BANDS] = {0};
    float pO[NUM_OF_BANDS] = {0}, nO[NUM_OF_BANDS] = {0}, pO2[NUM_OF_BANDS] = {0}, nO2[NUM_OF_BANDS] = {0};
    const short halfHeight = (heightOfLSP - 1) / 2;
    const short realWidth = (short)w, imageWidth = (short)(w - 1), imageHeight = (short)(h - 1);
    const short lengthOfLSP = (short)kl.num_px;
    const short halfWidth = (lengthOfLSP - 1) / 2;
    const float midX = (float)(0.5 * (kl.sx + kl.ex)), midY = (float)(0.5 * (kl.sy + kl.ey));
    float dL[2], dO[2];
    if (libm_trig) {
        dL[0] = (float)std::cos((double)kl.angle);
        dL[1] = (float)std::sin((double)kl.angle);
    } else {
        double s, c;
        orc_lsd_det_sincos((double)kl.angle, &s, &c);
        dL[0] = (float)c;
        dL[1] = (float)s;
    }
    dO[0] = -dL[1];
    dO[1] = dL[0];
    float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + midX;
    float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + midY;
    for (short hID = 0; hID < heightOfLSP; hID++) {
        float sCorX = sCorX0, sCorY = sCorY0;
        float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
        for (short wID = 0; wID < lengthOfLSP; wID++) {
            short tempCor = (short)std::round(sCorX);
            const short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
            tempCor = (short)std::round(sCorY);
            const short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
            const short dx = pdx[yCor * realWidth + xCor], dy = pdy[yCor * realWidth + xCor];
            const float gDL = dx * dL[0] + dy * dL[1];
            const float gDO = dx * dO[0] + dy * dO[1];
            if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
            if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
            sCorX += dL[0];
            sCorY += dL[1];
        }
        sCorX0 -= dL[1];
        sCorY0 += dL[0];
        float coef = wt.G[hID];
        pgdLRowSum = coef * pgdLRowSum;
        ngdLRowSum = coef * ngdLRowSum;
        const float pgdL2RowSum = pgdLRowSum * pgdLRowSum, ngdL2RowSum = ngdLRowSum * ngdLRowSum;
        pgdORowSum = coef * pgdORowSum;
        ngdORowSum = coef * ngdORowSum;
        const float pgdO2RowSum = pgdORowSum * pgdORowSum, ngdO2RowSum = ngdORowSum * ngdORowSum;
        auto add = [&](int band, float c) {
            pL[band] += c * pgdLRowSum;
            nL[band] += c * ngdLRowSum;
            pL2[band] += c * c * pgdL2RowSum;
            nL2[band] += c * c * ngdL2RowSum;
            pO[band] += c * pgdORowSum;
            nO[band] += c * ngdORowSum;
            pO2[band] += c * c * pgdO2RowSum;
            nO2[band] += c * c * ngdO2RowSum;
        };
        short bandID = (short)(hID / WIDTH_OF_BAND);
        add(bandID, wt.L[hID % WIDTH_OF_BAND + WIDTH_OF_BAND]);
        bandID--;
        if (bandID >= 0) add(bandID, wt.L[hID % WIDTH_OF_BAND + 2 * WIDTH_OF_BAND]);
        bandID = bandID + 2;
        if (bandID < NUM_OF_BANDS) add(bandID, wt.L[hID % WIDTH_OF_BAND]);
    }
    const float invN2 = (float)(1.0 / (WIDTH_OF_BAND * 2.0)), invN3 = (float)(1.0 / (WIDTH_OF_BAND * 3.0));
    for (int b = 0; b < NUM_OF_BANDS; b++) {
        const float invN = (b == 0 || b == NUM_OF_BANDS - 1) ? invN2 : invN3;
        float* d = des + b * 8;
        float temp = pL[b] * invN;
        d[0] = temp;
        d[4] = std::sqrt(pL2[b] * invN - temp * temp);
        temp = nL[b] * invN;
        d[1] = temp;
        d[5] = std::sqrt(nL2[b] * invN - temp * temp);
        temp = pO[b] * invN;
        d[2] = temp;
        d[6] = std::sqrt(pO2[b] * invN - temp * temp);
        temp = nO[b] * invN;
        d[3] = temp;
        d[7] = std::sqrt(nO2[b] * invN - temp * temp);
    }
    float tempM = 0, tempS = 0;
    for (int b = 0; b < NUM_OF_BANDS; b++) {
        const float* d = des + b * 8;
        tempM += d[0] * d[0];
        tempM += d[1] * d[1];
        tempM += d[2] * d[2];
        tempM += d[3] * d[3];
        tempS += d[4] * d[4];
        tempS += d[5] * d[5];
        tempS += d[6] * d[6];
        tempS += d[7] * d[7];
    }
    tempM = 1 / std::sqrt(tempM);
    tempS = 1 / std::sqrt(tempS);
    for (int b = 0; b < NUM_OF_BANDS; b++) {
        float* d = des + b * 8;
        for (int i = 0; i < 4; i++) d[i] = d[i] * tempM;
        for (int i = 4; i < 8; i++) d[i] = d[i] * tempS;
    }
    for (int i = 0; i < NUM_OF_BANDS * 8; i++)
        if (des[i] > 0.4) des[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < NUM_OF_BANDS * 8; i++) temp += des[i] * des[i];
    temp = 1 / std::sqrt(temp);
    for (int i = 0; i < NUM_OF_BANDS * 8; i++) des[i] = des[i] * temp;
}

}  // namespace

extern "C" {

// 4 (default): cv2 4.x taps; 3: the taps of OpenCV <= 3.4.0, the generation the reference's committed outputs were produced with
void orc_lbd_set_blur_generation(int gen) { g_blur_generation = gen == 3 ? 3 : 4; }

// blurred frame (optional) and the two int16 gradient images the descriptor samples
void orc_lbd_gradients(const uint8_t* gray, int w, int h, uint8_t* blur_out, int16_t* dx_out, int16_t* dy_out) {
    std::vector<uint8_t> b;
    blur5(gray, w, h, b);
    if (blur_out) std::memcpy(blur_out, b.data(), b.size());
    sobel3(b, w, h, dx_out, dy_out);
}

void orc_lbd_weights(float* g63, float* l21) {
    const Weights wt;
    std::memcpy(g63, wt.G, sizeof wt.G);
    std::memcpy(l21, wt.L, sizeof wt.L);
}

// lines: n x 4 float [x1 y1 x2 y2] in frame coordinates (the LSD key lines of the frame).  desc72_out: n x 72 float (optional),
// desc32_out: n x 32 bytes (optional), keyline_out: n x 3 float {angle, numOfPixels, lineLength} (optional).  Returns n.
int orc_lbd_describe(const uint8_t* gray, int w, int h, const float* lines, int n, int libm_trig, float* desc72_out, uint8_t* desc32_out,
                     float* keyline_out) {
    if (n <= 0) return 0;  // computeImpl: "keypoint list is empty" -> descriptors untouched (binary_descriptor.cpp:623-628)
    std::vector<int16_t> dx((size_t)w * h), dy((size_t)w * h);
    orc_lbd_gradients(gray, w, h, nullptr, dx.data(), dy.data());
    const Weights wt;
    for (int i = 0; i < n; i++) {
        const KeyLine kl = make_keyline(lines + 4 * (size_t)i, w, h, libm_trig);
        float des[72];
        compute_lbd(kl, dx.data(), dy.data(), w, h, libm_trig, wt, des);
        if (desc72_out) std::memcpy(desc72_out + 72 * (size_t)i, des, sizeof des);
        if (desc32_out)
            for (int c = 0; c < 32; c++) {  // binaryConversion (binary_descriptor.cpp:405-417)
                const float *f1 = des + 8 * COMB[c][0], *f2 = des + 8 * COMB[c][1];
                unsigned r = 0;
                for (int k = 0; k < 8; k++)
                    if (f1[k] > f2[k]) r += 1u << k;
                desc32_out[32 * (size_t)i + c] = (uint8_t)r;
            }
        if (keyline_out) {
            keyline_out[3 * (size_t)i] = kl.angle;
            keyline_out[3 * (size_t)i + 1] = (float)kl.num_px;
            keyline_out[3 * (size_t)i + 2] = kl.length;
        }
    }
    return n;
}

// Descriptors for key lines whose fields come from a detector other than LSDDetector (EDLines: OctaveKeyLines fills direction =
// lineDirection_, numOfPixels = the pixel count of the fitted chain segment, and the end points as projected, unclamped;
// binary_descriptor.cpp:1045-1140).  fields: n x 2 float {direction, numOfPixels}.
int orc_lbd_describe_keylines(const uint8_t* gray, int w, int h, const float* lines, const float* fields, int n, int libm_trig, float* desc72_out,
                              uint8_t* desc32_out) {
    if (n <= 0) return 0;
    std::vector<int16_t> dx((size_t)w * h), dy((size_t)w * h);
    orc_lbd_gradients(gray, w, h, nullptr, dx.data(), dy.data());
    const Weights wt;
    for (int i = 0; i < n; i++) {
        KeyLine kl;
        kl.sx = lines[4 * (size_t)i]; kl.sy = lines[4 * (size_t)i + 1]; kl.ex = lines[4 * (size_t)i + 2]; kl.ey = lines[4 * (size_t)i + 3];
        kl.angle = fields[2 * (size_t)i];
        kl.num_px = (int)fields[2 * (size_t)i + 1];
        kl.length = 0;
        float des[72];
        compute_lbd(kl, dx.data(), dy.data(), w, h, libm_trig, wt, des);
        if (desc72_out) std::memcpy(desc72_out + 72 * (size_t)i, des, sizeof des);
        if (desc32_out)
            for (int c = 0; c < 32; c++) {
                const float *f1 = des + 8 * COMB[c][0], *f2 = des + 8 * COMB[c][1];
                unsigned r = 0;
                for (int k = 0; k < 8; k++)
                    if (f1[k] > f2[k]) r += 1u << k;
                desc32_out[32 * (size_t)i + c] = (uint8_t)r;
            }
    }
    return n;
}

}  // extern "C"
