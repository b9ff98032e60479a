// oracle_edlines.cpp -- TEST INFRASTRUCTURE ONLY (never linked or called by the product).
//
// CPU restatement of the EDLines branch of the reference's line_lbd package (use_LSD = false: what object_slam selects,
// object_slam/src/main_obj.cpp:503-505), one octave:
//   line_lbd_detect::detect_raw_lines / filter_lines        line_lbd/class/line_lbd_allclass.cpp:130-149, 200-208
//   BinaryDescriptor::detectImpl                            line_lbd/libs/binary_descriptor.cpp:486-590
//   BinaryDescriptor::OctaveKeyLines                        :796-1148  (GaussianBlur 5x5 sigma 1, EDline, end-point order by direction)
//   EDLineDetector::EdgeDrawing                             :1583-2380 (Sobel, gradient / direction maps, anchors, smart routing, chains)
//   EDLineDetector::EDline                                  :2383-2630 (initial least-squares segment, extension, end points)
//   EDLineDetector::LeastSquaresLineFit_ (both overloads)   :2632-2790
//   EDLineDetector::LineValidation_ + nfa                   :2793-2873, line_descriptor/descriptor.hpp:695-848
// Parameters as the reference's constructor sets them (:1515-1525): gradient threshold 80, anchor threshold 8, scan interval 2,
// minimum line length 15, line-fit error 1.6.
//
// PARITY STATUS: PINNED through the reference's own committed outputs.  The reference ships no EDLines segment file, but its object_slam
// node ran this detector (main_obj.cpp:504) to produce output_obj_poses.txt / output_cam_poses.txt, and tests/test_reference_replay.py
// reproduces all 58 rows to their printed digits with this oracle (and the OpenCV <= 3.4.0 blur taps, oracle_lbd.cpp) feeding the cuboid proposals -- with the LSD
// branch instead, the landmark history leaves the committed one at the second frame: the best proposal of a frame depends on the exact
// line table.  The OpenCV arithmetic underneath (blur, Sobel) is the same cv2-pinned code as oracle_lbd.cpp (orc_lbd_gradients).
//
// Quirks of the reference that are kept because they decide which pixels join a chain:
//   * the routing compares neighbours through `(unsigned char) pgImg[...]` (:1746-1748 ...): gradient values above 255 would wrap
//     (after the 5x5 blur (|dx| + |dy|) / 4 stays below 256 on every frame tried, so the cast is kept but never bites);
//   * `gImg_ = gImg_ / 4` is a cv::MatExpr on CV_16S: scale by 0.25 and round half to even (not an integer division);
//   * anchors are scanned column by column (:1643-1670);
//   * the least-squares sums live in float matrices (cv::Mat_<float>, descriptor.hpp:629-639): every sum is rounded to float once
//     (cv::gemm accumulates float products in double), the running ATA / ATV additions are float additions.
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle_math.h"

extern "C" void orc_lbd_gradients(const uint8_t* gray, int w, int h, uint8_t* blur_out, int16_t* dx_out, int16_t* dy_out);  // oracle_lbd.cpp

namespace {

constexpr int HORIZONTAL = 255, VERTICAL = 0;
constexpr int UP = 1, RIGHT = 2, DOWN = 3, LEFT = 4;
constexpr int TRY_TIME = 6, SKIP_EDGE_POINT = 2;
constexpr short GRADIENT_THRESHOLD = 80;
constexpr int ANCHOR_THRESHOLD = 8, SCAN_INTERVALS = 2, MIN_LINE_LEN = 15;
constexpr double LINE_FIT_ERR_THRESHOLD = 1.6;
constexpr double MLN10 = 2.30258509299404568402;

// Specified arithmetic (shared with csrc/edlines_dev.cuh): LineValidation_ calls libm atan2 per chain pixel and for the line direction; both
// sides use det_atan2 (oracle_math.h / csb_math.cuh) unless g_libm_trig selects the literal libm call.
int g_libm_trig = 0;
inline double sp_atan2(double y, double x) { return g_libm_trig ? std::atan2(y, x) : orc::det_atan2(y, x); }

struct Maps {
    int w = 0, h = 0;
    std::vector<int16_t> dx, dy, g;   // g = thresholded (|dx| + |dy|) / 4
    std::vector<uint8_t> dir;         // 255: |dx| < |dy| (horizontal edge), 0 otherwise
};

// binary_descriptor.cpp:1624-1637
void build_maps(const uint8_t* gray, int w, int h, Maps& m) {
    m.w = w; m.h = h;
    const size_t n = (size_t)w * h;
    m.dx.resize(n); m.dy.resize(n); m.g.resize(n); m.dir.resize(n);
    orc_lbd_gradients(gray, w, h, nullptr, m.dx.data(), m.dy.data());  // GaussianBlur 5x5 sigma 1 (:815-816, increaseSigma = 1), Sobel 3x3
    for (size_t i = 0; i < n; i++) {
        const int ax = std::abs((int)m.dx[i]), ay = std::abs((int)m.dy[i]);
        const int sum = ax + ay;
        const int t = sum > GRADIENT_THRESHOLD + 1 ? sum : 0;           // cv::threshold(..., THRESH_TOZERO)
        m.g[i] = (int16_t)std::nearbyint(t * 0.25);                     // MatExpr `/ 4`: saturate_cast<short>(x * 0.25), round half to even
        m.dir[i] = ax < ay ? HORIZONTAL : VERTICAL;                     // cv::compare(dxABS, dyABS, CMP_LT)
    }
}

struct Chains {
    std::vector<unsigned> x, y, sid;  // sid[k] .. sid[k+1]: pixels of chain k
    unsigned n = 0;
};

// EdgeDrawing :1638-2380.  Returns 1, or -1 on the reference's capacity errors.
int edge_drawing(const Maps& m, Chains& out, std::vector<unsigned>* anchors_xy) {
    const unsigned W = m.w, H = m.h;
    const unsigned pixelNum = W * H, edgePixelArraySize = pixelNum / 5, maxNumOfEdge = edgePixelArraySize / 20;
    const int16_t* g = m.g.data();
    const uint8_t* dir = m.dir.data();
    std::vector<unsigned> ax, ay;
    for (unsigned w = 1; w + 1 < W; w += SCAN_INTERVALS)
        for (unsigned h = 1; h + 1 < H; h += SCAN_INTERVALS) {
            const int i = h * W + w;
            if (dir[i] == HORIZONTAL) {
                if (g[i] >= g[i - (int)W] + ANCHOR_THRESHOLD && g[i] >= g[i + (int)W] + ANCHOR_THRESHOLD) { ax.push_back(w); ay.push_back(h); }
            } else {
                if (g[i] >= g[i - 1] + ANCHOR_THRESHOLD && g[i] >= g[i + 1] + ANCHOR_THRESHOLD) { ax.push_back(w); ay.push_back(h); }
            }
        }
    if (ax.size() > edgePixelArraySize) return -1;
    if (anchors_xy) {
        anchors_xy->clear();
        for (size_t i = 0; i < ax.size(); i++) { anchors_xy->push_back(ax[i]); anchors_xy->push_back(ay[i]); }
    }
    std::vector<uint8_t> edge((size_t)pixelNum, 0);
    std::vector<unsigned> fx, fy, sx, sy, fs, ss;  // first / second parts and their start offsets
    unsigned lastX = 0, lastY = 0;                 // carried from walk to walk, as in the reference

    // one walk of the smart routing (the reference repeats this body four times: :1716-1850, :1858-1990, :2000-2160, :2170-2312)
    auto walk = [&](unsigned x, unsigned y, int lastDirection, std::vector<unsigned>& px, std::vector<unsigned>& py) {
        int idx = y * W + x;
        while (g[idx] > 0 && !edge[idx]) {
            edge[idx] = 1;
            px.push_back(x);
            py.push_back(y);
            int shouldGo = 0;
            auto gv = [&](int off) { return (unsigned char)g[idx + off]; };
            if (dir[idx] == HORIZONTAL) {
                if (lastDirection == UP || lastDirection == DOWN) shouldGo = (x > lastX) ? RIGHT : LEFT;
                lastX = x; lastY = y;
                if (lastDirection == RIGHT || shouldGo == RIGHT) {
                    if (x == W - 1 || y == 0 || y == H - 1) break;
                    const unsigned char g1 = gv(-(int)W + 1), g2 = gv(1), g3 = gv((int)W + 1);
                    if (g1 >= g2 && g1 >= g3) { x = x + 1; y = y - 1; }
                    else if (g3 >= g2 && g3 >= g1) { x = x + 1; y = y + 1; }
                    else { x = x + 1; }
                    lastDirection = RIGHT;
                } else if (lastDirection == LEFT || shouldGo == LEFT) {
                    if (x == 0 || y == 0 || y == H - 1) break;
                    const unsigned char g1 = gv(-(int)W - 1), g2 = gv(-1), g3 = gv((int)W - 1);
                    if (g1 >= g2 && g1 >= g3) { x = x - 1; y = y - 1; }
                    else if (g3 >= g2 && g3 >= g1) { x = x - 1; y = y + 1; }
                    else { x = x - 1; }
                    lastDirection = LEFT;
                }
            } else {
                if (lastDirection == RIGHT || lastDirection == LEFT) shouldGo = (y > lastY) ? DOWN : UP;
                lastX = x; lastY = y;
                if (lastDirection == DOWN || shouldGo == DOWN) {
                    if (x == 0 || x == W - 1 || y == H - 1) break;
                    const unsigned char g1 = gv((int)W + 1), g2 = gv((int)W), g3 = gv((int)W - 1);
                    if (g1 >= g2 && g1 >= g3) { x = x + 1; y = y + 1; }
                    else if (g3 >= g2 && g3 >= g1) { x = x - 1; y = y + 1; }
                    else { y = y + 1; }
                    lastDirection = DOWN;
                } else if (lastDirection == UP || shouldGo == UP) {
                    if (x == 0 || x == W - 1 || y == 0) break;
                    const unsigned char g1 = gv(-(int)W + 1), g2 = gv(-(int)W), g3 = gv(-(int)W - 1);
                    if (g1 >= g2 && g1 >= g3) { x = x + 1; y = y - 1; }
                    else if (g3 >= g2 && g3 >= g1) { x = x - 1; y = y - 1; }
                    else { y = y - 1; }
                    lastDirection = UP;
                }
            }
            idx = y * W + x;
        }
    };

    for (size_t a = 0; a < ax.size(); a++) {
        const unsigned x = ax[a], y = ay[a];
        const int idx = y * W + x;
        if (edge[idx]) continue;
        const size_t f0 = fx.size(), s0 = sx.size();
        const bool horiz = dir[idx] == HORIZONTAL;
        walk(x, y, horiz ? RIGHT : DOWN, fx, fy);
        edge[idx] = 0;  // the anchor starts the second part too
        walk(x, y, horiz ? LEFT : UP, sx, sy);
        const int lenF = (int)(fx.size() - f0), lenS = (int)(sx.size() - s0);
        if (lenF + lenS < MIN_LINE_LEN + 1) {  // short edge: its pixels stay marked, the chain is dropped
            fx.resize(f0); fy.resize(f0); sx.resize(s0); sy.resize(s0);
        } else {
            fs.push_back((unsigned)f0);
            ss.push_back((unsigned)s0);
        }
    }
    if (fs.size() > maxNumOfEdge) return -1;
    if (fx.size() > edgePixelArraySize || sx.size() > edgePixelArraySize) return -1;
    fs.push_back((unsigned)fx.size());
    ss.push_back((unsigned)sx.size());
    out.x.clear(); out.y.clear(); out.sid.clear();
    const size_t nE = fs.size() - 1;
    for (size_t e = 0; e < nE; e++) {
        out.sid.push_back((unsigned)out.x.size());
        for (int t = (int)fs[e + 1] - 1; t >= (int)fs[e]; t--) { out.x.push_back(fx[t]); out.y.push_back(fy[t]); }   // first part, reversed
        for (int t = (int)ss[e] + 1; t < (int)ss[e + 1]; t++) { out.x.push_back(sx[t]); out.y.push_back(sy[t]); }     // second part without the anchor
    }
    out.sid.push_back((unsigned)out.x.size());
    out.n = (unsigned)nE;
    return 1;
}

// descriptor.hpp:655-671, 695-848
bool double_equal(double a, double b) {
    if (a == b) return true;
    const double abs_diff = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
    double abs_max = aa > bb ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
double log_gamma_lanczos(double x) {
    static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
    double b = 0.0;
    for (int n = 0; n < 7; n++) {
        a -= std::log(x + (double)n);
        b += q[n] * std::pow(x, (double)n);
    }
    return a + std::log(b);
}
double log_gamma_windschitl(double x) {
    return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
}
double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }
double nfa(int n, int k, double p, double logNT) {
    const double tolerance = 0.1;
    if (n == 0 || k == 0) return -logNT;
    if (n == k) return -logNT - (double)n * std::log10(p);
    const double p_term = p / (1.0 - p);
    const double log1term = log_gamma((double)n + 1.0) - log_gamma((double)k + 1.0) - log_gamma((double)(n - k) + 1.0) + (double)k * std::log(p) +
                            (double)(n - k) * std::log(1.0 - p);
    double term = std::exp(log1term);
    if (double_equal(term, 0.0)) {
        if ((double)k > (double)n * p) return -log1term / MLN10 - logNT;
        return -logNT;
    }
    double bin_tail = term;
    for (int i = k + 1; i <= n; i++) {
        const double bin_term = (double)(n - i + 1) / (double)i;
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1.0) {
            const double err = term * ((1.0 - std::pow(mult_term, (double)(n - i + 1))) / (1.0 - mult_term) - 1.0);
            if (err < tolerance * std::fabs(-std::log10(bin_tail) - logNT) * bin_tail) break;
        }
    }
    return -std::log10(bin_tail) - logNT;
}

struct Line {
    float ep[4];       // lineEndpoints_
    float direction;   // lineDirection_
    double eq[3];      // lineEquations_
    unsigned s, e;     // pixels [s, e) in the line arrays
};

struct Detector {
    const Maps& m;
    float ATA[4] = {0, 0, 0, 0}, ATV[2] = {0, 0};  // cv::Mat_<float> members: state carried from the initial fit to the extensions
    double logNT = 0;
    explicit Detector(const Maps& mm) : m(mm) {}

    // :2632-2710: fit through the first MIN_LINE_LEN pixels of the chain starting at offsetS; returns the fit error
    double fit_initial(const unsigned* xs, const unsigned* ys, unsigned offsetS, double* eq) {
        const bool horiz = m.dir[ys[offsetS] * m.w + xs[offsetS]] == HORIZONTAL;
        double sa2 = 0, sa = 0, sab = 0, sb = 0;  // exact in double: integer coordinates
        for (int i = 0; i < MIN_LINE_LEN; i++) {
            const double a = (double)(float)(horiz ? xs[offsetS + i] : ys[offsetS + i]), b = (double)(float)(horiz ? ys[offsetS + i] : xs[offsetS + i]);
            sa2 += a * a; sa += a; sab += a * b; sb += b;
        }
        ATA[0] = (float)sa2; ATA[1] = (float)sa; ATA[2] = (float)sa; ATA[3] = (float)(double)MIN_LINE_LEN;
        ATV[0] = (float)sab; ATV[1] = (float)sb;
        solve(eq);
        double fitError = 0;
        for (int i = 0; i < MIN_LINE_LEN; i++) {
            const unsigned o = offsetS + i;
            const double c = horiz ? double(ys[o]) - double(xs[o]) * eq[0] - eq[1] : double(xs[o]) - double(ys[o]) * eq[0] - eq[1];
            fitError += c * c;
        }
        return std::sqrt(fitError);
    }
    // :2712-2790: add the pixels [newOffsetS, offsetE) of the line to the running sums and solve again
    void fit_extend(const unsigned* xs, const unsigned* ys, unsigned offsetS, unsigned newOffsetS, unsigned offsetE, double* eq) {
        const int length = (int)(offsetE - offsetS), newLength = (int)(offsetE - newOffsetS);
        if (length <= 0 || newLength <= 0) return;  // the reference prints an error and returns -1 without touching the equation
        const bool horiz = m.dir[ys[offsetS] * m.w + xs[offsetS]] == HORIZONTAL;
        double sa2 = 0, sa = 0, sab = 0, sb = 0;
        for (int i = 0; i < newLength; i++) {
            const double a = (double)(float)(horiz ? xs[newOffsetS + i] : ys[newOffsetS + i]), b = (double)(float)(horiz ? ys[newOffsetS + i] : xs[newOffsetS + i]);
            sa2 += a * a; sa += a; sab += a * b; sb += b;
        }
        const float t[4] = {(float)sa2, (float)sa, (float)sa, (float)(double)newLength}, tv[2] = {(float)sab, (float)sb};
        for (int i = 0; i < 4; i++) ATA[i] = ATA[i] + t[i];
        for (int i = 0; i < 2; i++) ATV[i] = ATV[i] + tv[i];
        solve(eq);
    }
    void solve(double* eq) const {
        const double coef = 1.0 / (double(ATA[0]) * double(ATA[3]) - double(ATA[1]) * double(ATA[2]));
        eq[0] = coef * (double(ATA[3]) * double(ATV[0]) - double(ATA[1]) * double(ATV[1]));
        eq[1] = coef * (double(ATA[0]) * double(ATV[1]) - double(ATA[2]) * double(ATV[0]));
    }

    // :2793-2873
    bool validate(const unsigned* xs, const unsigned* ys, unsigned offsetS, unsigned offsetE, const double* lineEqu, float& direction) const {
        const int n = (int)(offsetE - offsetS);
        int meanGX = 0, meanGY = 0;
        std::vector<double> pointDirection;
        pointDirection.reserve(n);
        for (int i = 0; i < n; i++) {
            const int index = ys[offsetS + i] * m.w + xs[offsetS + i];
            meanGX += m.dx[index];
            meanGY += m.dy[index];
            pointDirection.push_back(sp_atan2(-(double)m.dx[index], (double)m.dy[index]));
        }
        const double dx = std::fabs(lineEqu[1]), dy = std::fabs(lineEqu[0]);
        if (meanGX == 0 && meanGY == 0) return false;
        if (meanGX > 0 && meanGY >= 0) direction = (float)sp_atan2(-dy, dx);
        if (meanGX <= 0 && meanGY > 0) direction = (float)sp_atan2(dy, dx);
        if (meanGX < 0 && meanGY <= 0) direction = (float)sp_atan2(dy, -dx);
        if (meanGX >= 0 && meanGY < 0) direction = (float)sp_atan2(-dy, -dx);
        if (std::fabs(direction) < 0.15 || M_PI - std::fabs(direction) < 0.15) {
            if (std::fabs(lineEqu[2]) < 10 || std::fabs((double)(unsigned)m.h - std::fabs(lineEqu[2])) < 10) return false;
        }
        if (std::fabs(std::fabs(direction) - M_PI * 0.5) < 0.15) {
            if (std::fabs(lineEqu[2]) < 10 || std::fabs((double)(unsigned)m.w - std::fabs(lineEqu[2])) < 10) return false;
        }
        int k = 0;
        for (int i = 0; i < n; i++) {
            const double dis = std::fabs(direction - pointDirection[i]);
            if (std::fabs(2 * M_PI - dis) < 0.392699 || dis < 0.392699) k++;
        }
        return nfa(n, k, 0.125, logNT) > 0;
    }

    // EDline :2383-2630
    void run(const Chains& ch, std::vector<Line>& lines) {
        lines.clear();
        if (ch.n == 0) return;
        logNT = 2.0 * (std::log10((double)(unsigned)m.w) + std::log10((double)(unsigned)m.h));
        const unsigned* ex = ch.x.data();
        const unsigned* ey = ch.y.data();
        std::vector<unsigned> lx(ch.x.size() + 1), ly(ch.x.size() + 1);
        unsigned offsetInLineArray = 0, newOffsetS = 0;
        double lineFitErr = 0;
        double eq[2] = {0, 0};
        for (unsigned edgeID = 0; edgeID < ch.n; edgeID++) {
            unsigned S = ch.sid[edgeID];
            const unsigned E = ch.sid[edgeID + 1];
            while (E > S + MIN_LINE_LEN) {
                while (E > S + MIN_LINE_LEN) {
                    lineFitErr = fit_initial(ex, ey, S, eq);
                    if (lineFitErr <= LINE_FIT_ERR_THRESHOLD) break;
                    S += SKIP_EDGE_POINT;
                }
                if (lineFitErr > LINE_FIT_ERR_THRESHOLD) break;
                const unsigned lineStart = offsetInLineArray;
                const bool horiz = m.dir[ey[S] * m.w + ex[S]] == HORIZONTAL;
                double coef1 = 0;
                bool bExtended = true, bFirstTry = true;
                int tryTimes = 0;
                while (bExtended) {
                    tryTimes++;
                    if (bFirstTry) {
                        bFirstTry = false;
                        for (int i = 0; i < MIN_LINE_LEN; i++) { lx[offsetInLineArray] = ex[S]; ly[offsetInLineArray++] = ey[S++]; }
                    } else {
                        fit_extend(lx.data(), ly.data(), lineStart, newOffsetS, offsetInLineArray, eq);
                    }
                    coef1 = horiz ? 1 / std::sqrt(eq[0] * eq[0] + 1) : 1 / std::sqrt(1 + eq[0] * eq[0]);
                    int numOfOutlier = 0;
                    newOffsetS = offsetInLineArray;
                    while (E > S) {
                        const double d = horiz ? std::fabs(eq[0] * ex[S] - ey[S] + eq[1]) * coef1 : std::fabs(ex[S] - eq[0] * ey[S] - eq[1]) * coef1;
                        lx[offsetInLineArray] = ex[S];
                        ly[offsetInLineArray++] = ey[S++];
                        if (d > LINE_FIT_ERR_THRESHOLD) {
                            numOfOutlier++;
                            if (numOfOutlier > 3) break;
                        } else {
                            numOfOutlier = 0;
                        }
                    }
                    offsetInLineArray -= numOfOutlier;
                    S -= numOfOutlier;
                    if (!(offsetInLineArray - newOffsetS > 0 && tryTimes < TRY_TIME)) bExtended = false;
                }
                Line L;
                if (horiz) { L.eq[0] = eq[0] * coef1; L.eq[1] = -1 * coef1; L.eq[2] = eq[1] * coef1; }
                else { L.eq[0] = 1 * coef1; L.eq[1] = -eq[0] * coef1; L.eq[2] = -eq[1] * coef1; }
                float direction = 0;
                if (validate(lx.data(), ly.data(), lineStart, offsetInLineArray, L.eq, direction)) {
                    const double a1 = L.eq[1] * L.eq[1], a2 = L.eq[0] * L.eq[0], a3 = L.eq[0] * L.eq[1], a4 = L.eq[2] * L.eq[0], a5 = L.eq[2] * L.eq[1];
                    unsigned Px = lx[lineStart], Py = ly[lineStart];
                    L.ep[0] = (float)(a1 * Px - a3 * Py - a4);
                    L.ep[1] = (float)(a2 * Py - a3 * Px - a5);
                    Px = lx[offsetInLineArray - 1]; Py = ly[offsetInLineArray - 1];
                    L.ep[2] = (float)(a1 * Px - a3 * Py - a4);
                    L.ep[3] = (float)(a2 * Py - a3 * Px - a5);
                    L.direction = direction;
                    L.s = lineStart; L.e = offsetInLineArray;
                    lines.push_back(L);
                } else {
                    offsetInLineArray = lineStart;
                }
            }
        }
    }
};

}  // namespace

extern "C" {

void orc_edlines_set_libm_trig(int on) { g_libm_trig = on; }

// gradient map (thresholded, / 4), direction map (255 / 0) and the anchors in scan order (x0 y0 x1 y1 ...); returns the anchor count
int orc_edlines_maps(const uint8_t* gray, int w, int h, int16_t* g_out, uint8_t* dir_out, uint32_t* anchors_out, int anchor_cap) {
    Maps m;
    build_maps(gray, w, h, m);
    if (g_out) std::memcpy(g_out, m.g.data(), m.g.size() * 2);
    if (dir_out) std::memcpy(dir_out, m.dir.data(), m.dir.size());
    Chains ch;
    std::vector<unsigned> anc;
    if (edge_drawing(m, ch, &anc) != 1) return -1;
    const int n = (int)(anc.size() / 2);
    if (anchors_out)
        for (int i = 0; i < 2 * n && i < 2 * anchor_cap; i++) anchors_out[i] = anc[i];
    return n;
}

// edge chains: xy_out = x0 y0 x1 y1 ... of all chains back to back, sid_out[n_chains + 1]; returns n_chains (-1: the reference's capacity error)
int orc_edlines_chains(const uint8_t* gray, int w, int h, uint32_t* xy_out, int px_cap, uint32_t* sid_out, int chain_cap, int* n_px) {
    Maps m;
    build_maps(gray, w, h, m);
    Chains ch;
    if (edge_drawing(m, ch, nullptr) != 1) return -1;
    if (n_px) *n_px = (int)ch.x.size();
    if (xy_out)
        for (size_t i = 0; i < ch.x.size() && (int)i < px_cap; i++) { xy_out[2 * i] = ch.x[i]; xy_out[2 * i + 1] = ch.y[i]; }
    if (sid_out)
        for (size_t i = 0; i < ch.sid.size() && (int)i <= chain_cap; i++) sid_out[i] = ch.sid[i];
    return (int)ch.n;
}

// filter = 1: line_lbd_detect::detect_filter_lines (octave 0, lineLength > length_thres); 0: every key line of BinaryDescriptor::detect.
// lines_out: n x 4 float [startPointX startPointY endPointX endPointY] (keylines_to_mat), extra_out (optional): n x 3 {direction, numOfPixels, lineLength}.
int orc_edlines_detect(const uint8_t* gray, int w, int h, int filter, float length_thres, float* lines_out, float* extra_out, int cap) {
    Maps m;
    build_maps(gray, w, h, m);
    Chains ch;
    if (edge_drawing(m, ch, nullptr) != 1) return -1;
    Detector det(m);
    std::vector<Line> lines;
    det.run(ch, lines);
    int n_out = 0;
    for (const Line& L : lines) {
        // OctaveKeyLines :836-851 (length), :1075-1140 (end-point order by direction); scale[0] = 1
        const float dx = std::fabs(L.ep[0] - L.ep[2]), dy = std::fabs(L.ep[1] - L.ep[3]);
        const float len = std::sqrt(dx * dx + dy * dy);
        const float s1 = L.ep[0], s2 = L.ep[1], e1 = L.ep[2], e2 = L.ep[3];
        const float ddx = e1 - s1, ddy = e2 - s2;
        const float direction = L.direction;
        bool change = false;
        if (direction >= -0.75 * M_PI && direction < -0.25 * M_PI && ddy > 0) change = true;
        if (direction >= -0.25 * M_PI && direction < 0.25 * M_PI && ddx < 0) change = true;
        if (direction >= 0.25 * M_PI && direction < 0.75 * M_PI && ddy < 0) change = true;
        if (((direction >= 0.75 * M_PI && direction < M_PI) || (direction >= -M_PI && direction < -0.75 * M_PI)) && ddx > 0) change = true;
        if (filter && !(len > length_thres)) continue;  // line_lbd_allclass.cpp:204-207
        if (n_out < cap) {
            float* o = lines_out + 4 * (size_t)n_out;
            if (change) { o[0] = e1; o[1] = e2; o[2] = s1; o[3] = s2; }
            else { o[0] = s1; o[1] = s2; o[2] = e1; o[3] = e2; }
            if (extra_out) {
                extra_out[3 * (size_t)n_out] = direction;
                extra_out[3 * (size_t)n_out + 1] = (float)(L.e - L.s);
                extra_out[3 * (size_t)n_out + 2] = len;
            }
        }
        n_out++;
    }
    return n_out;
}

}  // extern "C"
