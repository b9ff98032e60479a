// oracle_lsd.cpp -- TEST INFRASTRUCTURE ONLY (SURVEY.md 8 "next" row f-2).
//
// CPU restatement of the LSD branch of line_lbd_detect::detect_filter_lines (reference: line_lbd/class/line_lbd_allclass.cpp:130-149,
// 200-235 -> LSDDetector::detectImpl line_lbd/libs/LSDDetector.cpp:154-293 -> LineSegmentDetectorImpl line_lbd/libs/lsd.cpp:414-1180).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it; the product never does.
//
// Third-party arithmetic that is NOT in the reference tree (OpenCV, version unpinned by line_lbd/CMakeLists.txt) and is restated here
// from OpenCV's published algorithms:
//   Mat::convertTo(CV_64F)                      u8 -> double, exact
//   cv::getGaussianKernel(7, 0.75, CV_64F)      exp(-x^2 / (2 sigma^2)) normalised (the 3.x formula); a caller may pass the 7 taps instead
//   cv::GaussianBlur on CV_64F                  separable: RowFilter (taps summed left to right) then SymmColumnFilter
//                                               (centre tap, then f_k (S[+k] + S[-k])), BORDER_REFLECT_101
//   cv::resize(.., 0.8, 0.8, INTER_LINEAR)      float coefficients, double data: HResizeLinear then VResizeLinear
//   cv::fastAtan2                               the degree-7 float polynomial of core/src/mathfuncs_core
//   cv::LineIterator                            only its pixel count, which does not reach the output matrix -> not restated
//
// How it is pinned (tests/test_lsd_oracle.py; fixtures tests/golden/lsd_demo.npz + lsd_cv2.npz made by tests/golden/make_lsd_golden.py):
//  (1) the reference's OWN golden vector: detect_3d_cuboid/data/edge_detection/LSD/0000_edge.txt holds the 271 segments the reference's
//      line_lbd node wrote for the bundled image detect_3d_cuboid/data/0000_rgb_raw.jpg.  This restatement reproduces the file: 271
//      segments in the same order, 1082 of 1084 coordinates equal to the 6 printed digits, the worst 7e-4 px off (author's OpenCV and JPEG
//      decoder vs cv2 4.13).  The file pins everything at once -- front end, seed order, region growing, refinement, the NFA search with
//      the reference copy's quirks (integer-division scan-line slopes, `n + 1` instead of log_gamma(n + 1)), border and length filters.
//  (2) cv2 4.13, which ships a newer revision of the same detector (cv2.createLineSegmentDetector): it visits seeds by descending bin of
//      the quantised gradient norm (raster order inside a bin) where the reference's copy walks `list[i]` in raster order (its bin "sort"
//      only relinks ->next pointers it never follows, lsd.cpp:478-481 vs :618-634), runs its front end on the u8 image, and has a
//      repaired NFA.  seed_order = 1 + cv2's own scaled image as input reproduce cv2's LSD_REFINE_NONE and LSD_REFINE_STD outputs bit
//      for bit (everything up to and including refine()); the CV_64F blur + resize and fastAtan2 are compared with cv2's directly.
//
// Specified arithmetic shared with the device path (csrc/lsd.cu): the running region angle sums cosf/sinf of every accepted
// pixel in FLOAT (lsd.cpp:679-680), so a 1-ulp libm difference changes which pixels join a region.  Both sides therefore use det_sincos()
// (Cody-Waite reduction + the fdlibm kernel polynomials, plain IEEE double operations, rounded once to float) instead of libm;
// libm_trig = 1 selects the literal libm calls.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr double CV_PI_ = 3.1415926535897932384626433832795;
constexpr double NOTDEF = -1024.0;
constexpr double M_3_2_PI = (3 * CV_PI_) / 2;
constexpr double M_2__PI = 2 * CV_PI_;
constexpr double DEG_TO_RADS = CV_PI_ / 180;
constexpr double LN10 = 2.30258509299404568402;

// ---- specified sin/cos (see header) --------------------------------------------------------------------------------------------
inline void det_sincos(double x, double* s, double* c) {
    const double two_over_pi = 6.36619772367581382433e-01, pio2_hi = 1.57079632673412561417e+00, pio2_lo = 6.07710050650619224932e-11;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double kf = std::floor(x * two_over_pi + 0.5);
    const int k = (int)kf;
    double r = x - kf * pio2_hi;  // exact: pio2_hi has 33 significant bits, |k| < 2^20
    r = r - kf * pio2_lo;
    const double z = r * r;
    const double sn = r + (z * r) * (S1 + z * (S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)))));
    const double cs = 1.0 - (0.5 * z - z * (z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))))));
    switch (k & 3) {
        case 0: *s = sn; *c = cs; break;
        case 1: *s = cs; *c = -sn; break;
        case 2: *s = -sn; *c = -cs; break;
        default: *s = -cs; *c = sn; break;
    }
}

// cv::fastAtan2 (degrees in [0, 360)), float arithmetic, no contraction
inline float fast_atan2f(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / CV_PI_), p3 = -0.3258083974640975f * (float)(180 / CV_PI_),
                       p5 = 0.1555786518463281f * (float)(180 / CV_PI_), p7 = -0.04432655554792128f * (float)(180 / CV_PI_);
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// lsd.cpp:62-127
inline double distSq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
inline double dist(double x1, double y1, double x2, double y2) { return std::sqrt(distSq(x1, y1, x2, y2)); }
inline double angle_diff_signed(double a, double b) {
    double diff = a - b;
    while (diff <= -CV_PI_) diff += M_2__PI;
    while (diff > CV_PI_) diff -= M_2__PI;
    return diff;
}
inline double angle_diff(double a, double b) { return std::fabs(angle_diff_signed(a, b)); }
inline bool double_equal(double a, double b) {
    if (a == b) return true;
    double abs_diff = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
    double abs_max = (aa > bb) ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
inline double log_gamma_windschitl(double x) {
    return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
}
inline double log_gamma_lanczos(double x) {
    static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
    double b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= std::log(x + double(n));
        b += q[n] * std::pow(x, double(n));
    }
    return a + std::log(b);
}
inline double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

struct RegionPoint {
    int x, y;
    double angle, modgrad;
};
struct Rect {
    double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
};
struct Edge {
    int x, y;
    bool taken;
};
struct NormPoint {
    int x, y, norm;
};

struct LSD {
    int libm_trig = 0;
    int W = 0, H = 0;  // scaled image
    std::vector<double> scaled, angles, modgrad;
    std::vector<uint8_t> used;
    double LOG_NT = 0;
    // speculation study (orc_lsd_spec_sim): when rec is set, every used-flag whose value can matter is logged as read, every change as
    // written (with the old value, so that a speculative seed can be rolled back)
    struct Rec { std::vector<int> reads; std::vector<std::pair<int, uint8_t>> writes; };
    Rec* rec = nullptr;
    void set_used(int a, uint8_t v) {
        if (rec) rec->writes.push_back({a, used[a]});
        used[a] = v;
    }

    void sincosd(double x, double* s, double* c) const {
        if (libm_trig) { *s = std::sin(x); *c = std::cos(x); }
        else det_sincos(x, s, c);
    }
    float cosf_(float x) const {
        if (libm_trig) return std::cos(x);
        double s, c;
        det_sincos((double)x, &s, &c);
        return (float)c;
    }
    float sinf_(float x) const {
        if (libm_trig) return std::sin(x);
        double s, c;
        det_sincos((double)x, &s, &c);
        return (float)s;
    }

    // lsd.cpp:1151-1167
    bool isAligned(int address, double theta, double prec) const {
        if (address < 0) return false;
        const double a = angles[address];
        if (a == NOTDEF) return false;
        double n_theta = theta - a;
        if (n_theta < 0) n_theta = -n_theta;
        if (n_theta > M_3_2_PI) {
            n_theta -= M_2__PI;
            if (n_theta < 0) n_theta = -n_theta;
        }
        return n_theta <= prec;
    }

    // lsd.cpp:538-590 (gradient part of ll_angle)
    double ll_angle(double threshold) {
        angles.assign((size_t)W * H, NOTDEF);
        modgrad.assign((size_t)W * H, 0.0);
        double max_grad = -1;
        for (int y = 0; y < H - 1; ++y)
            for (int addr = y * W, addr_end = addr + W - 1; addr < addr_end; ++addr) {
                double DA = scaled[addr + W + 1] - scaled[addr];
                double BC = scaled[addr + 1] - scaled[addr + W];
                double gx = DA + BC, gy = DA - BC;
                double norm = std::sqrt((gx * gx + gy * gy) / 4);
                modgrad[addr] = norm;
                if (norm <= threshold) angles[addr] = NOTDEF;
                else {
                    angles[addr] = fast_atan2f(float(gx), float(-gy)) * DEG_TO_RADS;
                    if (norm > max_grad) max_grad = norm;
                }
            }
        return max_grad;
    }

    // lsd.cpp:637-688
    void region_grow(int sx, int sy, std::vector<RegionPoint>& reg, int& reg_size, double& reg_angle, double prec) {
        reg_size = 1;
        int addr = sx + sy * W;
        reg[0].x = sx;
        reg[0].y = sy;
        reg_angle = angles[addr];
        reg[0].angle = reg_angle;
        reg[0].modgrad = modgrad[addr];
        double s0, c0;
        sincosd(reg_angle, &s0, &c0);
        float sumdx = float(c0), sumdy = float(s0);
        set_used(addr, 1);
        for (int i = 0; i < reg_size; ++i) {
            const int px = reg[i].x, py = reg[i].y;
            int xx_min = std::max(px - 1, 0), xx_max = std::min(px + 1, W - 1);
            int yy_min = std::max(py - 1, 0), yy_max = std::min(py + 1, H - 1);
            for (int yy = yy_min; yy <= yy_max; ++yy) {
                int c_addr = xx_min + yy * W;
                for (int xx = xx_min; xx <= xx_max; ++xx, ++c_addr) {
                    // (the alignment test comes first here: same result, and only aligned pixels' used flags influence the region)
                    const bool al = isAligned(c_addr, reg_angle, prec);
                    if (al && rec) rec->reads.push_back(c_addr);
                    if (al && used[c_addr] != 1) {
                        set_used(c_addr, 1);
                        RegionPoint& rp = reg[reg_size];
                        rp.x = xx;
                        rp.y = yy;
                        rp.modgrad = modgrad[c_addr];
                        const double angle = angles[c_addr];
                        rp.angle = angle;
                        ++reg_size;
                        sumdx += cosf_(float(angle));
                        sumdy += sinf_(float(angle));
                        reg_angle = fast_atan2f(sumdy, sumdx) * DEG_TO_RADS;
                    }
                }
            }
        }
    }

    // lsd.cpp:748-784
    double get_theta(const std::vector<RegionPoint>& reg, int reg_size, double x, double y, double reg_angle, double prec) const {
        double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
        for (int i = 0; i < reg_size; ++i) {
            const double regx = reg[i].x, regy = reg[i].y, weight = reg[i].modgrad;
            double dx = regx - x, dy = regy - y;
            Ixx += dy * dy * weight;
            Iyy += dx * dx * weight;
            Ixy -= dx * dy * weight;
        }
        double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? double(fast_atan2f(float(lambda - Ixx), float(Ixy)))
                                                         : double(fast_atan2f(float(Ixy), float(lambda - Iyy)));
        theta *= DEG_TO_RADS;
        if (angle_diff(theta, reg_angle) > prec) theta += CV_PI_;
        return theta;
    }

    // lsd.cpp:690-746
    void region2rect(const std::vector<RegionPoint>& reg, int reg_size, double reg_angle, double prec, double p, Rect& rec) const {
        double x = 0, y = 0, sum = 0;
        for (int i = 0; i < reg_size; ++i) {
            const double weight = reg[i].modgrad;
            x += double(reg[i].x) * weight;
            y += double(reg[i].y) * weight;
            sum += weight;
        }
        x /= sum;
        y /= sum;
        double theta = get_theta(reg, reg_size, x, y, reg_angle, prec);
        double dx, dy;
        sincosd(theta, &dy, &dx);
        double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
        for (int i = 0; i < reg_size; ++i) {
            double regdx = double(reg[i].x) - x, regdy = double(reg[i].y) - y;
            double l = regdx * dx + regdy * dy;
            double w = -regdx * dy + regdy * dx;
            if (l > l_max) l_max = l;
            else if (l < l_min) l_min = l;
            if (w > w_max) w_max = w;
            else if (w < w_min) w_min = w;
        }
        rec.x1 = x + l_min * dx;
        rec.y1 = y + l_min * dy;
        rec.x2 = x + l_max * dx;
        rec.y2 = y + l_max * dy;
        rec.width = w_max - w_min;
        rec.x = x;
        rec.y = y;
        rec.theta = theta;
        rec.dx = dx;
        rec.dy = dy;
        rec.prec = prec;
        rec.p = p;
        if (rec.width < 1.0) rec.width = 1.0;
    }

    // lsd.cpp:834-871
    bool reduce_region_radius(std::vector<RegionPoint>& reg, int& reg_size, double reg_angle, double prec, double p, Rect& rec, double density,
                              double density_th) {
        double xc = double(reg[0].x), yc = double(reg[0].y);
        double radSq1 = distSq(xc, yc, rec.x1, rec.y1), radSq2 = distSq(xc, yc, rec.x2, rec.y2);
        double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
        while (density < density_th) {
            radSq *= 0.75 * 0.75;
            for (int i = 0; i < reg_size; ++i) {
                if (distSq(xc, yc, double(reg[i].x), double(reg[i].y)) > radSq) {
                    set_used(reg[i].x + reg[i].y * W, 0);
                    std::swap(reg[i], reg[reg_size - 1]);
                    --reg_size;
                    --i;
                }
            }
            if (reg_size < 2) return false;
            region2rect(reg, reg_size, reg_angle, prec, p, rec);
            density = double(reg_size) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        }
        return true;
    }

    // lsd.cpp:786-832
    bool refine(std::vector<RegionPoint>& reg, int& reg_size, double reg_angle, double prec, double p, Rect& rec, double density_th) {
        double density = double(reg_size) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= density_th) return true;
        double xc = double(reg[0].x), yc = double(reg[0].y);
        const double ang_c = reg[0].angle;
        double sum = 0, s_sum = 0;
        int n = 0;
        for (int i = 0; i < reg_size; ++i) {
            set_used(reg[i].x + reg[i].y * W, 0);
            if (dist(xc, yc, reg[i].x, reg[i].y) < rec.width) {
                double ang_d = angle_diff_signed(reg[i].angle, ang_c);
                sum += ang_d;
                s_sum += ang_d * ang_d;
                ++n;
            }
        }
        double mean_angle = sum / double(n);
        double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
        region_grow(reg[0].x, reg[0].y, reg, reg_size, reg_angle, tau);
        if (reg_size < 2) return false;
        region2rect(reg, reg_size, reg_angle, prec, p, rec);
        density = double(reg_size) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density < density_th) return reduce_region_radius(reg, reg_size, reg_angle, prec, p, rec, density, density_th);
        return true;
    }

    // lsd.cpp:1100-1149
    double nfa(int n, int k, double p) const {
        if (n == 0 || k == 0) return -LOG_NT;
        if (n == k) return -LOG_NT - double(n) * std::log10(p);
        double p_term = p / (1 - p);
        double log1term = (double(n) + 1) - log_gamma(double(k) + 1) - log_gamma(double(n - k) + 1) + double(k) * std::log(p) +
                          double(n - k) * std::log(1.0 - p);
        double term = std::exp(log1term);
        if (double_equal(term, 0)) {
            if (k > n * p) return -log1term / LN10 - LOG_NT;
            return -LOG_NT;
        }
        double bin_tail = term, tolerance = 0.1;
        for (int i = k + 1; i <= n; ++i) {
            double bin_term = double(n - i + 1) / double(i);
            double mult_term = bin_term * p_term;
            term *= mult_term;
            bin_tail += term;
            if (bin_term < 1) {
                double err = term * ((1 - std::pow(mult_term, double(n - i + 1))) / (1 - mult_term) - 1);
                if (err < tolerance * std::fabs(-std::log10(bin_tail) - LOG_NT) * bin_tail) break;
            }
        }
        return -std::log10(bin_tail) - LOG_NT;
    }

    // lsd.cpp:977-1098 (integer-division slopes and the tailp->p.x comparisons are the reference's)
    double rect_nfa(const Rect& rec) const {
        int total_pts = 0, alg_pts = 0;
        double half_width = rec.width / 2.0;
        double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
        Edge o[4];
        o[0] = {int(rec.x1 - dyhw), int(rec.y1 + dxhw), false};
        o[1] = {int(rec.x2 - dyhw), int(rec.y2 + dxhw), false};
        o[2] = {int(rec.x2 + dyhw), int(rec.y2 - dxhw), false};
        o[3] = {int(rec.x1 + dyhw), int(rec.y1 - dxhw), false};
        std::sort(o, o + 4, [](const Edge& a, const Edge& b) { return a.x == b.x ? a.y < b.y : a.x < b.x; });
        Edge *min_y = &o[0], *max_y = &o[0];
        for (int i = 1; i < 4; ++i) {
            if (min_y->y > o[i].y) min_y = &o[i];
            if (max_y->y < o[i].y) max_y = &o[i];
        }
        min_y->taken = true;
        Edge* leftmost = nullptr;
        for (int i = 0; i < 4; ++i)
            if (!o[i].taken) {
                if (!leftmost) leftmost = &o[i];
                else if (leftmost->x > o[i].x) leftmost = &o[i];
            }
        leftmost->taken = true;
        Edge* rightmost = nullptr;
        for (int i = 0; i < 4; ++i)
            if (!o[i].taken) {
                if (!rightmost) rightmost = &o[i];
                else if (rightmost->x < o[i].x) rightmost = &o[i];
            }
        rightmost->taken = true;
        Edge* tailp = nullptr;
        for (int i = 0; i < 4; ++i)
            if (!o[i].taken) {
                if (!tailp) tailp = &o[i];
                else if (tailp->x > o[i].x) tailp = &o[i];
            }
        tailp->taken = true;
        double flstep = (min_y->y != leftmost->y) ? (min_y->x - leftmost->x) / (min_y->y - leftmost->y) : 0;
        double slstep = (leftmost->y != tailp->x) ? (leftmost->x - tailp->x) / (leftmost->y - tailp->x) : 0;
        double frstep = (min_y->y != rightmost->y) ? (min_y->x - rightmost->x) / (min_y->y - rightmost->y) : 0;
        double srstep = (rightmost->y != tailp->x) ? (rightmost->x - tailp->x) / (rightmost->y - tailp->x) : 0;
        double lstep = flstep, rstep = frstep;
        double left_x = min_y->x, right_x = min_y->x;
        for (int y = min_y->y; y <= max_y->y; ++y) {
            if (y < 0 || y >= H) continue;
            int adx = y * W + int(left_x);
            for (int x = int(left_x); x <= int(right_x); ++x, ++adx) {
                if (x < 0 || x >= W) continue;
                ++total_pts;
                if (isAligned(adx, rec.theta, rec.prec)) ++alg_pts;
            }
            if (y >= leftmost->y) lstep = slstep;
            if (y >= rightmost->y) rstep = srstep;
            left_x += lstep;
            right_x += rstep;
        }
        return nfa(total_pts, alg_pts, rec.p);
    }

    // lsd.cpp:873-975
    double rect_improve(Rect& rec, double LOG_EPS) const {
        const double delta = 0.5, delta_2 = delta / 2.0;
        double log_nfa = rect_nfa(rec);
        if (log_nfa > LOG_EPS) return log_nfa;
        Rect r = rec;
        for (int n = 0; n < 5; ++n) {
            r.p /= 2;
            r.prec = r.p * CV_PI_;
            double v = rect_nfa(r);
            if (v > log_nfa) { log_nfa = v; rec = r; }
        }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.width -= delta;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.x1 += -r.dy * delta_2;
                r.y1 += r.dx * delta_2;
                r.x2 += -r.dy * delta_2;
                r.y2 += r.dx * delta_2;
                r.width -= delta;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.x1 -= -r.dy * delta_2;
                r.y1 -= r.dx * delta_2;
                r.x2 -= -r.dy * delta_2;
                r.y2 -= r.dx * delta_2;
                r.width -= delta;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.p /= 2;
                r.prec = r.p * CV_PI_;
                double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        return log_nfa;
    }
};

inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

// cv::GaussianBlur(CV_64F, 7x7) -- see header
void gaussian_blur7(const uint8_t* gray, int w, int h, const double k[7], std::vector<double>& out) {
    std::vector<double> rowf((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = gray + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            double acc = k[0] * double(s[reflect101(x - 3, w)]);
            for (int t = 1; t < 7; ++t) acc += k[t] * double(s[reflect101(x - 3 + t, w)]);
            rowf[(size_t)y * w + x] = acc;
        }
    }
    out.resize((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double acc = k[3] * rowf[(size_t)y * w + x];
            for (int t = 1; t <= 3; ++t)
                acc += k[3 + t] * (rowf[(size_t)reflect101(y + t, h) * w + x] + rowf[(size_t)reflect101(y - t, h) * w + x]);
            out[(size_t)y * w + x] = acc;
        }
}

// cv::resize(.., Size(), 0.8, 0.8, INTER_LINEAR) on CV_64F -- see header
void resize_linear(const std::vector<double>& src, int sw, int sh, double scale, std::vector<double>& dst, int& dw, int& dh) {
    dw = (int)std::lrint(sw * scale);
    dh = (int)std::lrint(sh * scale);
    const double inv = 1.0 / scale;
    std::vector<int> xofs(dw);
    std::vector<float> xa(dw);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * inv - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        xa[dx] = fx;
    }
    dst.resize((size_t)dw * dh);
    std::vector<double> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * inv - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        const float b0 = 1.f - fy, b1 = fy;
        const int y0 = std::min(std::max(sy, 0), sh - 1), y1 = std::min(std::max(sy + 1, 0), sh - 1);
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx];
            const float a1 = xa[dx], a0 = 1.f - a1;
            if (sx + 1 < sw) {
                r0[dx] = src[(size_t)y0 * sw + sx] * a0 + src[(size_t)y0 * sw + sx + 1] * a1;
                r1[dx] = src[(size_t)y1 * sw + sx] * a0 + src[(size_t)y1 * sw + sx + 1] * a1;
            } else {
                r0[dx] = src[(size_t)y0 * sw + sx];
                r1[dx] = src[(size_t)y1 * sw + sx];
            }
        }
        for (int dx = 0; dx < dw; ++dx) dst[(size_t)dy * dw + dx] = r0[dx] * b0 + r1[dx] * b1;
    }
}

}  // namespace

extern "C" {

// cv::getGaussianKernel(7, sigma, CV_64F), OpenCV 3.x formula (sigma = SIGMA_SCALE / SCALE, lsd.cpp:453-457)
void orc_lsd_gauss_kernel(double* k7) {
    const double sigma = 0.6 / 0.8;
    const double scale2X = -0.5 / (sigma * sigma);
    double sum = 0;
    for (int i = 0; i < 7; ++i) {
        double x = i - 3.0;
        k7[i] = std::exp(scale2X * x * x);
        sum += k7[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < 7; ++i) k7[i] *= sum;
}

float orc_lsd_fast_atan2(float y, float x) { return fast_atan2f(y, x); }
void orc_lsd_det_sincos(double x, double* s, double* c) { det_sincos(x, s, c); }

// stage outputs for parity checks of the streaming kernels: scaled image, gradient norm, level-line angle (-1024 = undefined)
int orc_lsd_maps(const uint8_t* gray, int w, int h, const double* gauss7, double* scaled_out, double* modgrad_out, double* angles_out, int* sw,
                 int* sh) {
    double k7[7];
    if (gauss7) std::memcpy(k7, gauss7, sizeof k7);
    else orc_lsd_gauss_kernel(k7);
    std::vector<double> blur;
    gaussian_blur7(gray, w, h, k7, blur);
    LSD L;
    resize_linear(blur, w, h, 0.8, L.scaled, L.W, L.H);
    const double prec = CV_PI_ * 22.5 / 180;
    L.ll_angle(2.0 / std::sin(prec));
    *sw = L.W;
    *sh = L.H;
    const size_t n = (size_t)L.W * L.H;
    if (scaled_out) std::memcpy(scaled_out, L.scaled.data(), n * 8);
    if (modgrad_out) std::memcpy(modgrad_out, L.modgrad.data(), n * 8);
    if (angles_out) std::memcpy(angles_out, L.angles.data(), n * 8);
    return 0;
}

// mode 0: LineSegmentDetectorImpl::detect (lsd.cpp:414-438) -- every segment, [x1 y1 x2 y2] float in image coordinates.
// mode 1: line_lbd_detect::detect_filter_lines (line_lbd_allclass.cpp:200-235): clamp (LSDDetector.cpp:80-101), drop segments with both
//         ends within 10 px of one border (:206, 228-232), keep lineLength > length_thres (line_lbd_allclass.cpp:206).
// seed_order 0: raster (reference), 1: descending gradient-norm bin, raster inside a bin (cv2 4.x).  refine: 0 LSD_REFINE_NONE, 1 _STD, 2 _ADV (what
// LSDDetector.cpp:174 constructs).  Returns the number of segments (<= cap written),
// nfa_out (optional) receives 12 doubles per written segment: log_nfa, width, p (cv2's nfa / width / prec outputs), then the
// rectangle in scaled-image coordinates x1 y1 x2 y2 (image coordinates: (v + 0.5) / 0.8) theta dx dy prec and the seed's pixel index.
int orc_lsd_detect(const uint8_t* gray, int w, int h, const double* gauss7, const double* scaled_override, int seed_order, int libm_trig,
                   int refine, int mode, float length_thres, float* lines_out, double* nfa_out, int cap) {
    const double SCALE = 0.8, ANG_TH = 22.5, LOG_EPS = 0, DENSITY_TH = 0.7;
    const int N_BINS = 1024;
    double k7[7];
    if (gauss7) std::memcpy(k7, gauss7, sizeof k7);
    else orc_lsd_gauss_kernel(k7);
    LSD L;
    L.libm_trig = libm_trig;
    if (scaled_override) {  // pinning tests only: start from a scaled image produced elsewhere (cv2's own u8 front end)
        L.W = (int)std::lrint(w * SCALE);
        L.H = (int)std::lrint(h * SCALE);
        L.scaled.assign(scaled_override, scaled_override + (size_t)L.W * L.H);
    } else {
        std::vector<double> blur;
        gaussian_blur7(gray, w, h, k7, blur);
        resize_linear(blur, w, h, SCALE, L.scaled, L.W, L.H);
    }
    const double prec = CV_PI_ * ANG_TH / 180, p = ANG_TH / 180, rho = 2.0 / std::sin(prec);
    const double max_grad = L.ll_angle(rho);
    const int W = L.W, H = L.H;
    L.LOG_NT = 5 * (std::log10(double(W)) + std::log10(double(H))) / 2 + std::log10(11.0);
    const int min_reg_size = int(-L.LOG_NT / std::log10(p));
    L.used.assign((size_t)W * H, 0);
    std::vector<RegionPoint> reg((size_t)W * H);

    std::vector<NormPoint> order;
    order.reserve((size_t)W * H);
    if (seed_order == 1) {
        const double bin_coef = (max_grad > 0) ? double(N_BINS - 1) / max_grad : 0;
        for (int y = 0; y < H - 1; ++y)
            for (int x = 0; x < W - 1; ++x) order.push_back({x, y, int(L.modgrad[(size_t)y * W + x] * bin_coef)});
        // cv2 4.13 visits equal-bin seeds in raster order (found empirically: bit-identical segment lists on every test image with a
        // stable sort, a swapped pair here and there with std::sort)
        std::stable_sort(order.begin(), order.end(), [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; });
    } else {
        // lsd.cpp:478-481: list[i] in the order the points were appended = raster order over (H-1) x (W-1)
        for (int y = 0; y < H - 1; ++y)
            for (int x = 0; x < W - 1; ++x) order.push_back({x, y, 0});
    }

    int n_out = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        const int sx = order[i].x, sy = order[i].y;
        const int adx = sx + sy * W;
        if (L.used[adx] != 0 || L.angles[adx] == NOTDEF) continue;
        int reg_size;
        double reg_angle;
        L.region_grow(sx, sy, reg, reg_size, reg_angle, prec);
        if (reg_size < min_reg_size) continue;
        Rect rec;
        L.region2rect(reg, reg_size, reg_angle, prec, p, rec);
        double log_nfa = -1;
        if (refine > 0) {  // lsd.cpp:493-504
            if (!L.refine(reg, reg_size, reg_angle, prec, p, rec, DENSITY_TH)) continue;
            if (refine >= 2) {
                log_nfa = L.rect_improve(rec, LOG_EPS);
                if (log_nfa <= LOG_EPS) continue;
            }
        }
        rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
        rec.x1 /= SCALE; rec.y1 /= SCALE; rec.x2 /= SCALE; rec.y2 /= SCALE;
        float e[4] = {float(rec.x1), float(rec.y1), float(rec.x2), float(rec.y2)};
        if (mode == 1) {
            // LSDDetector.cpp:80-101, 219-232; octaveScale = 1
            if (e[0] < 0) e[0] = 0;
            if (e[0] >= w) e[0] = (float)w - 1.0f;
            if (e[2] < 0) e[2] = 0;
            if (e[2] >= w) e[2] = (float)w - 1.0f;
            if (e[1] < 0) e[1] = 0;
            if (e[1] >= h) e[1] = (float)h - 1.0f;
            if (e[3] < 0) e[3] = 0;
            if (e[3] >= h) e[3] = (float)h - 1.0f;
            const float thr = 10;
            if (((e[0] < thr) && (e[2] < thr)) || ((e[0] > w - thr) && (e[2] > w - thr)) || ((e[1] < thr) && (e[3] < thr)) ||
                ((e[1] > h - thr) && (e[3] > h - thr)))
                continue;
            const double ddx = double(e[0] - e[2]), ddy = double(e[1] - e[3]);
            const float len = (float)std::sqrt(ddx * ddx + ddy * ddy);  // :239
            if (!(len > length_thres)) continue;                        // line_lbd_allclass.cpp:206
        }
        if (n_out < cap) {
            std::memcpy(lines_out + 4 * (size_t)n_out, e, sizeof e);
            if (nfa_out) {
                double* q = nfa_out + 12 * (size_t)n_out;
                q[0] = log_nfa; q[1] = rec.width / SCALE; q[2] = rec.p; q[3] = rec.x1; q[4] = rec.y1; q[5] = rec.x2; q[6] = rec.y2;
                q[7] = rec.theta; q[8] = rec.dx; q[9] = rec.dy; q[10] = rec.prec; q[11] = double(sx + sy * W);
            }
        }
        ++n_out;
    }
    return n_out;
}


// Study for the next optimisation step of the GPU path: can consecutive seeds be processed speculatively in parallel?
// Seeds are taken K at a time in raster order; each one is processed from the state at the start of the wave while its reads (used
// flags of aligned pixels) and writes are logged, then rolled back; they commit in order, and a seed is valid iff nothing it read or
// wrote was written by a seed committed before it in the same wave -- then it behaves exactly as in the sequential order.  The wave
// ends at the first invalid seed.  Returns the number of segments (identical to orc_lsd_detect by construction; the caller checks) and
// fills stats: [0] seeds processed sequentially, [1] waves, [2] speculative seed executions, [3] work of the committed seeds
// (logged reads + writes), [4] sum over waves of the largest work in the wave (the parallel critical path), [5] total speculative work.
int orc_lsd_spec_sim(const uint8_t* gray, int w, int h, int K, float* lines_out, int cap, double* stats) {
    const double SCALE = 0.8, ANG_TH = 22.5, LOG_EPS = 0, DENSITY_TH = 0.7;
    double k7[7];
    orc_lsd_gauss_kernel(k7);
    LSD L;
    {
        std::vector<double> blur;
        gaussian_blur7(gray, w, h, k7, blur);
        resize_linear(blur, w, h, SCALE, L.scaled, L.W, L.H);
    }
    const double prec = CV_PI_ * ANG_TH / 180, p = ANG_TH / 180, rho = 2.0 / std::sin(prec);
    L.ll_angle(rho);
    const int W = L.W, H = L.H;
    L.LOG_NT = 5 * (std::log10(double(W)) + std::log10(double(H))) / 2 + std::log10(11.0);
    const int min_reg_size = int(-L.LOG_NT / std::log10(p));
    L.used.assign((size_t)W * H, 0);
    std::vector<RegionPoint> reg((size_t)W * H);
    struct Spec { int seed; LSD::Rec rec; std::vector<std::pair<int, uint8_t>> finals; bool line; float e[4]; };
    auto run_seed = [&](int adx, Spec& S) {
        S.seed = adx;
        S.line = false;
        S.rec.reads.clear();
        S.rec.writes.clear();
        L.rec = &S.rec;
        S.rec.reads.push_back(adx);
        int reg_size;
        double reg_angle;
        L.region_grow(adx % W, adx / W, reg, reg_size, reg_angle, prec);
        if (reg_size >= min_reg_size) {
            Rect rec;
            L.region2rect(reg, reg_size, reg_angle, prec, p, rec);
            if (L.refine(reg, reg_size, reg_angle, prec, p, rec, DENSITY_TH)) {
                const double log_nfa = L.rect_improve(rec, LOG_EPS);
                if (log_nfa > LOG_EPS) {
                    rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
                    rec.x1 /= SCALE; rec.y1 /= SCALE; rec.x2 /= SCALE; rec.y2 /= SCALE;
                    S.e[0] = float(rec.x1); S.e[1] = float(rec.y1); S.e[2] = float(rec.x2); S.e[3] = float(rec.y2);
                    S.line = true;
                }
            }
        }
        L.rec = nullptr;
        // final values of everything written, then roll back
        S.finals.clear();
        for (auto& wv : S.rec.writes) S.finals.push_back({wv.first, L.used[wv.first]});
        for (size_t i = S.rec.writes.size(); i-- > 0;) L.used[S.rec.writes[i].first] = S.rec.writes[i].second;
    };
    std::vector<Spec> wave(K);
    std::vector<int> mark((size_t)W * H, -1);
    for (int i = 0; i < 6; i++) stats[i] = 0;
    int n_out = 0, pos = 0, wave_id = 0;
    const int N = W * H;
    while (pos < N) {
        int n = 0;
        for (int j = pos; j < N && n < K; j++)
            if (L.used[j] == 0 && L.angles[j] != NOTDEF) { run_seed(j, wave[n]); n++; }
        if (n == 0) break;
        double wmax = 0;
        for (int k = 0; k < n; k++) {
            const double wk = double(wave[k].rec.reads.size() + wave[k].rec.writes.size());
            wmax = std::max(wmax, wk);
            stats[5] += wk;
        }
        stats[2] += n;
        stats[1] += 1;
        stats[4] += wmax;
        int committed = 0;
        for (int k = 0; k < n; k++) {
            bool ok = true;
            if (k > 0) {
                for (int a : wave[k].rec.reads) if (mark[a] == wave_id) { ok = false; break; }
                if (ok) for (auto& wv : wave[k].rec.writes) if (mark[wv.first] == wave_id) { ok = false; break; }
            }
            if (!ok) break;
            for (auto& fv : wave[k].finals) { L.used[fv.first] = fv.second; mark[fv.first] = wave_id; }
            if (wave[k].line) {
                if (n_out < cap) std::memcpy(lines_out + 4 * (size_t)n_out, wave[k].e, 16);
                ++n_out;
            }
            stats[3] += double(wave[k].rec.reads.size() + wave[k].rec.writes.size());
            stats[0] += 1;
            committed++;
        }
        pos = wave[committed - 1].seed + 1;
        wave_id++;
    }
    return n_out;
}

}  // extern "C"
