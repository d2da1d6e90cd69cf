// yavo_oracle_klt.cpp — CPU ORACLE for the sparse pyramidal Lucas-Kanade tracker (test infrastructure
// only; see yavo_oracle.h).
//
// The reference tracks map points from the last frame into the current one with
//     cv::calcOpticalFlowPyrLK(last, cur, lastKpt, curKpt, status, error, cv::Size(11,11), 3,
//                              TermCriteria(COUNT+EPS, 30, 0.01), 0, 0.001)        (src/LoopHandler.cc:372-375)
// — the only place an image pyramid is built.  The arithmetic is OpenCV's (module video,
// lkpyramid.cpp, and imgproc pyrDown), which is not under the reference checkout; the reference pins no OpenCV
// version.  This file restates the published algorithm as OpenCV 4.x executes it on 8-bit single-channel images:
//   * pyramid level k+1 = cv::pyrDown(level k): separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8;
//     levels stop when a side would be <= the window side;
//   * Scharr derivatives of the previous image's level ([3 10 3] x [-1 0 1], reflect-101 inside the image,
//     ZERO outside it), the image itself REFLECT_101-padded by the window size;
//   * per point and level: 14-bit fixed-point bilinear weights (cvRound = round half to even), patch and
//     derivative samples descaled to 5 / 0 fractional bits, the 2x2 gradient matrix, minimum-eigenvalue gate,
//     up to maxCount Newton steps, the 0.01 px oscillation stop, and the mean absolute patch difference as err.
//   * the float32 sums A11, A12, A22, b1, b2 are accumulated in the ORDER of OpenCV's 128-bit universal-intrinsic
//     loops (x86 build).  Columns are taken in groups of 8 (x < 8*floor(w/8)); the remaining columns go to a
//     scalar float accumulator in scan order.  Gradient matrix: four lane accumulators, lane k takes columns
//     k and k+4 of every group (float products, exact), result = tail + ((q0+q2)+(q1+q3)).  Mismatch vector: the
//     int32 pair sums (x, x+4) of every group are converted to float and added to four accumulators, result =
//     tail + ((p04+p26)+(p15+p37)); tail terms are (float)(int product).  Float addition is not associative, so
//     this order is part of the definition (found by matching cv2 bit for bit over window sizes 5..31).
// Parity status: pinned BIT FOR BIT against cv2 4.13.0 (this image's opencv-python-headless: SSE3 baseline,
// lkpyramid.cpp not dispatched) by tests/golden/klt_golden.npz, made by tests/golden/make_klt_golden.py: pyramid
// levels, Scharr planes, tracked positions, status flags and err.  An OpenCV built for NEON, or one whose
// lkpyramid.cpp vectorises differently, can differ in the last float bits of a step (and then, for badly
// conditioned points, in where the iteration ends); the reference pins no OpenCV version.
// Coordinates are OpenCV's: point = (x = column, y = row).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "yavo_oracle.h"

namespace {

inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

void pyr_down(const uint8_t *src, int H, int W, uint8_t *dst) {
    const int oh = (H + 1) / 2, ow = (W + 1) / 2;
    std::vector<int> hp((size_t)H * ow);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < ow; x++) {
            const uint8_t *r = src + (size_t)y * W;
            hp[(size_t)y * ow + x] = r[reflect101(2 * x - 2, W)] + 4 * r[reflect101(2 * x - 1, W)] + 6 * r[reflect101(2 * x, W)] +
                                     4 * r[reflect101(2 * x + 1, W)] + r[reflect101(2 * x + 2, W)];
        }
    for (int y = 0; y < oh; y++)
        for (int x = 0; x < ow; x++) {
            const int s = hp[(size_t)reflect101(2 * y - 2, H) * ow + x] + 4 * hp[(size_t)reflect101(2 * y - 1, H) * ow + x] +
                          6 * hp[(size_t)reflect101(2 * y, H) * ow + x] + 4 * hp[(size_t)reflect101(2 * y + 1, H) * ow + x] +
                          hp[(size_t)reflect101(2 * y + 2, H) * ow + x];
            dst[(size_t)y * ow + x] = (uint8_t)((s + 128) >> 8);
        }
}

void scharr(const uint8_t *img, int H, int W, int16_t *dx, int16_t *dy) {
    std::vector<int> t0(W), t1(W);
    for (int y = 0; y < H; y++) {
        const uint8_t *up = img + (size_t)reflect101(y - 1, H) * W, *mid = img + (size_t)y * W,
                      *dn = img + (size_t)reflect101(y + 1, H) * W;
        for (int x = 0; x < W; x++) {
            t0[x] = (up[x] + dn[x]) * 3 + mid[x] * 10;
            t1[x] = dn[x] - up[x];
        }
        for (int x = 0; x < W; x++) {
            const int l = reflect101(x - 1, W), r = reflect101(x + 1, W);
            dx[(size_t)y * W + x] = (int16_t)(t0[r] - t0[l]);
            dy[(size_t)y * W + x] = (int16_t)((t1[r] + t1[l]) * 3 + t1[x] * 10);
        }
    }
}

struct Level {
    int H, W;
    std::vector<uint8_t> I, J;
    std::vector<int16_t> dx, dy;
};

inline int cv_round(float v) { return (int)std::nearbyintf(v); }  // default rounding mode: half to even
inline int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

struct Weights {
    int w00, w01, w10, w11;
};
inline Weights weights(float a, float b) {
    Weights w;
    w.w00 = cv_round((1.f - a) * (1.f - b) * (1 << 14));
    w.w01 = cv_round(a * (1.f - b) * (1 << 14));
    w.w10 = cv_round((1.f - a) * b * (1 << 14));
    w.w11 = (1 << 14) - w.w00 - w.w01 - w.w10;
    return w;
}

inline int pix(const std::vector<uint8_t> &im, int H, int W, int y, int x) {  // REFLECT_101 padding
    return im[(size_t)reflect101(y, H) * W + reflect101(x, W)];
}
inline int der(const std::vector<int16_t> &d, int H, int W, int y, int x) {  // zero padding
    return (y >= 0 && y < H && x >= 0 && x < W) ? d[(size_t)y * W + x] : 0;
}

long long g_iterations = 0;  // Newton steps of the last yavo_oracle_klt call (workload statistics for the benchmarks)

}  // namespace

extern "C" {

long long yavo_oracle_klt_iterations(void) { return g_iterations; }


void yavo_oracle_pyr_down(const uint8_t *img, int H, int W, uint8_t *out) { pyr_down(img, H, W, out); }

void yavo_oracle_scharr(const uint8_t *img, int H, int W, int16_t *dx, int16_t *dy) { scharr(img, H, W, dx, dy); }

int yavo_oracle_klt_levels(int H, int W, int win_w, int win_h, int max_level) {
    int lv = 0;
    while (lv < max_level) {
        const int nh = (H + 1) / 2, nw = (W + 1) / 2;
        if (nw <= win_w || nh <= win_h) break;
        H = nh;
        W = nw;
        lv++;
    }
    return lv;
}

int yavo_oracle_klt(const uint8_t *prev, const uint8_t *next, int H, int W, const float *prev_xy, int n, float *next_xy,
                    uint8_t *status, float *err, int win_w, int win_h, int max_level, int crit_type, int max_count,
                    double epsilon, int flags, double min_eig_threshold) {
    // criteria clamping of SparsePyrLKOpticalFlowImpl::calc
    if ((crit_type & 1) == 0) max_count = 30;
    else max_count = std::min(std::max(max_count, 0), 100);
    if ((crit_type & 2) == 0) epsilon = 0.01;
    else epsilon = std::min(std::max(epsilon, 0.), 10.);
    epsilon *= epsilon;
    const bool use_initial = (flags & 4) != 0, get_min_eig = (flags & 8) != 0;
    g_iterations = 0;

    const int levels = yavo_oracle_klt_levels(H, W, win_w, win_h, max_level);
    std::vector<Level> pyr(levels + 1);
    pyr[0].H = H;
    pyr[0].W = W;
    pyr[0].I.assign(prev, prev + (size_t)H * W);
    pyr[0].J.assign(next, next + (size_t)H * W);
    for (int l = 1; l <= levels; l++) {
        pyr[l].H = (pyr[l - 1].H + 1) / 2;
        pyr[l].W = (pyr[l - 1].W + 1) / 2;
        pyr[l].I.resize((size_t)pyr[l].H * pyr[l].W);
        pyr[l].J.resize((size_t)pyr[l].H * pyr[l].W);
        pyr_down(pyr[l - 1].I.data(), pyr[l - 1].H, pyr[l - 1].W, pyr[l].I.data());
        pyr_down(pyr[l - 1].J.data(), pyr[l - 1].H, pyr[l - 1].W, pyr[l].J.data());
    }
    for (int l = 0; l <= levels; l++) {
        pyr[l].dx.resize((size_t)pyr[l].H * pyr[l].W);
        pyr[l].dy.resize((size_t)pyr[l].H * pyr[l].W);
        scharr(pyr[l].I.data(), pyr[l].H, pyr[l].W, pyr[l].dx.data(), pyr[l].dy.data());
    }
    for (int i = 0; i < n; i++) {
        status[i] = 1;
        if (err) err[i] = 0.f;
    }
    const float FLT_SCALE = 1.f / (1 << 20);
    const float halfx = (win_w - 1) * 0.5f, halfy = (win_h - 1) * 0.5f;
    const int area = win_w * win_h;
    std::vector<int> Iw(area), Ix(area), Iy(area);

    for (int level = levels; level >= 0; level--) {
        const Level &L = pyr[level];
        const int LH = L.H, LW = L.W;
        for (int i = 0; i < n; i++) {
            const float sc = (float)(1. / (1 << level));
            float px = prev_xy[2 * i] * sc, py = prev_xy[2 * i + 1] * sc;
            float nx, ny;
            if (level == levels) {
                if (use_initial) {
                    nx = next_xy[2 * i] * sc;
                    ny = next_xy[2 * i + 1] * sc;
                } else {
                    nx = px;
                    ny = py;
                }
            } else {
                nx = next_xy[2 * i] * 2.f;
                ny = next_xy[2 * i + 1] * 2.f;
            }
            next_xy[2 * i] = nx;
            next_xy[2 * i + 1] = ny;
            px -= halfx;
            py -= halfy;
            const int ix = (int)std::floor(px), iy = (int)std::floor(py);
            if (ix < -win_w || ix >= LW || iy < -win_h || iy >= LH) {
                if (level == 0) {
                    status[i] = 0;
                    if (err) err[i] = 0.f;
                }
                continue;
            }
            Weights w = weights(px - ix, py - iy);
            float qA[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}, tA[3] = {0, 0, 0};
            const int vecA = (win_w / 8) * 8;
            for (int y = 0; y < win_h; y++)
                for (int x = 0; x < win_w; x++) {
                    const int yy = iy + y, xx = ix + x;
                    const int iv = descale(pix(L.I, LH, LW, yy, xx) * w.w00 + pix(L.I, LH, LW, yy, xx + 1) * w.w01 +
                                               pix(L.I, LH, LW, yy + 1, xx) * w.w10 + pix(L.I, LH, LW, yy + 1, xx + 1) * w.w11,
                                           14 - 5);
                    const int gx = descale(der(L.dx, LH, LW, yy, xx) * w.w00 + der(L.dx, LH, LW, yy, xx + 1) * w.w01 +
                                               der(L.dx, LH, LW, yy + 1, xx) * w.w10 + der(L.dx, LH, LW, yy + 1, xx + 1) * w.w11,
                                           14);
                    const int gy = descale(der(L.dy, LH, LW, yy, xx) * w.w00 + der(L.dy, LH, LW, yy, xx + 1) * w.w01 +
                                               der(L.dy, LH, LW, yy + 1, xx) * w.w10 + der(L.dy, LH, LW, yy + 1, xx + 1) * w.w11,
                                           14);
                    Iw[y * win_w + x] = (int16_t)iv;
                    Ix[y * win_w + x] = (int16_t)gx;
                    Iy[y * win_w + x] = (int16_t)gy;
                    const float fx = (float)gx, fy = (float)gy;
                    if (x < vecA) {
                        qA[0][x & 3] = fx * fx + qA[0][x & 3];
                        qA[1][x & 3] = fx * fy + qA[1][x & 3];
                        qA[2][x & 3] = fy * fy + qA[2][x & 3];
                    } else {
                        tA[0] += (float)(gx * gx);
                        tA[1] += (float)(gx * gy);
                        tA[2] += (float)(gy * gy);
                    }
                }
            for (int k = 0; k < 3; k++) tA[k] += (qA[k][0] + qA[k][2]) + (qA[k][1] + qA[k][3]);  // v_reduce_sum
            const float A11 = tA[0] * FLT_SCALE, A12 = tA[1] * FLT_SCALE, A22 = tA[2] * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            const float minEig =
                (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win_w * win_h);
            if (err && get_min_eig) err[i] = minEig;
            if (minEig < (float)min_eig_threshold || D < FLT_EPSILON) {
                if (level == 0) status[i] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= halfx;
            ny -= halfy;
            float pdx = 0.f, pdy = 0.f;
            for (int j = 0; j < max_count; j++) {
                const int jx = (int)std::floor(nx), jy = (int)std::floor(ny);
                if (jx < -win_w || jx >= LW || jy < -win_h || jy >= LH) {
                    if (level == 0) status[i] = 0;
                    break;
                }
                w = weights(nx - jx, ny - jy);
                g_iterations++;
                float qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0}, tb1 = 0, tb2 = 0;
                const int vecB = (win_w / 8) * 8;
                for (int y = 0; y < win_h; y++) {
                    int pr[8][2];
                    for (int x = 0; x < win_w; x++) {
                        const int yy = jy + y, xx = jx + x;
                        const int diff = descale(pix(L.J, LH, LW, yy, xx) * w.w00 + pix(L.J, LH, LW, yy, xx + 1) * w.w01 +
                                                     pix(L.J, LH, LW, yy + 1, xx) * w.w10 + pix(L.J, LH, LW, yy + 1, xx + 1) * w.w11,
                                                 14 - 5) -
                                         Iw[y * win_w + x];
                        if (x < vecB) {
                            pr[x & 7][0] = diff * Ix[y * win_w + x];
                            pr[x & 7][1] = diff * Iy[y * win_w + x];
                            if ((x & 7) == 7) {
                                qb0[0] += (float)(pr[0][0] + pr[4][0]);
                                qb0[1] += (float)(pr[0][1] + pr[4][1]);
                                qb0[2] += (float)(pr[1][0] + pr[5][0]);
                                qb0[3] += (float)(pr[1][1] + pr[5][1]);
                                qb1[0] += (float)(pr[2][0] + pr[6][0]);
                                qb1[1] += (float)(pr[2][1] + pr[6][1]);
                                qb1[2] += (float)(pr[3][0] + pr[7][0]);
                                qb1[3] += (float)(pr[3][1] + pr[7][1]);
                            }
                        } else {
                            tb1 += (float)(diff * Ix[y * win_w + x]);
                            tb2 += (float)(diff * Iy[y * win_w + x]);
                        }
                    }
                }
                tb1 += (qb0[0] + qb1[0]) + (qb0[2] + qb1[2]);
                tb2 += (qb0[1] + qb1[1]) + (qb0[3] + qb1[3]);
                const float b1 = tb1 * FLT_SCALE, b2 = tb2 * FLT_SCALE;
                const float ddx = (float)((A12 * b2 - A22 * b1) * D), ddy = (float)((A12 * b1 - A11 * b2) * D);
                nx += ddx;
                ny += ddy;
                next_xy[2 * i] = nx + halfx;
                next_xy[2 * i + 1] = ny + halfy;
                if ((double)ddx * ddx + (double)ddy * ddy <= epsilon) break;
                if (j > 0 && std::abs(ddx + pdx) < 0.01 && std::abs(ddy + pdy) < 0.01) {
                    next_xy[2 * i] -= ddx * 0.5f;
                    next_xy[2 * i + 1] -= ddy * 0.5f;
                    break;
                }
                pdx = ddx;
                pdy = ddy;
            }
            if (status[i] && err && level == 0 && !get_min_eig) {
                const float qx = next_xy[2 * i] - halfx, qy = next_xy[2 * i + 1] - halfy;
                const int jx = (int)std::floor(qx), jy = (int)std::floor(qy);
                if (jx < -win_w || jx >= LW || jy < -win_h || jy >= LH) {
                    status[i] = 0;
                    continue;
                }
                w = weights(qx - jx, qy - jy);
                long long e = 0;
                for (int y = 0; y < win_h; y++)
                    for (int x = 0; x < win_w; x++) {
                        const int yy = jy + y, xx = jx + x;
                        const int diff = descale(pix(L.J, LH, LW, yy, xx) * w.w00 + pix(L.J, LH, LW, yy, xx + 1) * w.w01 +
                                                     pix(L.J, LH, LW, yy + 1, xx) * w.w10 + pix(L.J, LH, LW, yy + 1, xx + 1) * w.w11,
                                                 14 - 5) -
                                         Iw[y * win_w + x];
                        e += std::abs(diff);
                    }
                err[i] = (float)e * 1.f / (32 * win_w * win_h);
            }
        }
    }
    return levels;
}

}  // extern "C"
