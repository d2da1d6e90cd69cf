// yavo_oracle_geom.cpp — CPU ORACLE for the inlier count of the reference's fundamental-matrix RANSAC (test
// infrastructure only; see yavo_oracle.h).  SURVEY 8f-4.
//
// _3DHandler::getFRANSAC (src/3DHandler.cc:145-195) draws 8 random matches per iteration (std::random_device —
// not reproducible), fits F with the normalised 8-point algorithm (cv::SVD, stays on the host) and counts the
// matches whose algebraic epipolar residual is below the threshold (:163-186):
//     p1 = (pt1.x, pt1.y, 1), p2 = (pt2.x, pt2.y, 1) as doubles from the integer keypoint coordinates;
//     error = p2.t() * F * p1;   inlier iff fabs(error) < threshold;   the first maximum over iterations wins (:187).
// The products are two cv::gemm calls; for these 1x3 * 3x3 and 1x3 * 3x1 shapes OpenCV sums the three products of a
// dot product left to right in double, no FMA (x86 baseline build).  That order is pinned bit for bit against
// cv2.gemm (4.13.0) by tests/golden/epipolar_golden.npz, made by tests/golden/make_epipolar_golden.py.
#include <climits>
#include <cmath>
#include <cstdint>

#include "yavo_oracle.h"

extern "C" {

double yavo_oracle_epipolar_residual(const double *F, int x1, int y1, int x2, int y2) {
    const double a[3] = {(double)x2, (double)y2, 1.0}, b[3] = {(double)x1, (double)y1, 1.0};
    double r[3];
    for (int j = 0; j < 3; j++) r[j] = (a[0] * F[j] + a[1] * F[3 + j]) + a[2] * F[6 + j];  // p2.t() * F
    return (r[0] * b[0] + r[1] * b[1]) + r[2] * b[2];                                       // (...) * p1
}

void yavo_oracle_epipolar_inliers(const double *F, int m, const int32_t *x1, const int32_t *y1, const int32_t *x2,
                                  const int32_t *y2, int n, double threshold, int32_t *counts, double *residuals,
                                  int32_t *best, int32_t *best_count) {
    int bi = -1, bc = INT_MIN;
    for (int i = 0; i < m; i++) {
        int c = 0;
        for (int k = 0; k < n; k++) {
            const double e = yavo_oracle_epipolar_residual(F + 9 * i, x1[k], y1[k], x2[k], y2[k]);
            if (residuals) residuals[(size_t)i * n + k] = e;
            if (std::fabs(e) < threshold) c++;
        }
        counts[i] = c;
        if (c > bc) {  // strict: the first maximum wins (src/3DHandler.cc:187)
            bc = c;
            bi = i;
        }
    }
    if (best) *best = bi;
    if (best_count) *best_count = bc;
}

}  // extern "C"
