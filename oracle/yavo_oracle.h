/*
 * yavo_oracle.h — CPU ORACLE for the YA_VO ORB-style front end.
 *
 * TEST INFRASTRUCTURE ONLY.  This library restates, on the CPU, the algorithm
 * of the reference's hot path (FastDetector / Brief / Image).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  Nothing under ya_vo_b200/ links, imports or calls it.
 *
 * Parity status: the two OpenCV-owned steps (8-bit GaussianBlur 9x9 sigma 2.5
 * and cv::eigen on a 2x2 float32 matrix) are pinned against cv2 4.13.0 through
 * the committed fixtures in tests/golden/ (made by tests/golden/make_golden.py);
 * the reference-owned control flow is pinned against the reference's own
 * sources compiled unmodified over a stub OpenCV tree (oracle/ref_shim,
 * output oracle/_ref/) and against the reference's known-answer tests
 * (tests/FastDetectorTest.cc:6-80, tests/ImageTest.cc:23-37).
 *
 * Coordinate convention (reference): Point.x = row, Point.y = col.
 */
#ifndef YAVO_ORACLE_H
#define YAVO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/FastDetector.cc:50-112 — 16 ring points around (xc,yc); out_xy = x0,y0,x1,y1,... */
void yavo_oracle_ring(int xc, int yc, int32_t *out_xy);
/* literal restatement of the std::set based generator (same lines), for cross-checking the table */
int yavo_oracle_ring_literal(int xc, int yc, int32_t *out_xy);

/* src/FastDetector.cc:155-161 */
int yavo_oracle_check_in_between(uint8_t cent, uint8_t cond);
/* src/FastDetector.cc:135-153 — ring values in ring order */
int yavo_oracle_check_contiguous(uint8_t cent, const uint8_t ring_vals[16]);

/* src/FastDetector.cc:164-214 — Sobel planes with the reference's loop bounds (last two rows/cols stay 0) */
void yavo_oracle_sobel(const uint8_t *img, int H, int W, float *Ix, float *Iy);
/* OpenCV JacobiImpl_<float> for a symmetric 2x2 [a b; b c]; l1 >= l2 */
void yavo_oracle_eigen2x2(float a, float b, float c, float *l1, float *l2);
/* src/FastDetector.cc:270 applied to eigenvalues of [a b; b c] */
float yavo_oracle_score_from_tensor(float a, float b, float c);
/* src/FastDetector.cc:244-273 at pixel (x=row, y=col), x in [1,H-2], y in [1,W-2] */
float yavo_oracle_harris(const uint8_t *img, int H, int W, int x, int y);

/* src/FastDetector.cc:298-335 — candidates in scan (row-major) order. Returns N_cand
 * (may exceed cap; only the first cap are written). */
int yavo_oracle_fast_candidates(const uint8_t *img, int H, int W, int cap,
                                int32_t *rows, int32_t *cols, float *scores);
/* src/FastDetector.cc:277-369 — std::sort (libstdc++) by score desc, first max_kp.
 * Returns the number written; *n_cand receives the candidate count. */
int yavo_oracle_fast_detect(const uint8_t *img, int H, int W, int max_kp,
                            int32_t *rows, int32_t *cols, float *scores, int *n_cand);
/* std::sort replay on an explicit (score, payload) list, same comparator as :343-345 */
void yavo_oracle_std_sort_desc(float *scores, int32_t *payload, int n);
/* transparent restatement of libstdc++'s introsort restricted to the ranges that
 * decide the first k outputs (the model of the CUDA select kernel) */
void yavo_oracle_introsort_topk(float *scores, int32_t *payload, int n, int k);
/* same with an explicit initial depth limit (forces the std::make_heap/std::sort_heap fallback) */
void yavo_oracle_introsort_topk_depth(float *scores, int32_t *payload, int n, int k, int depth);

/* cv::GaussianBlur(u8, 9x9, 2.5) as OpenCV 4.x computes it (fixed point), BORDER_REFLECT_101 */
void yavo_oracle_gaussian_blur(const uint8_t *img, int H, int W, uint8_t *out);

/* src/BriefDescriptor.cc:86-136. offsets = 256x4 int32 {drow1,dcol1,drow2,dcol2}.
 * desc = n x 32 bytes, valid = n bytes (checkBoundry), *n_oob = keypoints that read
 * at a linear index >= H*W (reference UB; defined here as pixel value 0).
 * `blurred` may be NULL (then it is computed from img). */
void yavo_oracle_brief(const uint8_t *img, const uint8_t *blurred, int H, int W,
                       const int32_t *offsets, const int32_t *rows, const int32_t *cols,
                       int n, uint8_t *desc, uint8_t *valid, int *n_oob);

/* src/BriefDescriptor.cc:139-160 */
int yavo_oracle_popcount(uint8_t v);
int yavo_oracle_hamming(const uint8_t *a, const uint8_t *b);
/* src/BriefDescriptor.cc:163-183 — first minimum wins. n2 == 0 -> idx -1, dist INT_MAX.
 * Extensions (no reference counterpart, F9): second = second smallest distance
 * (INT_MAX if n2 < 2), rev_idx[j] = first i minimising d(i,j) (may be NULL). */
void yavo_oracle_match(const uint8_t *d1, int n1, const uint8_t *d2, int n2,
                       int32_t *idx, int32_t *dist, int32_t *second, int32_t *rev_idx);
/* src/BriefDescriptor.cc:213-231 — keep[i] = dist[i] < max(2*min, threshold). returns kept count */
int yavo_oracle_remove_outliers(const int32_t *dist, int n, int threshold, uint8_t *keep);

/* whole front end on F frames (frame-parallel over nthreads std::threads): FAST + BRIEF on
 * each, and (if do_match) match frame f-1 -> f.  Outputs may be NULL (timing only).
 * kp_* are F x max_kp; n_kp is F; match_* are F x max_kp (row 0 unused). Returns 0. */
int yavo_oracle_pipeline(const uint8_t *frames, int F, int H, int W, const int32_t *offsets,
                         int max_kp, int do_match, int nthreads,
                         int32_t *kp_rows, int32_t *kp_cols, float *kp_scores,
                         uint8_t *kp_desc, int32_t *n_kp,
                         int32_t *match_idx, int32_t *match_dist);

/* ---- sparse pyramidal Lucas-Kanade (yavo_oracle_klt.cpp) ----------------------------------
 * cv::calcOpticalFlowPyrLK as src/LoopHandler.cc:372-375 calls it, restated for 8-bit single-channel images.
 * Points are OpenCV's (x = column, y = row), interleaved x0,y0,x1,y1,...  crit_type: 1 = COUNT, 2 = EPS.
 * flags: 4 = OPTFLOW_USE_INITIAL_FLOW (next_xy is read), 8 = OPTFLOW_LK_GET_MIN_EIGENVALS.
 * err is meaningful only where status == 1 (OpenCV leaves the rest uninitialised; here 0 or the last value).
 * Returns the top pyramid level actually used. */
void yavo_oracle_pyr_down(const uint8_t *img, int H, int W, uint8_t *out /* ((H+1)/2) x ((W+1)/2) */);
void yavo_oracle_scharr(const uint8_t *img, int H, int W, int16_t *dx, int16_t *dy);
int yavo_oracle_klt_levels(int H, int W, int win_w, int win_h, int max_level);
long long yavo_oracle_klt_iterations(void); /* Newton steps taken by the last yavo_oracle_klt call */
int yavo_oracle_klt(const uint8_t *prev, const uint8_t *next, int H, int W, const float *prev_xy, int n, float *next_xy,
                    uint8_t *status, float *err, int win_w, int win_h, int max_level, int crit_type, int max_count,
                    double epsilon, int flags, double min_eig_threshold);

/* ---- inlier count of the reference's F-matrix RANSAC (yavo_oracle_geom.cpp) ---------------------
 * src/3DHandler.cc:163-188: residual = p2.t() * F * p1 with p = (x, y, 1) from the integer keypoint coordinates,
 * inlier iff fabs(residual) < threshold, first maximum over the m candidate matrices wins.  F: m x 9 row-major. */
double yavo_oracle_epipolar_residual(const double *F, int x1, int y1, int x2, int y2);
void yavo_oracle_epipolar_inliers(const double *F, int m, const int32_t *x1, const int32_t *y1, const int32_t *x2,
                                  const int32_t *y2, int n, double threshold, int32_t *counts, double *residuals,
                                  int32_t *best, int32_t *best_count);

#ifdef __cplusplus
}
#endif
#endif
