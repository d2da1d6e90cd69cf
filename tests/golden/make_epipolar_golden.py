#!/usr/bin/env python
"""Generates tests/golden/epipolar_golden.npz: the algebraic epipolar residual p2.t() * F * p1 of the reference's
F-matrix RANSAC (src/3DHandler.cc:163-186) evaluated by the real OpenCV (cv2 4.13.0 cv2.gemm, the two products the
reference's cv::Mat expression performs), for 16 candidate matrices x 400 integer point pairs, plus the inlier counts
at the reference's threshold 0.1.  Run in the build container only (imports cv2).

    python tests/golden/make_epipolar_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(88)
    n, m = 400, 16
    x1 = rng.integers(0, 376, n).astype(np.int32)      # the reference's (x, y) = (row, col)
    y1 = rng.integers(0, 1241, n).astype(np.int32)
    x2 = np.clip(x1 + rng.integers(-3, 4, n), 0, 375).astype(np.int32)
    y2 = np.clip(y1 + rng.integers(-6, 7, n), 0, 1240).astype(np.int32)
    F = np.zeros((m, 3, 3))
    for i in range(m):
        # plausible fundamental matrices of a small sideways motion, scaled like the reference's (F[2][2] = 1),
        # with magnitudes that put a good share of the residuals near the 0.1 threshold
        t = rng.normal(0, 1, 3)
        K = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) * 10.0 ** rng.uniform(-5, -2)
        F[i] = K + rng.normal(0, 1e-7, (3, 3))
        F[i, 2, 2] = 1.0 if i % 2 else rng.normal(0, 1e-3)
    res = np.zeros((m, n))
    for i in range(m):
        for k in range(n):
            p1 = np.array([[x1[k]], [y1[k]], [1.0]])
            p2 = np.array([[x2[k]], [y2[k]], [1.0]])
            r = cv2.gemm(p2, F[i], 1, None, 0, flags=cv2.GEMM_1_T)
            res[i, k] = cv2.gemm(r, p1, 1, None, 0)[0, 0]
    counts = (np.abs(res) < 0.1).sum(1).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "epipolar_golden.npz"), F=F, x1=x1, y1=y1, x2=x2, y2=y2, residuals=res, counts=counts,
                        threshold=0.1)
    print("counts", counts)


if __name__ == "__main__":
    main()
