"""Inlier count of the reference's F-matrix RANSAC (src/3DHandler.cc:163-190) against residuals computed
by the real OpenCV (cv2.gemm, tests/golden/epipolar_golden.npz): oracle on the CPU, CUDA path through the C ABI."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "epipolar_golden.npz")
INT_MIN = -2**31


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def check(gold, counts, best, bc, res):
    assert np.array_equal(res.view(np.uint64), gold["residuals"].view(np.uint64))   # doubles, bit for bit
    assert np.array_equal(counts, gold["counts"])
    assert best == int(np.argmax(gold["counts"])) and bc == int(gold["counts"].max())


def test_oracle_reproduces_opencv_residuals(oracle, gold):
    counts, best, bc, res = oracle.epipolar_inliers(gold["F"], gold["x1"], gold["y1"], gold["x2"], gold["y2"], float(gold["threshold"]), True)
    check(gold, counts, best, bc, res)
    assert 0 < gold["counts"].max() < gold["x1"].size and (gold["counts"] == 0).any()


def test_oracle_first_maximum_and_empty_inputs(oracle, gold):
    F = np.stack([gold["F"][2], gold["F"][0], gold["F"][2]])   # equal maxima at 0 and 2: the first wins (strict >)
    counts, best, bc = oracle.epipolar_inliers(F, gold["x1"], gold["y1"], gold["x2"], gold["y2"])
    assert counts[0] == counts[2] > counts[1] and best == 0
    counts, best, bc = oracle.epipolar_inliers(F, gold["x1"][:0], gold["y1"][:0], gold["x2"][:0], gold["y2"][:0])
    assert list(counts) == [0, 0, 0] and best == 0 and bc == 0
    counts, best, bc = oracle.epipolar_inliers(np.zeros((0, 9)), gold["x1"], gold["y1"], gold["x2"], gold["y2"])
    assert counts.size == 0 and best == -1 and bc == INT_MIN  # the reference's maxInliers = INT_MIN start


@pytest.mark.gpu
def test_cuda_reproduces_opencv_residuals(cuda_lib, gold):
    with cuda_lib.Context(device=0, n_slots=1, max_rows=64, max_cols=64, max_kp=64) as ctx:
        counts, best, bc, res = ctx.epipolar_inliers(gold["F"], gold["x1"], gold["y1"], gold["x2"], gold["y2"], float(gold["threshold"]), True)
        check(gold, counts, best, bc, res)
        F = np.stack([gold["F"][2], gold["F"][0], gold["F"][2]])
        counts, best, bc = ctx.epipolar_inliers(F, gold["x1"], gold["y1"], gold["x2"], gold["y2"])
        assert counts[0] == counts[2] > counts[1] and best == 0
        counts, best, bc = ctx.epipolar_inliers(np.zeros((0, 9)), gold["x1"], gold["y1"], gold["x2"], gold["y2"])
        assert counts.size == 0 and best == -1 and bc == INT_MIN
        counts, best, bc = ctx.epipolar_inliers(F, gold["x1"][:0], gold["y1"][:0], gold["x2"][:0], gold["y2"][:0])
        assert list(counts) == [0, 0, 0] and best == 0 and bc == 0


@pytest.mark.gpu
def test_cuda_matches_oracle_at_the_reference_scale(cuda_lib, oracle):
    """400 candidate matrices x 2000 matches (src/LoopHandler.cc:225: getFRANSAC(filterMatches, F, 400, 0.1))."""
    rng = np.random.default_rng(9)
    n, m = 2000, 400
    x1 = rng.integers(0, 376, n).astype(np.int32)
    y1 = rng.integers(0, 1241, n).astype(np.int32)
    x2 = np.clip(x1 + rng.integers(-3, 4, n), 0, 375).astype(np.int32)
    y2 = np.clip(y1 + rng.integers(-6, 7, n), 0, 1240).astype(np.int32)
    F = rng.normal(0, 1, (m, 3, 3)) * 10.0 ** rng.uniform(-7, -2, (m, 1, 1))
    F[:, 2, 2] = 1.0
    with cuda_lib.Context(device=0, n_slots=1, max_rows=64, max_cols=64, max_kp=64) as ctx:
        g = ctx.epipolar_inliers(F, x1, y1, x2, y2, 0.1, True)
    o = oracle.epipolar_inliers(F, x1, y1, x2, y2, 0.1, True)
    assert np.array_equal(g[3].view(np.uint64), o[3].view(np.uint64))
    assert np.array_equal(g[0], o[0]) and g[1] == o[1] and g[2] == o[2]
