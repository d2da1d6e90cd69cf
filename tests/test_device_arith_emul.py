"""CPU: the per-thread device arithmetic (fast_core.h, select_serial.h), built for the host,
against the oracle — byte-SIMD segment test, dp4a/dp2a Gaussian, Harris response, std::sort replay."""
import ctypes as C
import os

import numpy as np
import pytest

from ya_vo_b200 import synth


def p(a):
    return a.ctypes.data_as(C.c_void_p)


CASES = [("U", 0, 376, 1241), ("G30", 1, 376, 1241), ("B4", 2, 376, 1241), ("U", 5, 37, 50), ("U", 6, 41, 133),
         ("U", 7, 9, 9), ("U", 8, 12, 300)]


@pytest.mark.parametrize("kind,seed,H,W", CASES)
def test_swar_segment_test_and_harris(emul, oracle, kind, seed, H, W):
    img = synth.synth_frame(kind, seed, H, W)
    m = np.zeros((H, W), np.uint8)
    emul.emul_fast_mask(p(img), H, W, p(m))
    r, c, s = oracle.fast_candidates(img)
    rr, cc = np.nonzero(m)
    assert np.array_equal(rr, r) and np.array_equal(cc, c)
    step = max(1, r.size // 1500)
    for i in range(0, r.size, step):
        got = np.float32(emul.emul_harris(p(img), W, int(r[i]), int(c[i])))
        assert got.tobytes() == s[i].tobytes()


def test_kitti_frame(emul, oracle, kitti):
    H, W = kitti.shape
    m = np.zeros((H, W), np.uint8)
    emul.emul_fast_mask(p(kitti), H, W, p(m))
    r, c, s = oracle.fast_candidates(kitti)
    rr, cc = np.nonzero(m)
    assert np.array_equal(rr, r) and np.array_equal(cc, c)
    b = np.zeros((H, W), np.uint8)
    emul.emul_blur(p(kitti), H, W, p(b))
    assert np.array_equal(b, oracle.gaussian_blur(kitti))


@pytest.mark.parametrize("kind,seed,H,W", CASES)
def test_dp4a_dp2a_blur(emul, oracle, kind, seed, H, W):
    img = synth.synth_frame(kind, seed, H, W)
    b = np.zeros((H, W), np.uint8)
    emul.emul_blur(p(img), H, W, p(b))
    assert np.array_equal(b, oracle.gaussian_blur(img))


def test_score_golden_through_device_arith(emul):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eigen_golden.npz"))
    t, score = g["tensors"], g["score"]
    for i in range(t.shape[0]):
        got = np.float32(emul.emul_score_from_tensor(int(t[i, 0]), int(t[i, 1]), int(t[i, 2])))
        assert got.tobytes() == score[i].tobytes(), t[i]


def test_select_model_equals_std_sort(emul, oracle):
    rng = np.random.default_rng(0)
    for t in range(150):
        n = int(rng.integers(1, 6000))
        s = rng.integers(0, max(2, int(rng.integers(2, 300))), n).astype(np.float32)
        if t % 5 == 0:
            s = rng.standard_normal(n).astype(np.float32)
        if t % 7 == 0:
            s = np.sort(s)  # adversarial orders
        if t % 11 == 0:
            s = np.sort(s)[::-1].copy()
        pay = np.arange(n, dtype=np.int32)
        K = int(rng.integers(1, n + 50))
        a1, b1 = oracle.std_sort_desc(s, pay)
        s2, p2 = s.copy(), pay.copy()
        emul.emul_select(p(s2), p(p2), n, K)
        assert np.array_equal(b1[:K], p2[:K]), (t, n, K)
        for depth in (-1, 0, 2):
            o_s, o_p = oracle.introsort_topk(s, pay, 1 << 30, depth)
            s3, p3 = s.copy(), pay.copy()
            emul.emul_serial_sort(p(s3), p(p3), n, 1 << 30, depth)
            assert np.array_equal(o_p, p3), (t, n, depth)


def test_necessary_condition_never_drops_a_corner(emul, kitti):
    """yavo_fast4_core (ring 0,4,7,8) must hold wherever the full segment test fires: the detect kernel only runs
    the full test on quads that pass it."""
    for img in (kitti, synth.synth_frame("U", 0), synth.synth_frame("G30", 1), synth.synth_frame("B4", 2)):
        H, W = img.shape
        core = C.c_longlong()
        full = C.c_longlong()
        bad = emul.emul_fast_core_violations(p(img), H, W, C.byref(core), C.byref(full))
        assert bad == 0
        assert full.value <= core.value
