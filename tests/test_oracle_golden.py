"""CPU: the oracle against the committed golden fixtures and the reference's known answers."""
import hashlib
import os

import numpy as np
import pytest

from ya_vo_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_ring_matches_reference_golden_image(oracle):
    """tests/FastDetectorTest.cc:6-31 BresenhamCircleCheck + tests/ImageTest.cc:23-37."""
    bres = np.load(os.path.join(GOLDEN, "bresenham_50x50.npy"))
    pts = oracle.ring(25, 25)
    assert pts.shape == (16, 2)
    painted = np.zeros((50, 50), np.uint8)
    for x, y in pts:
        painted[y, x] = 255  # putPixel: rawImage.at(pt) -> row y, col x
    lit = set(map(tuple, np.argwhere(bres > 0).tolist())) - {(25, 25)}
    assert set(map(tuple, np.argwhere(painted > 0).tolist())) == lit
    for x, y in pts:  # GetPixelMethod: getPixelVal(p.x, p.y) == 255 on the golden image
        assert bres[x, y] == 255
    n, lit_pts = oracle.ring_literal(25, 25)
    assert n == 16 and np.array_equal(lit_pts, pts)
    for xc, yc in [(4, 4), (100, 7), (9, 300)]:
        n, l = oracle.ring_literal(xc, yc)
        assert n == 16 and np.array_equal(l, oracle.ring(xc, yc))


def test_check_contiguous_known_answers(oracle):
    """tests/FastDetectorTest.cc:38-80."""
    assert oracle.check_contiguous(0, [255] * 16) is True
    assert oracle.check_contiguous(255, [255] * 16) is False
    assert oracle.check_contiguous(0, [255] * 11 + [0] * 5) is False
    assert oracle.check_contiguous(0, [255] * 12 + [0] * 4) is True
    assert oracle.check_contiguous(0, [0] * 4 + [255] * 12) is True
    assert oracle.check_contiguous(0, [255] * 6 + [0] * 4 + [255] * 6) is False  # no wrap-around
    # threshold: |c-p| >= 40 differs
    assert oracle.check_in_between(100, 139) and not oracle.check_in_between(100, 140)
    assert oracle.check_in_between(100, 61) and not oracle.check_in_between(100, 60)


def test_blur_golden(oracle):
    g = np.load(os.path.join(GOLDEN, "blur_golden.npz"))
    names = sorted(k[:-3] for k in g.files if k.endswith("_in") and "_sha_" not in k)
    assert len(names) >= 6
    for n in names:
        assert np.array_equal(oracle.gaussian_blur(g[n + "_in"]), g[n + "_out"]), n
    for name, img in (("U_seed0", synth.synth_frame("U", 0)), ("G30_seed1", synth.synth_frame("G30", 1)),
                      ("B4_seed2", synth.synth_frame("B4", 2))):
        assert sha(img) == str(g[name + "_sha_in"]), "synthetic generator drifted: " + name
        assert sha(oracle.gaussian_blur(img)) == str(g[name + "_sha_out"]), name


def test_blur_golden_kitti(oracle, kitti):
    g = np.load(os.path.join(GOLDEN, "blur_golden.npz"))
    assert sha(kitti) == str(g["kitti_sha_in"])
    assert sha(oracle.gaussian_blur(kitti)) == str(g["kitti_sha_out"])


def test_eigen_and_score_golden(oracle):
    g = np.load(os.path.join(GOLDEN, "eigen_golden.npz"))
    t, eig, score = g["tensors"], g["eig"], g["score"]
    assert t.shape[0] > 4000
    for i in range(t.shape[0]):
        l1, l2 = oracle.eigen2x2(*t[i])
        assert l1.tobytes() == eig[i, 0].tobytes() and l2.tobytes() == eig[i, 1].tobytes(), t[i]
        assert oracle.score_from_tensor(*t[i]).tobytes() == score[i].tobytes(), t[i]


def test_candidate_counts_match_survey_probes(oracle, kitti):
    """3,791 candidates on the KITTI fixture, 17,205 on uniform noise seed 0."""
    r, c, s = oracle.fast_candidates(kitti)
    assert r.size == 3791
    r, c, s = oracle.fast_candidates(synth.synth_frame("U", 0))
    assert r.size == 17205
    assert r.min() >= 4 and r.max() <= 376 - 5 and c.min() >= 4 and c.max() <= 1241 - 5
    lin = r.astype(np.int64) * 1241 + c
    assert np.all(np.diff(lin) > 0)  # scan order


def test_harris_paths_agree(oracle, kitti):
    r, c, s = oracle.fast_candidates(kitti)
    for i in range(0, r.size, 97):
        assert oracle.harris(kitti, r[i], c[i]).tobytes() == s[i].tobytes()


def test_sobel_loop_bounds(oracle):
    img = synth.synth_frame("U", 3, 20, 24)
    ix, iy = oracle.sobel(img)
    assert not ix[-2:].any() and not ix[:, -2:].any() and not iy[-2:].any() and not iy[:, -2:].any()
    p = img.astype(np.int32)
    r, c = 7, 9
    gx = (p[r - 1, c + 1] - p[r - 1, c - 1]) + 2 * (p[r, c + 1] - p[r, c - 1]) + (p[r + 1, c + 1] - p[r + 1, c - 1])
    assert ix[r, c] == gx
    assert ix[0, 0] == 2 * p[0, 1] + p[1, 1]  # zero padding


def test_fast_detect_is_std_sort_of_candidates(oracle, kitti):
    r, c, s = oracle.fast_candidates(kitti)
    kr, kc, ks, nc = oracle.fast_detect(kitti, 2000)
    assert nc == r.size and kr.size == 2000
    ss, pp = oracle.std_sort_desc(s, np.arange(s.size))
    assert np.array_equal(kr, r[pp[:2000]]) and np.array_equal(kc, c[pp[:2000]])
    assert np.all(np.diff(ks) <= 0)
    # fewer candidates than the cap
    small = synth.synth_frame("U", 9, 40, 60)
    kr, kc, ks, nc = oracle.fast_detect(small, 2000)
    assert kr.size == nc
    flat = np.full((30, 30), 7, np.uint8)
    kr, kc, ks, nc = oracle.fast_detect(flat, 2000)
    assert nc == 0 and kr.size == 0


def test_introsort_restatement_equals_std_sort(oracle):
    rng = np.random.default_rng(1)
    for t in range(60):
        n = int(rng.integers(1, 5000))
        s = rng.integers(0, int(rng.integers(2, 400)), n).astype(np.float32)
        k = int(rng.integers(1, n + 10))
        a, p = oracle.std_sort_desc(s, np.arange(n))
        b, q = oracle.introsort_topk(s, np.arange(n), k)
        assert np.array_equal(p[:k], q[:k])


def test_brief_semantics(oracle, offsets):
    img = synth.synth_frame("U", 4, 64, 80)
    H, W = img.shape
    blurred = oracle.gaussian_blur(img)
    rows = np.array([8, 7, 30, H - 8, H - 9, 30, 30, 56], np.int32)
    cols = np.array([8, 30, 7, 30, 30, W - 8, W - 7, 72], np.int32)
    desc, valid, oob = oracle.brief(img, offsets, rows, cols)
    assert valid.tolist() == [True, False, False, True, True, True, False, True]
    assert not desc[~valid].any()
    # bit j of keypoint 0 by the definition in src/BriefDescriptor.cc:98-118
    flat = blurred.reshape(-1)
    for j in (0, 1, 17, 100, 255):
        a = flat[(8 + offsets[j, 0]) * W + 8 + offsets[j, 1]]
        b = flat[(8 + offsets[j, 2]) * W + 8 + offsets[j, 3]]
        assert ((desc[0, j // 8] >> (j % 8)) & 1) == int(a > b)
    # row + 8 == H reads past the buffer in the reference; counted, defined as 0
    assert oob >= 1
    # col + 8 == W wraps to column 0 of the next row (linear indexing)
    j = int(np.argmax(offsets[:, 1] == 8))
    a = flat[(30 + offsets[j, 0]) * W + (W - 8) + 8]
    b = flat[(30 + offsets[j, 2]) * W + (W - 8) + offsets[j, 3]]
    assert ((desc[5, j // 8] >> (j % 8)) & 1) == int(a > b)


def test_match_semantics(oracle):
    d1 = synth.synth_descriptors(50, 1)
    d2 = synth.synth_descriptors(70, 2)
    d2[10] = d1[3]
    d2[40] = d1[3]  # duplicate minimum: lowest index wins
    idx, dist, sec, rev = oracle.match(d1, d2, extensions=True)
    assert idx[3] == 10 and dist[3] == 0 and sec[3] == 0
    full = np.unpackbits(d1[:, None, :] ^ d2[None, :, :], axis=2).sum(axis=2)
    assert np.array_equal(idx, full.argmin(axis=1)) and np.array_equal(dist, full.min(axis=1))
    assert np.array_equal(rev, full.argmin(axis=0))
    assert np.array_equal(sec, np.sort(full, axis=1)[:, 1])
    assert oracle.hamming(d1[0], d2[0]) == full[0, 0]
    assert [oracle.popcount(v) for v in (0, 1, 255, 0x5A)] == [0, 1, 8, 4]
    idx, dist = oracle.match(d1, d2[:0])
    assert np.all(idx == -1) and np.all(dist == 2**31 - 1)
    keep = oracle.remove_outliers(np.array([5, 9, 10, 19, 20, 40]), 20)
    assert keep.tolist() == [True, True, True, True, False, False]
    keep = oracle.remove_outliers(np.array([15, 29, 30, 31]), 20)
    assert keep.tolist() == [True, True, False, False]
