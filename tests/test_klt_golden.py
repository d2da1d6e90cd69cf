"""Sparse pyramidal Lucas-Kanade (cv::calcOpticalFlowPyrLK at src/LoopHandler.cc:372-375) against
outputs of the real OpenCV (cv2 4.13.0) stored in tests/golden/klt_golden.npz:
  * CPU: the oracle reproduces pyramid levels, Scharr planes, tracked positions, status flags and err BIT FOR BIT;
  * GPU: the CUDA path reproduces the same fixtures through the C ABI (tests/test_gpu_klt.py adds oracle-vs-CUDA
    runs on seeded inputs the fixture does not hold).
err is compared where status == 1 only (OpenCV leaves it uninitialised elsewhere).
"""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "klt_golden.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def images(gold, kitti):
    from ya_vo_b200 import synth
    kitti2 = synth.shifted_pair(kitti, 2)
    assert sha(kitti2) == str(gold["kitti2_sha"])  # the seeded second frame is reproducible
    return {"kitti": (kitti, kitti2), "small": (gold["small1"], gold["small2"]), "tiny": (gold["tiny1"], gold["tiny2"])}


def case_args(gold, key):
    w, h, lv, ct, mc, eps, flags, me = gold[key + "_cfg"]
    init = gold[key + "_init"] if (key + "_init") in gold.files else None
    return dict(win=(int(w), int(h)), max_level=int(lv), crit_type=int(ct), max_count=int(mc), epsilon=float(eps),
                flags=int(flags), min_eig=float(me), init_pts=init)


def check_case(gold, key, nxt, st, er):
    g_st = gold[key + "_status"]
    assert np.array_equal(st, g_st), key
    ok = g_st == 1
    assert np.array_equal(nxt.view(np.uint32), gold[key + "_next"].view(np.uint32)), key  # every point, tracked or not
    assert np.array_equal(er[ok].view(np.uint32), gold[key + "_err"][ok].view(np.uint32)), key


def test_fixture_holds_the_reference_call(gold):
    names = [str(n) for n in gold["names"]]
    assert "kitti_fast_c0" in names and len(names) >= 24
    assert list(gold["kitti_fast_c0_cfg"][:5]) == [11, 11, 3, 3, 30]  # Size(11,11), maxLevel 3, COUNT+EPS, 30
    assert gold["kitti_fast_c0_pts"].shape == (2000, 2)
    assert 0 < gold["kitti_rand_c0_status"].mean() < 1  # both outcomes present


def test_oracle_pyr_down_and_scharr_match_opencv(oracle, gold, images):
    lv = images["kitti"][0]
    for k in range(1, 5):
        lv = oracle.pyr_down(lv)
        assert list(lv.shape) == list(gold["kitti_pyr%d_shape" % k])
        assert sha(lv) == str(gold["kitti_pyr%d_sha" % k])
    dx, dy = oracle.scharr(images["kitti"][0])
    assert sha(dx) == str(gold["kitti_scharr_dx_sha"]) and sha(dy) == str(gold["kitti_scharr_dy_sha"])
    for name in ("small1", "tiny1", "row", "col", "px2"):  # includes 1-pixel-high / -wide inputs
        im = gold["img_" + name]
        assert np.array_equal(oracle.pyr_down(im), gold["pyr_" + name]), name
        dx, dy = oracle.scharr(im)
        assert np.array_equal(dx, gold["dx_" + name]) and np.array_equal(dy, gold["dy_" + name]), name


def test_oracle_tracks_bit_exactly_like_opencv(oracle, gold, images):
    for key in [str(n) for n in gold["names"]]:
        a, b = images[key.split("_")[0]]
        nxt, st, er = oracle.klt_track(a, b, gold[key + "_pts"], **case_args(gold, key))
        check_case(gold, key, nxt, st, er)


def test_oracle_pyramid_depth_rule(oracle):
    # levels stop when a side of the next level would be <= the window side (cv::buildOpticalFlowPyramid)
    assert oracle.klt_levels(376, 1241, (11, 11), 3) == 3
    assert oracle.klt_levels(376, 1241, (11, 11), 10) == 5   # 376 -> 188 -> 94 -> 47 -> 24 -> 12 -> (6 <= 11)
    assert oracle.klt_levels(30, 41, (11, 11), 3) == 1       # 15 x 21 ok, 8 x 11 not
    assert oracle.klt_levels(20, 20, (11, 11), 3) == 0


def test_empty_point_list(oracle, gold):
    nxt, st, er = oracle.klt_track(gold["small1"], gold["small2"], np.zeros((0, 2), np.float32))
    assert nxt.shape == (0, 2) and st.size == 0 and er.size == 0


def random_cases(n_trials=24, seed=4242):
    """Seeded random frames, windows (3..31), depths, criteria and flags for the live comparisons with cv2."""
    from ya_vo_b200 import synth
    rng = np.random.default_rng(seed)
    for trial in range(n_trials):
        H, W = int(rng.integers(24, 200)), int(rng.integers(24, 260))
        kind = ("U", "G30", "B4")[trial % 3]
        a = synth.synth_frame(kind, 9000 + trial, H, W)
        b = synth.shifted_pair(a, 9100 + trial, drow=int(rng.integers(-2, 3)), dcol=int(rng.integers(-3, 4)))
        win = (int(rng.integers(1, 16)) * 2 + 1 if trial % 4 else int(rng.integers(3, 32)), int(rng.integers(3, 32)))
        lv = int(rng.integers(0, 5))
        ct = int(rng.integers(1, 4))
        mc, eps = int(rng.integers(0, 40)), float(rng.choice([0.0, 0.001, 0.01, 0.03, 0.3]))
        flags = int(rng.choice([0, 4, 8, 12]))
        me = float(rng.choice([1e-4, 1e-3, 1e-2]))
        n = 150
        pts = np.stack([rng.uniform(-10, W + 10, n), rng.uniform(-10, H + 10, n)], 1).astype(np.float32)
        init = (pts + rng.normal(0, 1.0, pts.shape)).astype(np.float32) if flags & 4 else None
        yield a, b, pts, init, dict(win=win, max_level=lv, crit_type=ct, max_count=mc, epsilon=eps, flags=flags, min_eig=me)


def live_opencv(cv2, a, b, pts, init, kw):
    nxt, st, er = cv2.calcOpticalFlowPyrLK(a, b, pts.copy(), None if init is None else init.copy(), winSize=kw["win"],
                                           maxLevel=kw["max_level"], criteria=(kw["crit_type"], kw["max_count"], kw["epsilon"]),
                                           flags=kw["flags"], minEigThreshold=kw["min_eig"])
    return nxt.astype(np.float32), st.ravel(), er.ravel()


def test_oracle_matches_live_opencv_on_random_configurations(oracle):
    """Where cv2 imports (this image and the GPU box's), the oracle is also compared with the live
    cv2.calcOpticalFlowPyrLK on seeded random frames, windows, depths, criteria and flags: bit for bit."""
    cv2 = pytest.importorskip("cv2")
    for a, b, pts, init, kw in random_cases():
        c_next, c_st, c_err = live_opencv(cv2, a, b, pts, init, kw)
        o_next, o_st, o_err = oracle.klt_track(a, b, pts, init_pts=init, **kw)
        assert np.array_equal(o_st, c_st), kw
        assert np.array_equal(o_next.view(np.uint32), c_next.view(np.uint32)), kw
        ok = o_st == 1
        assert np.array_equal(o_err[ok].view(np.uint32), c_err[ok].view(np.uint32)), kw
