"""GPU parity of the tracking step (cv::calcOpticalFlowPyrLK, src/LoopHandler.cc:372-375) through the
C ABI: bit-exact against the cv2 4.13 fixtures (tests/golden/klt_golden.npz) and against the CPU oracle on seeded
inputs the fixtures do not hold; size-independent properties at full frame size."""
import hashlib

import numpy as np
import pytest

from test_klt_golden import GOLDEN, case_args, check_case, live_opencv, random_cases

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def ctx(cuda_lib):
    with cuda_lib.Context(device=0, n_slots=6, max_rows=376, max_cols=1241, max_kp=2000) as c:
        yield c


@pytest.fixture(scope="module")
def images(gold, kitti):
    from ya_vo_b200 import synth
    return {"kitti": (kitti, synth.shifted_pair(kitti, 2)), "small": (gold["small1"], gold["small2"]),
            "tiny": (gold["tiny1"], gold["tiny2"])}


def test_cuda_pyramid_matches_opencv(ctx, gold, images, oracle):
    ctx.upload(0, images["kitti"][0])
    assert ctx.build_pyramid(0, 1, (11, 11), 4) == 4
    for k in range(1, 5):
        lv = ctx.pyramid_level(0, k)
        assert list(lv.shape) == list(gold["kitti_pyr%d_shape" % k])
        assert sha(lv) == str(gold["kitti_pyr%d_sha" % k]), k
    for name in ("small1", "tiny1"):
        im = gold["img_" + name]
        ctx.upload(1, im)
        assert ctx.build_pyramid(1, 1, (3, 3), 1) == 1
        assert np.array_equal(ctx.pyramid_level(1, 1), gold["pyr_" + name]), name
    # odd sizes down to a few pixels against the oracle (itself pinned to cv2.pyrDown)
    rng = np.random.default_rng(5)
    for shape in ((9, 13), (8, 8), (5, 64), (33, 7)):
        im = rng.integers(0, 256, shape, dtype=np.uint8)
        ctx.upload(1, im)
        n = ctx.build_pyramid(1, 1, (3, 3), 2)
        cur = im
        for k in range(1, n + 1):
            cur = oracle.pyr_down(cur)
            assert np.array_equal(ctx.pyramid_level(1, k), cur), (shape, k)


def test_cuda_tracks_bit_exactly_like_opencv(ctx, gold, images):
    loaded = None
    for key in [str(n) for n in gold["names"]]:
        fam = key.split("_")[0]
        if fam != loaded:
            ctx.upload(0, images[fam][0])
            ctx.upload(1, images[fam][1])
            loaded = fam
        nxt, st, er = ctx.klt_track(0, 1, gold[key + "_pts"], **case_args(gold, key))
        check_case(gold, key, nxt, st, er)


@pytest.mark.parametrize("kind,seed,shape", [("U", 21, (376, 1241)), ("G30", 22, (240, 320)), ("B4", 23, (97, 131)),
                                             ("G30", 24, (31, 45))])
def test_cuda_matches_oracle_on_seeded_frames(ctx, oracle, kind, seed, shape):
    from ya_vo_b200 import synth
    a = synth.synth_frame(kind, seed, *shape)
    b = synth.shifted_pair(a, seed + 100)
    ctx.upload(2, a)
    ctx.upload(3, b)
    rng = np.random.default_rng(seed)
    H, W = shape
    pts = np.stack([rng.uniform(-25, W + 25, 700), rng.uniform(-25, H + 25, 700)], 1).astype(np.float32)
    pts[:50] = np.round(pts[:50])              # integer positions (the reference tracks integer keypoints)
    pts[50:60] = [[0, 0]] * 10 + rng.uniform(-1, 1, (10, 2)).astype(np.float32) * 1e-3   # image corner, sub-pixel
    for win, lv, ct, mc, eps, flags, me in (((11, 11), 3, 3, 30, 0.01, 0, 1e-3), ((9, 13), 2, 3, 20, 0.02, 0, 1e-4),
                                            ((16, 8), 5, 1, 7, 0.0, 8, 1e-3), ((3, 3), 1, 2, 0, 0.1, 4, 1e-3)):
        init = (pts + rng.normal(0, 2, pts.shape)).astype(np.float32) if flags & 4 else None
        kw = dict(win=win, max_level=lv, crit_type=ct, max_count=mc, epsilon=eps, flags=flags, min_eig=me, init_pts=init)
        g_next, g_st, g_err = ctx.klt_track(2, 3, pts, **kw)
        o_next, o_st, o_err = oracle.klt_track(a, b, pts, **kw)
        assert np.array_equal(g_st, o_st), (win, lv)
        assert np.array_equal(g_next.view(np.uint32), o_next.view(np.uint32)), (win, lv)
        ok = o_st == 1
        assert np.array_equal(g_err[ok].view(np.uint32), o_err[ok].view(np.uint32)), (win, lv)
        assert 0 < ok.sum() < pts.shape[0]


def test_cuda_matches_live_opencv_on_random_configurations(ctx):
    """The CUDA path against the live cv2.calcOpticalFlowPyrLK of this image (the GPU box has the same cv2 4.13):
    seeded random frame sizes, windows 3..31, 0-4 levels, criteria and flags — bit for bit."""
    cv2 = pytest.importorskip("cv2")
    for a, b, pts, init, kw in random_cases(n_trials=32, seed=777):
        ctx.upload(2, a)
        ctx.upload(3, b)
        c_next, c_st, c_err = live_opencv(cv2, a, b, pts, init, kw)
        g_next, g_st, g_err = ctx.klt_track(2, 3, pts, init_pts=init, **kw)
        assert np.array_equal(g_st, c_st), kw
        assert np.array_equal(g_next.view(np.uint32), c_next.view(np.uint32)), kw
        ok = c_st == 1
        assert np.array_equal(g_err[ok].view(np.uint32), c_err[ok].view(np.uint32)), kw


def test_batch_form_tracks_the_frontend_keypoints(ctx, oracle, offsets):
    """yavo_frontend_batch leaves each frame's top-K keypoints on the device; yavo_klt_track_batch tracks them into
    the next frame with the reference's parameters (what LoopHandler::trackLastFrame feeds OpenCV)."""
    from ya_vo_b200 import synth
    f0 = synth.synth_frame("B4", 40, 200, 320)
    frames = np.stack([f0, synth.shifted_pair(f0, 41), synth.synth_frame("G30", 42, 200, 320), synth.synth_frame("B4", 43, 200, 320)])
    ctx.set_brief_offsets(offsets)
    ctx.upload_batch(0, frames)
    ctx.frontend_batch(0, 4, False)
    ctx.klt_track_batch(0, 4)
    xy, st, er = ctx.klt_fetch(0, 4)
    for f in range(3):
        r, c, s, nc = oracle.fast_detect(frames[f], 2000)
        pts = np.stack([c, r], 1).astype(np.float32)
        o_next, o_st, o_err = oracle.klt_track(frames[f], frames[f + 1], pts)
        k = r.size
        assert k > 100
        assert np.array_equal(st[f, :k], o_st)
        assert np.array_equal(xy[f, :k].view(np.uint32), o_next.view(np.uint32))
        ok = o_st == 1
        assert np.array_equal(er[f, :k][ok].view(np.uint32), o_err[ok].view(np.uint32))


def test_fullsize_batch_tracking_matches_oracle(cuda_lib, oracle, offsets, kitti):
    """BASELINE frame size (1241x376), 2000 keypoints per frame, 5 frames resident: the KITTI fixture followed by
    shifted / noisy successors and one unrelated frame (tracks that fail or wander must agree too)."""
    from ya_vo_b200 import synth
    frames = [kitti]
    for f in range(3):
        frames.append(synth.shifted_pair(frames[-1], 300 + f, drow=f - 1, dcol=2 * f + 1))
    frames.append(synth.synth_frame("G30", 305, 376, 1241))
    frames = np.stack(frames)
    with cuda_lib.Context(device=0, n_slots=5, max_rows=376, max_cols=1241, max_kp=2000) as c:
        c.set_brief_offsets(offsets)
        c.upload_batch(0, frames)
        c.frontend_batch(0, 5, False)
        c.klt_track_batch(0, 5)
        xy, st, er = c.klt_fetch(0, 5)
    total = 0
    for f in range(4):
        r, col, s, nc = oracle.fast_detect(frames[f], 2000)
        pts = np.stack([col, r], 1).astype(np.float32)
        o_next, o_st, o_err = oracle.klt_track(frames[f], frames[f + 1], pts)
        k = r.size
        assert np.array_equal(st[f, :k], o_st), f
        assert np.array_equal(xy[f, :k].view(np.uint32), o_next.view(np.uint32)), f
        ok = o_st == 1
        assert np.array_equal(er[f, :k][ok].view(np.uint32), o_err[ok].view(np.uint32)), f
        total += k
    assert total > 7000


def test_identical_frames_leave_integer_keypoints_in_place(ctx, kitti):
    """Size-independent property at full frame size: with prev == next every mismatch sum is zero, so each accepted
    point stays exactly where it was and its err is 0."""
    ctx.upload(4, kitti)
    ctx.upload(5, kitti)
    r, c, s, nc = ctx.fast_detect(4)
    pts = np.stack([c, r], 1).astype(np.float32)
    nxt, st, er = ctx.klt_track(4, 5, pts)
    ok = st == 1
    assert ok.sum() > 1500
    assert np.array_equal(nxt[ok], pts[ok]) and np.all(er[ok] == 0)


def test_integer_shift_is_recovered(ctx):
    """Frame B = frame A moved by (+3 cols, +1 row): well-textured interior corners are tracked to p + shift."""
    from ya_vo_b200 import synth
    a = synth.synth_frame("B4", 77, 376, 1241)
    b = np.empty_like(a)
    b[1:, 3:] = a[:-1, :-3]
    b[0, :] = b[1, :]
    b[:, :3] = b[:, 3:4]
    ctx.upload(4, a)
    ctx.upload(5, b)
    r, c, s, nc = ctx.fast_detect(4)
    keep = (r > 40) & (r < 330) & (c > 40) & (c < 1200)
    pts = np.stack([c[keep], r[keep]], 1).astype(np.float32)
    nxt, st, er = ctx.klt_track(4, 5, pts)
    ok = st == 1
    d = np.abs(nxt[ok] - (pts[ok] + np.float32([3, 1]))).max(1)
    assert ok.mean() > 0.95 and np.mean(d < 0.05) > 0.95


def test_argument_errors(ctx, cuda_lib, kitti):
    ctx.upload(0, kitti)
    ctx.upload(1, kitti[:100, :200].copy())
    p = np.float32([[10, 10]])
    with pytest.raises(cuda_lib.YavoError):
        ctx.klt_track(0, 1, p)                      # different frame sizes
    ctx.upload(1, kitti)
    with pytest.raises(cuda_lib.YavoError):
        ctx.klt_track(0, 1, p, win=(33, 11))        # window too large
    with pytest.raises(cuda_lib.YavoError):
        ctx.klt_track(0, 1, p, win=(2, 11))         # OpenCV asserts sides > 2
    with pytest.raises(cuda_lib.YavoError):
        ctx.klt_track(0, 1, p, flags=1)             # unknown flag
    with pytest.raises(cuda_lib.YavoError):
        ctx.klt_track(0, 9, p)                      # slot out of range
    nxt, st, er = ctx.klt_track(0, 1, np.zeros((0, 2), np.float32))
    assert nxt.shape == (0, 2)
    with pytest.raises(cuda_lib.YavoError):
        ctx.pyramid_level(0, 7)                     # level never built


def test_python_mirror_reads_like_the_cv2_call(gold, images):
    """ya_vo_b200.tracking.calcOpticalFlowPyrLK with the reference's arguments (src/LoopHandler.cc:372-375)."""
    from ya_vo_b200 import tracking as tr
    a, b = images["kitti"]
    key = "kitti_fast_c0"
    nextPts, status, err = tr.calcOpticalFlowPyrLK(a, b, gold[key + "_pts"], None, winSize=(11, 11), maxLevel=3,
                                                   criteria=(tr.TERM_CRITERIA_COUNT + tr.TERM_CRITERIA_EPS, 30, 0.01),
                                                   flags=0, minEigThreshold=0.001)
    assert nextPts.shape == (2000, 2) and status.shape == (2000, 1) and err.shape == (2000, 1)
    check_case(gold, key, nextPts, status.ravel(), err.ravel())
    p3 = gold[key + "_pts"][:10].reshape(10, 1, 2)
    n3, s3, e3 = tr.calcOpticalFlowPyrLK(a, b, p3, None, winSize=(11, 11), maxLevel=3, minEigThreshold=0.001)
    assert n3.shape == (10, 1, 2) and np.array_equal(n3.reshape(10, 2), nextPts[:10])


def test_cpp_drop_in_tracks_like_opencv(cuda_lib, gold, images, tmp_path):
    """yavo::calcOpticalFlowPyrLK (ya_vo_b200/host/include/Tracking.hpp) on the FAST keypoints of the KITTI frame:
    the same points and parameters as fixture case kitti_fast_c0, so the answer is cv2's."""
    import os
    import subprocess
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ya_vo_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s"])
    a, b = images["kitti"]
    pa, pb, out = (tmp_path / n for n in ("a.bin", "b.bin", "out.bin"))
    a.tofile(pa)
    b.tofile(pb)
    r = subprocess.run([os.path.join(host, "host_tests"), "track", str(pa), str(pb), str(a.shape[0]), str(a.shape[1]), str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(out, "rb").read()
    n = int(np.frombuffer(buf, "<i4", 1)[0])
    rec = np.frombuffer(buf, np.dtype([("px", "<f4"), ("py", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("err", "<f4"), ("st", "<i4")]), n, 4)
    key = "kitti_fast_c0"
    assert n == 2000 and np.array_equal(np.stack([rec["px"], rec["py"]], 1), gold[key + "_pts"])
    check_case(gold, key, np.stack([rec["nx"], rec["ny"]], 1).copy(), rec["st"].astype(np.uint8), rec["err"].copy())
