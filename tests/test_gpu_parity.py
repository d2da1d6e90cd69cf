"""GPU: the CUDA path, called through the C ABI, against the oracle — bit-exact."""
import hashlib
import os

import numpy as np
import pytest

from ya_vo_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx(cuda_lib, offsets):
    c = cuda_lib.Context(device=0, n_slots=8, max_rows=376, max_cols=1241, max_kp=2000)
    c.set_brief_offsets(offsets)
    yield c
    c.close()


FRAMES = [("U", 0, 376, 1241), ("G30", 1, 376, 1241), ("B4", 2, 376, 1241), ("U", 5, 37, 50), ("U", 6, 41, 133),
          ("U", 7, 9, 9), ("U", 8, 12, 300), ("G30", 9, 376, 1241), ("U", 10, 200, 128), ("U", 11, 33, 129)]


def test_upload_download_roundtrip(ctx):
    img = synth.synth_frame("U", 42, 100, 333)
    ctx.upload(0, img)
    assert np.array_equal(ctx.download(0), img)
    strided = synth.synth_frame("U", 43, 64, 200)[:, 10:150]
    ctx.upload(1, np.ascontiguousarray(strided))
    assert np.array_equal(ctx.download(1), strided)


@pytest.mark.parametrize("kind,seed,H,W", FRAMES)
def test_blur_bit_exact(ctx, oracle, kind, seed, H, W):
    img = synth.synth_frame(kind, seed, H, W)
    ctx.upload(0, img)
    assert np.array_equal(ctx.blurred(0), oracle.gaussian_blur(img))


def test_blur_against_cv2_golden(ctx, kitti):
    g = np.load(os.path.join(GOLDEN, "blur_golden.npz"))
    for n in sorted(k[:-3] for k in g.files if k.endswith("_in") and "_sha_" not in k):
        ctx.upload(0, g[n + "_in"])
        assert np.array_equal(ctx.blurred(0), g[n + "_out"]), n
    ctx.upload(0, kitti)
    assert sha(ctx.blurred(0)) == str(g["kitti_sha_out"])


@pytest.mark.parametrize("kind,seed,H,W", FRAMES)
def test_fast_candidates_bit_exact(ctx, oracle, kind, seed, H, W):
    """positions in scan order and float32 score bits (src/FastDetector.cc:298-335)."""
    img = synth.synth_frame(kind, seed, H, W)
    ctx.upload(0, img)
    r, c, s = ctx.fast_candidates(0)
    er, ec, es = oracle.fast_candidates(img)
    assert np.array_equal(r, er) and np.array_equal(c, ec)
    assert np.array_equal(s.view(np.uint32), es.view(np.uint32))


def test_fast_candidates_kitti(ctx, oracle, kitti):
    ctx.upload(0, kitti)
    r, c, s = ctx.fast_candidates(0)
    er, ec, es = oracle.fast_candidates(kitti)
    assert r.size == 3791
    assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(s.view(np.uint32), es.view(np.uint32))


@pytest.mark.parametrize("kind,seed,H,W", FRAMES)
def test_fast_detect_order_bit_exact(ctx, oracle, kind, seed, H, W):
    """keypoints in the reference's std::sort order, including tied responses (:343-362)."""
    img = synth.synth_frame(kind, seed, H, W)
    ctx.upload(0, img)
    r, c, s, nc = ctx.fast_detect(0)
    er, ec, es, enc = oracle.fast_detect(img, 2000)
    assert nc == enc and r.size == er.size
    assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(s.view(np.uint32), es.view(np.uint32))


def test_fast_detect_kitti_and_small_caps(ctx, oracle, kitti):
    ctx.upload(0, kitti)
    for K in (2000, 1, 17, 500):
        r, c, s, nc = ctx.fast_detect(0, K)
        er, ec, es, enc = oracle.fast_detect(kitti, K)
        assert nc == 3791 and np.array_equal(r, er) and np.array_equal(c, ec)
    flat = np.full((40, 60), 9, np.uint8)
    ctx.upload(1, flat)
    r, c, s, nc = ctx.fast_detect(1)
    assert nc == 0 and r.size == 0


def test_fast_detect_heavy_ties(ctx, oracle):
    """blocky frames: many candidates share a response, the order is decided by introsort's history."""
    tied = 0
    for seed in range(20, 26):
        img = synth.synth_frame("B4", seed)
        ctx.upload(0, img)
        r, c, s, nc = ctx.fast_detect(0)
        er, ec, es, _ = oracle.fast_detect(img, 2000)
        tied += int((np.diff(es) == 0).sum())
        assert np.array_equal(r, er) and np.array_equal(c, ec)
    assert tied > 0


@pytest.mark.parametrize("kind,seed,H,W", FRAMES[:6])
def test_brief_bit_exact(ctx, oracle, offsets, kind, seed, H, W):
    img = synth.synth_frame(kind, seed, H, W)
    ctx.upload(0, img)
    er, ec, es, _ = oracle.fast_detect(img, 2000)
    # FAST points plus arbitrary ones (computeBrief accepts any list: src/LoopHandler.cc:488-510)
    rng = np.random.default_rng(seed)
    rows = np.concatenate([er, rng.integers(0, H, 64), [8, H - 8, H - 8, 8]]).astype(np.int32)
    cols = np.concatenate([ec, rng.integers(0, W, 64), [8, W - 8, 8, W - 8]]).astype(np.int32)
    d, v, oob = ctx.brief_describe(0, rows, cols)
    ed, ev, eoob = oracle.brief(img, offsets, rows, cols)
    assert np.array_equal(v, ev) and np.array_equal(d, ed) and oob == eoob


@pytest.mark.parametrize("n1,n2", [(1, 1), (50, 70), (2000, 2000), (333, 4097), (4096, 129), (2000, 0)])
def test_match_bit_exact(ctx, oracle, n1, n2):
    d1 = synth.synth_descriptors(n1, n1 * 131 + n2)
    d2 = synth.planted_descriptors(d1, n2, 7) if n2 else synth.synth_descriptors(0, 1)
    if n2 > 45:
        d2[40] = d2[10]  # duplicated train descriptor: lowest index must win
    idx, dist, sec, rev = ctx.match(d1, d2, extensions=True)
    eidx, edist, esec, erev = oracle.match(d1, d2, extensions=True)
    assert np.array_equal(idx, eidx) and np.array_equal(dist, edist)
    assert np.array_equal(sec, esec)
    assert np.array_equal(rev, erev)


@pytest.mark.parametrize("matcher", ["tc", "tc8", "popc"])
@pytest.mark.parametrize("n1,n2", [(1, 1), (7, 3), (127, 255), (128, 224), (128, 256), (129, 225), (129, 257), (500, 449), (2000, 2000), (1950, 1949),
                                   (333, 4097), (4096, 129), (300, 0), (0, 40), (9000, 7000)])
def test_match_tensor_core_and_integer_kernels(ctx, oracle, matcher, n1, n2):
    """Brief::matchFeatures (src/BriefDescriptor.cc:163-183) through both device matchers: the tcgen05 kernel works on
    +-1.0 FP8 expansions of the bits (every partial sum is an integer <= 256, exact in FP32), the other one on
    XOR / POPC; distances, indices and the first-minimum rule must agree with the oracle bit for bit, including
    all-equal descriptors (every distance 0 -> index 0), exact duplicates and tile-boundary sizes."""
    d1 = synth.synth_descriptors(n1, n1 * 131 + n2)
    d2 = synth.planted_descriptors(d1, n2, 7) if (n1 and n2) else synth.synth_descriptors(n2, 1)
    if n2 > 300:
        d2[290] = d2[10]   # duplicated train descriptors in different tiles: the lowest index wins
        d2[40] = d2[10]
    if n1 > 5 and n2 > 5:
        d1[3] = d2[5]      # distance 0
        d1[4] = ~d2[5]     # distance 256 to one train descriptor
    ctx.set_matcher(matcher)
    try:
        idx, dist = ctx.match(d1, d2)
    finally:
        ctx.set_matcher("tc")
    eidx, edist = oracle.match(d1, d2)
    assert np.array_equal(dist, edist)
    assert np.array_equal(idx, eidx)


def test_match_tensor_core_extreme_distances(ctx, oracle):
    """all-zero vs all-one descriptors (distance 256 everywhere) and identical sets (distance 0, ties on every row)"""
    z = np.zeros((300, 32), np.uint8)
    o = np.full((520, 32), 255, np.uint8)
    for a, b in ((z, o), (o, z), (z, z[:257]), (o, o)):
        idx, dist = ctx.match(a, b)
        eidx, edist = oracle.match(a, b)
        assert np.array_equal(idx, eidx) and np.array_equal(dist, edist)


def test_frontend_batch_matches_oracle_pipeline(cuda_lib, oracle, offsets, kitti):
    """configs 2/3 in miniature: consecutive frames, FAST+BRIEF on each, match f-1 -> f."""
    a = synth.synth_frame("G30", 1000)
    frames = np.stack([a, synth.shifted_pair(a, 1), kitti, synth.synth_frame("B4", 1002), synth.synth_frame("U", 3)])
    n = frames.shape[0]
    with cuda_lib.Context(device=0, n_slots=n, max_rows=376, max_cols=1241, max_kp=2000) as c:
        c.set_brief_offsets(offsets)
        out = c.process_host_batch(frames, do_match=True)
    exp = oracle.pipeline(frames, offsets, 2000, True, nthreads=4)
    assert np.array_equal(out["n_kp"], exp["n_kp"])
    for f in range(n):
        k = exp["n_kp"][f]
        assert np.array_equal(out["rows"][f, :k], exp["rows"][f, :k])
        assert np.array_equal(out["cols"][f, :k], exp["cols"][f, :k])
        assert np.array_equal(out["scores"][f, :k].view(np.uint32), exp["scores"][f, :k].view(np.uint32))
        assert np.array_equal(out["desc"][f, :k], exp["desc"][f, :k])
        if f > 0:
            kq = exp["n_kp"][f - 1]
            assert np.array_equal(out["match_idx"][f, :kq], exp["match_idx"][f, :kq])
            assert np.array_equal(out["match_dist"][f, :kq], exp["match_dist"][f, :kq])
    # the shifted pair has true matches: most survive removeOutliers(…, 20)
    keep = cuda_lib.remove_outliers(out["match_dist"][1, :exp["n_kp"][0]], 20)
    assert keep.sum() > 100


def test_reference_style_pipeline(cuda_lib, oracle, offsets):
    """tests/BriefDescriptorTest.cc:9-64 call order through the class mirror."""
    from ya_vo_b200.frontend import Brief, FastDetector, Image
    a = synth.synth_frame("G30", 77)
    b = synth.shifted_pair(a, 78)
    brief = Brief(256, offsets=offsets)
    t1, t2 = Image(a), Image(b)
    fd = FastDetector(12, 50)
    f1 = fd.getFastFeatures(t1)
    f2 = fd.getFastFeatures(t2)
    brief.computeBrief(f1, t1)
    brief.computeBrief(f2, t2)
    matches = brief.matchFeatures(t1, t2)
    filt = []
    brief.removeOutliers(matches, filt, 20.0)
    er, ec, es, _ = oracle.fast_detect(a, 2000)
    assert f1 == list(zip(er.tolist(), ec.tolist()))
    ed, ev, _ = oracle.brief(a, offsets, er, ec)
    assert len(t1.keypoints) == int(ev.sum())
    assert [k.id for k in t1.keypoints] == np.nonzero(ev)[0].tolist()
    assert np.array_equal(np.array([k.featVec for k in t1.keypoints]), ed[ev])
    assert len(matches) == len(t1.keypoints)
    assert len(filt) > 100 and all(m.pt1.matched and m.pt2.matched for m in filt)


def test_known_answers_through_mirror(cuda_lib):
    """tests/FastDetectorTest.cc:6-80 and tests/ImageTest.cc:23-37 against the class mirror."""
    from ya_vo_b200.frontend import FastDetector, Image
    bres = np.load(os.path.join(GOLDEN, "bresenham_50x50.npy"))
    img = Image(np.zeros((50, 50), np.uint8))
    fd = FastDetector(12, 50)
    pts = fd.getBresenhamCirclePoints(img, 25, 25)
    assert len(pts) == 16
    for p in pts:
        fd.putPixel(img, p)
    diff = img.rawImage.astype(np.int16) - bres.astype(np.int16)
    assert diff.min() >= -255 and np.abs(np.clip(diff, 0, 255)).max() <= 1  # cv::subtract saturates at 0
    gold = Image(bres)
    assert all(gold.getPixelVal(p[0], p[1]) == 255 for p in pts)
    assert fd.checkContiguousPixels(img.getPixelVal(25, 25), pts, img) is True
    fd.putPixel(img, (25, 25))
    assert fd.checkContiguousPixels(img.getPixelVal(25, 25), pts, img) is False
    img2 = Image(np.zeros((50, 50), np.uint8))
    for p in pts[:11]:
        fd.putPixel(img2, p)
    assert fd.checkContiguousPixels(img2.getPixelVal(25, 25), pts, img2) is False


def test_filter_pairs_matches_remove_outliers(cuda_lib, oracle, offsets, kitti):
    """removeOutliers (src/BriefDescriptor.cc:213-231) + the callers' point-pair conversion, fused on the device."""
    a = synth.synth_frame("G30", 1000)
    frames = np.stack([a, synth.shifted_pair(a, 1), kitti, synth.shifted_pair(kitti, 2, 0, 2), np.full((376, 1241), 9, np.uint8),
                       synth.synth_frame("U", 3)])
    n = frames.shape[0]
    exp = oracle.pipeline(frames, offsets, 2000, True, nthreads=4)
    with cuda_lib.Context(device=0, n_slots=n, max_rows=376, max_cols=1241, max_kp=2000) as c:
        c.set_brief_offsets(offsets)
        c.upload_batch(0, frames)
        c.frontend_batch(0, n, True)
        for thr in (20, 64):
            n_pairs, min_dist, pairs = c.filter_pairs(0, n, thr)
            for f in range(1, n):
                kq, kt = exp["n_kp"][f - 1], exp["n_kp"][f]
                d = exp["match_dist"][f, :kq]
                if kq == 0 or kt == 0:
                    assert n_pairs[f] == 0
                    continue
                keep = oracle.remove_outliers(d, thr)
                assert n_pairs[f] == keep.sum() and min_dist[f] == d.min()
                qi = np.nonzero(keep)[0]
                ti = exp["match_idx"][f, :kq][keep]
                got = pairs[f, :n_pairs[f]]
                assert np.array_equal(got[:, 5], qi) and np.array_equal(got[:, 6], ti) and np.array_equal(got[:, 4], d[keep])
                assert np.array_equal(got[:, 0], exp["rows"][f - 1, qi]) and np.array_equal(got[:, 1], exp["cols"][f - 1, qi])
                assert np.array_equal(got[:, 2], exp["rows"][f, ti]) and np.array_equal(got[:, 3], exp["cols"][f, ti])


@pytest.mark.parametrize("H,W", [(1, 1), (2, 3), (3, 9), (4, 4), (8, 8), (9, 130), (130, 9), (17, 257)])
def test_tiny_and_odd_frames(cuda_lib, oracle, offsets, H, W):
    """frames smaller than the blur kernel / the FAST border: reflected borders, empty interiors."""
    img = synth.synth_frame("U", 900 + H * 7 + W, H, W)
    with cuda_lib.Context(device=0, n_slots=1, max_rows=H, max_cols=W, max_kp=64) as c:
        c.set_brief_offsets(offsets)
        c.upload(0, img)
        assert np.array_equal(c.download(0), img)
        assert np.array_equal(c.blurred(0), oracle.gaussian_blur(img))
        r, cc, s, nc = c.fast_detect(0)
        er, ec, es, enc = oracle.fast_detect(img, 64)
        assert nc == enc and np.array_equal(r, er) and np.array_equal(cc, ec)
        rows = np.array([0, H // 2, H - 1, 8], np.int32)
        cols = np.array([0, W // 2, W - 1, 8], np.int32)
        d, v, oob = c.brief_describe(0, rows, cols)
        ed, ev, eoob = oracle.brief(img, offsets, rows, cols)
        assert np.array_equal(v, ev) and np.array_equal(d, ed) and oob == eoob


def test_argument_errors_are_reported(cuda_lib, offsets):
    with cuda_lib.Context(device=0, n_slots=2, max_rows=64, max_cols=64, max_kp=100) as c:
        with pytest.raises(cuda_lib.YavoError):
            c.fast_detect(0)  # nothing uploaded
        with pytest.raises(cuda_lib.YavoError):
            c.upload(5, np.zeros((8, 8), np.uint8))  # slot out of range
        with pytest.raises(cuda_lib.YavoError):
            c.upload(0, np.zeros((65, 8), np.uint8))  # larger than the context
        c.upload(0, synth.synth_frame("U", 1, 64, 64))
        with pytest.raises(cuda_lib.YavoError):
            c.brief_describe(0, [10], [10])  # offsets not set
        with pytest.raises(cuda_lib.YavoError):
            c.set_brief_offsets(np.full((256, 4), 9, np.int32))  # outside [-8, 8]
        c.set_brief_offsets(offsets)
        with pytest.raises(cuda_lib.YavoError):
            c.fast_detect(0, 101)  # more keypoints than the context was created for
        c.upload(1, synth.synth_frame("U", 2, 32, 64))
        with pytest.raises(cuda_lib.YavoError):
            c.frontend_batch(0, 2, True)  # slots of different sizes in one batch
        d, v, _ = c.brief_describe(0, [], [])
        assert d.shape == (0, 32)
    with pytest.raises(cuda_lib.YavoError):
        cuda_lib.Context(device=99)


def test_randomised_frames_against_oracle(cuda_lib, oracle, offsets):
    """40 random shapes / contents (noise, blocky ties, gradients with planted corners): candidates, order, scores,
    descriptors — all through one context, so stale slot state would show."""
    rng = np.random.default_rng(2024)
    with cuda_lib.Context(device=0, n_slots=2, max_rows=300, max_cols=700, max_kp=300) as c:
        c.set_brief_offsets(offsets)
        for t in range(40):
            H, W = int(rng.integers(10, 301)), int(rng.integers(10, 701))
            kind = ("U", "G30", "B4")[t % 3]
            img = synth.synth_frame(kind, 5000 + t, H, W)
            if t % 5 == 0:  # smooth ramp with a few bright squares: few, strong corners
                yy, xx = np.mgrid[0:H, 0:W]
                img = ((yy + xx) % 200).astype(np.uint8)
                for _ in range(6):
                    r0, c0 = int(rng.integers(0, max(1, H - 12))), int(rng.integers(0, max(1, W - 12)))
                    img[r0:r0 + 9, c0:c0 + 9] = 255
            slot = t & 1
            c.upload(slot, img)
            r, cc, s, nc = c.fast_detect(slot)
            er, ec, es, enc = oracle.fast_detect(img, 300)
            assert nc == enc, (t, H, W, kind)
            assert np.array_equal(r, er) and np.array_equal(cc, ec), (t, H, W, kind)
            assert np.array_equal(s.view(np.uint32), es.view(np.uint32)), (t, H, W, kind)
            d, v, oob = c.brief_describe(slot, r, cc)
            ed, ev, eoob = oracle.brief(img, offsets, er, ec)
            assert np.array_equal(v, ev) and np.array_equal(d, ed) and oob == eoob, (t, H, W, kind)
