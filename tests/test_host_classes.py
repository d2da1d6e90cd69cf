"""The C++ drop-in classes (ya_vo_b200/host: Image / FastDetector / Brief with the reference's signatures):
the reference's own known-answer tests on the CPU, the BriefDescriptorTest pipeline on the GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

from ya_vo_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ya_vo_b200", "host")
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def host_tests(cuda_lib):
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return os.path.join(HOST, "host_tests")


def test_reference_known_answers(host_tests, tmp_path):
    """tests/FastDetectorTest.cc:6-80, tests/ImageTest.cc:23-37 — host-side helpers only, no device."""
    p = tmp_path / "bres.bin"
    np.load(os.path.join(GOLDEN, "bresenham_50x50.npy")).tofile(p)
    r = subprocess.run([host_tests, "known", str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_headers_keep_reference_signatures():
    fd = open(os.path.join(HOST, "include", "FastDetector.hpp")).read()
    br = open(os.path.join(HOST, "include", "BriefDescriptor.hpp")).read()
    im = open(os.path.join(HOST, "include", "Image.hpp")).read()
    for sig in ("std::vector<cv::Point> getFastFeatures(const Image &img)",
                "std::vector<cv::Point> getBresenhamCirclePoints(const Image &img, int x, int y)",
                "bool checkContiguousPixels(uint8_t centPixel, const std::vector<cv::Point> &circlePoints, const Image &img)",
                "FastDetector(int _minDetectionThresold, uint8_t"):
        assert sig in fd, sig
    for sig in ("void computeBrief(const std::vector<cv::Point> &detectedCornerPoints, Image &img)",
                "std::vector<Matches> matchFeatures(Image &img1, Image &img2)",
                "void removeOutliers(std::vector<Matches> &matches, std::vector<Matches> &newMatches, int threshold)",
                "int hammingDistance(uchar featVec1[32], uchar featVec2[32])", "uchar featVec[32]"):
        assert sig in br, sig
    tk = open(os.path.join(HOST, "include", "Tracking.hpp")).read()
    for sig in ("void calcOpticalFlowPyrLK(const cv::Mat &prevImg, const cv::Mat &nextImg, const std::vector<cv::Point2f> &prevPts,",
                "cv::Size winSize = cv::Size(21, 21), int maxLevel = 3", "int flags = 0, double minEigThreshold = 1e-4"):
        assert sig in tk, sig
    for sig in ("Image(const cv::Mat &img)", "cv::Mat rawImage", "std::vector<KeyPoint> keypoints",
                "uint8_t getPixelVal(int i, int j) const"):
        assert sig in im, sig


def _read(buf, off, fmt):
    v = struct.unpack_from(fmt, buf, off)
    return v, off + struct.calcsize(fmt)


@pytest.mark.gpu
def test_brief_descriptor_test_pipeline(host_tests, oracle, offsets, tmp_path):
    """tests/BriefDescriptorTest.cc:9-64 through the C++ classes, compared with the oracle."""
    a = synth.synth_frame("G30", 77)
    b = synth.shifted_pair(a, 78)
    H, W = a.shape
    pa, pb, po_, out = (tmp_path / n for n in ("a.bin", "b.bin", "off.bin", "out.bin"))
    a.tofile(pa)
    b.tofile(pb)
    offsets.astype(np.int32).tofile(po_)
    r = subprocess.run([host_tests, "pipeline", str(pa), str(pb), str(H), str(W), str(po_), str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(out, "rb").read()
    off = 0
    (n1,), off = _read(buf, off, "<i")
    f1 = np.frombuffer(buf, dtype=np.dtype([("x", "<i4"), ("y", "<i4"), ("s", "<f4")]), count=n1, offset=off)
    off += 12 * n1
    (n2,), off = _read(buf, off, "<i")
    f2 = np.frombuffer(buf, dtype=np.dtype([("x", "<i4"), ("y", "<i4")]), count=n2, offset=off)
    off += 8 * n2
    kps = []
    kdt = np.dtype([("x", "<i4"), ("y", "<i4"), ("id", "<i4"), ("d", "u1", (32,))])
    for _ in range(2):
        (nk,), off = _read(buf, off, "<i")
        kps.append(np.frombuffer(buf, dtype=kdt, count=nk, offset=off))
        off += kdt.itemsize * nk
    (nm,), off = _read(buf, off, "<i")
    m = np.frombuffer(buf, dtype=np.dtype([("id1", "<i4"), ("id2", "<i4"), ("x2", "<i4"), ("y2", "<i4"), ("dist", "<i4"),
                                           ("matched", "<i4")]), count=nm, offset=off)
    off += 24 * nm
    (nf,), off = _read(buf, off, "<i")
    fm = np.frombuffer(buf, dtype=np.dtype([("id1", "<i4"), ("dist", "<i4")]), count=nf, offset=off)

    er, ec, es, _ = oracle.fast_detect(a, 2000)
    assert np.array_equal(f1["x"], er) and np.array_equal(f1["y"], ec)
    assert np.array_equal(f1["s"].view(np.uint32), es.view(np.uint32))
    er2, ec2, _, _ = oracle.fast_detect(b, 2000)
    assert np.array_equal(f2["x"], er2) and np.array_equal(f2["y"], ec2)
    d1, v1, _ = oracle.brief(a, offsets, er, ec)
    d2, v2, _ = oracle.brief(b, offsets, er2, ec2)
    assert np.array_equal(kps[0]["id"], np.nonzero(v1)[0]) and np.array_equal(kps[0]["d"], d1[v1])
    assert np.array_equal(kps[1]["id"], np.nonzero(v2)[0]) and np.array_equal(kps[1]["d"], d2[v2])
    idx, dist = oracle.match(d1[v1], d2[v2])
    assert nm == int(v1.sum())
    assert np.array_equal(m["dist"], dist)
    assert np.array_equal(m["id2"], np.nonzero(v2)[0][idx])
    assert np.array_equal(m["x2"], er2[v2][idx]) and np.array_equal(m["y2"], ec2[v2][idx])
    keep = oracle.remove_outliers(dist, 20)
    assert nf == int(keep.sum()) and np.array_equal(m["matched"].astype(bool), keep)
    assert np.array_equal(fm["dist"], dist[keep])


def test_reference_callers_compile_verbatim():
    """Source compatibility: tests/BriefDescriptorTest.cc:10-47 and the body of LoopHandler::insertFrameFeatures
    (src/LoopHandler.cc:469-484) are cut out of the reference checkout at build time and compiled unchanged against
    ya_vo_b200/host/include (ya_vo_b200/host/test/ref_callers.cc).  Where the checkout is absent (the GPU box) the binary
    built beside it must be there."""
    subprocess.check_call(["make", "-C", HOST, "-s"])
    assert os.access(os.path.join(HOST, "ref_callers"), os.X_OK)
    if os.path.isdir("/root/reference/src"):
        body = open(os.path.join(HOST, "test", "_gen", "insert_frame_features.inc")).read()
        assert "fd.getFastFeatures(*_frame)" in body and "brief.computeBrief(features, *_frame)" in body
        body = open(os.path.join(HOST, "test", "_gen", "brief_test_body.inc")).read()
        assert "brief.matchFeatures(testObj1, testObj2)" in body and "brief.removeOutliers(matches, filterMatches, 20.0)" in body
    # nothing of the reference is committed
    assert "_gen" in open(os.path.join(ROOT, ".gitignore")).read()


def _keypoints(buf, off):
    kdt = np.dtype([("x", "<i4"), ("y", "<i4"), ("id", "<i4"), ("d", "u1", (32,))])
    (nk,), off = _read(buf, off, "<i")
    k = np.frombuffer(buf, dtype=kdt, count=nk, offset=off)
    return k, off + kdt.itemsize * nk


@pytest.mark.gpu
def test_reference_callers_run_on_the_gpu(host_tests, oracle, tmp_path):
    """The verbatim reference callers on the GPU: the Brief objects draw their offset tables from std::random_device as the
    reference does, the harness dumps them, and the oracle replays the run with the same tables."""
    a = synth.synth_frame("G30", 91)
    b = synth.shifted_pair(a, 92)
    H, W = a.shape
    pa, pb, out = (tmp_path / n for n in ("a.bin", "b.bin", "out.bin"))
    a.tofile(pa)
    b.tofile(pb)
    r = subprocess.run([os.path.join(HOST, "ref_callers"), str(pa), str(pb), str(H), str(W), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Time taken for FAST feature detection" in r.stdout and "Descriptor generation cost time" in r.stdout
    buf = open(out, "rb").read()
    off = 0
    table = np.frombuffer(buf, dtype="<i4", count=1024, offset=off).reshape(256, 4)
    off += 4096
    feats = []
    for _ in range(2):
        (n,), off = _read(buf, off, "<i")
        feats.append(np.frombuffer(buf, dtype=np.dtype([("x", "<i4"), ("y", "<i4")]), count=n, offset=off))
        off += 8 * n
    k1, off = _keypoints(buf, off)
    k2, off = _keypoints(buf, off)
    (nm,), off = _read(buf, off, "<i")
    m = np.frombuffer(buf, dtype=np.dtype([("id1", "<i4"), ("id2", "<i4"), ("dist", "<i4")]), count=nm, offset=off)
    off += 12 * nm
    (nf,), off = _read(buf, off, "<i")
    fm = np.frombuffer(buf, dtype=np.dtype([("id1", "<i4"), ("dist", "<i4")]), count=nf, offset=off)
    off += 8 * nf
    (cols,), off = _read(buf, off, "<i")
    assert cols == 2 * W
    exp = []
    for img, f, k in ((a, feats[0], k1), (b, feats[1], k2)):
        er, ec, _, _ = oracle.fast_detect(img, 2000)
        assert np.array_equal(f["x"], er) and np.array_equal(f["y"], ec)
        d, v, _ = oracle.brief(img, table, er, ec)
        assert np.array_equal(k["id"], np.nonzero(v)[0]) and np.array_equal(k["d"], d[v])
        exp.append((d[v], np.nonzero(v)[0]))
    idx, dist = oracle.match(exp[0][0], exp[1][0])
    assert nm == len(dist) and np.array_equal(m["dist"], dist) and np.array_equal(m["id2"], exp[1][1][idx])
    keep = oracle.remove_outliers(dist, 20)
    assert nf == int(keep.sum()) and np.array_equal(fm["dist"], dist[keep])
    # LoopHandler::insertFrameFeatures on both frames, its own random table
    table2 = np.frombuffer(buf, dtype="<i4", count=1024, offset=off).reshape(256, 4)
    off += 4096
    for img in (a, b):
        k, off = _keypoints(buf, off)
        er, ec, _, _ = oracle.fast_detect(img, 2000)
        d, v, _ = oracle.brief(img, table2, er, ec)
        assert np.array_equal(k["id"], np.nonzero(v)[0]) and np.array_equal(k["d"], d[v])
        assert np.array_equal(k["x"], er[v]) and np.array_equal(k["y"], ec[v])
    assert off == len(buf)


@pytest.mark.gpu
def test_cpp_frame_stream_matches_oracle(host_tests, oracle, offsets, tmp_path):
    """yavo::FrameStream (ya_vo_b200/host/include/FrameStream.hpp): the C++ streaming caller over
    yavo_submit_host_batch / yavo_wait_batch — decoder thread, batches with seam frames, per-frame callbacks in order."""
    F = 11
    frames = synth.synth_batch(F, "G30", 500)
    frames[6] = synth.shifted_pair(frames[5], 3)
    H, W = frames.shape[1:]
    pf, po_, out = (tmp_path / n for n in ("f.bin", "off.bin", "out.bin"))
    frames.tofile(pf)
    offsets.astype(np.int32).tofile(po_)
    r = subprocess.run([host_tests, "stream", str(pf), str(F), str(H), str(W), str(po_), "4", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exp = oracle.pipeline(frames, offsets, 2000, True, nthreads=8)
    buf = open(out, "rb").read()
    off = 0
    for f in range(F):
        (fr, k), off = _read(buf, off, "<ii")
        assert fr == f and k == exp["n_kp"][f]
        rows = np.frombuffer(buf, "<i4", k, off); off += 4 * k
        cols = np.frombuffer(buf, "<i4", k, off); off += 4 * k
        desc = np.frombuffer(buf, "u1", 32 * k, off).reshape(k, 32); off += 32 * k
        assert np.array_equal(rows, exp["rows"][f, :k]) and np.array_equal(cols, exp["cols"][f, :k])
        assert np.array_equal(desc, exp["desc"][f, :k])
        (kq,), off = _read(buf, off, "<i")
        assert kq == (exp["n_kp"][f - 1] if f > 0 else 0)
        if kq:
            mi = np.frombuffer(buf, "<i4", kq, off); off += 4 * kq
            md = np.frombuffer(buf, "<i4", kq, off); off += 4 * kq
            assert np.array_equal(mi, exp["match_idx"][f, :kq]) and np.array_equal(md, exp["match_dist"][f, :kq])
    assert off == len(buf)


@pytest.mark.gpu
def test_single_frame_latency_mode_and_trusted_identity(host_tests, offsets, tmp_path):
    """host_tests latency: the reference's per-frame call shape through the drop-in classes (what bench.py reports as
    `latency`), with the byte-for-byte residency check (default) and with YAVO_TRUST_IMAGE_IDENTITY=1."""
    import json
    frames = synth.synth_batch(6, "G30", 900)
    H, W = frames.shape[1:]
    pf, po_ = tmp_path / "f.bin", tmp_path / "off.bin"
    frames.tofile(pf)
    offsets.astype(np.int32).tofile(po_)
    for env in ({}, {"YAVO_TRUST_IMAGE_IDENTITY": "1"}):
        r = subprocess.run([host_tests, "latency", str(pf), "40", str(H), str(W), str(po_)], capture_output=True, text=True,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout + r.stderr
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert d["frames"] == 40 and 1800 < d["mean_keypoints"] <= 2000 and d["getFastFeatures"]["calls"] == 40
        assert d["computeBrief"]["p50"] < d["getFastFeatures"]["p50"]  # the descriptors came with the detection
