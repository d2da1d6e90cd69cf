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
