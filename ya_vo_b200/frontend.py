"""Python mirror of the reference's hot-path classes over the C ABI.

Same names, argument meaning and quirks as the reference's C++ classes so that the parity tests
read like the reference's own tests (tests/FastDetectorTest.cc, tests/BriefDescriptorTest.cc,
tests/ImageTest.cc):

  Image         include/Image.hpp:14-28          rawImage / keypoints / getW / getH / getPixelVal
  KeyPoint      include/BriefDescriptor.hpp:11-24
  Matches       include/BriefDescriptor.hpp:27-39
  FastDetector  include/FastDetector.hpp:17-55   ctor args ignored exactly as the reference does
  Brief         include/BriefDescriptor.hpp:41-68

Points are (x=row, y=col) tuples, as the reference's cv::Point(i, j) with i = row.
All pixel work runs in the CUDA library; the C++ drop-in classes in ya_vo_b200/host/ are the
equivalent for C++ callers.
"""
import random

import numpy as np

from . import capi

INT_MAX = 2**31 - 1

_default_ctx = {}


def default_context(rows, cols, device=0):
    """A shared 2-slot context large enough for (rows, cols) frames."""
    key = device
    ctx = _default_ctx.get(key)
    if ctx is None or ctx.max_rows < rows or ctx.max_cols < cols:
        if ctx is not None:
            ctx.close()
        ctx = capi.Context(device=device, n_slots=2, max_rows=max(rows, 376), max_cols=max(cols, 1241), max_kp=2000)
        _default_ctx[key] = ctx
    return ctx


class KeyPoint:
    __slots__ = ("x", "y", "id", "matched", "featVec")

    def __init__(self, x=0, y=0, id=0):
        self.x, self.y, self.id = int(x), int(y), int(id)
        self.matched = False
        self.featVec = np.zeros(32, np.uint8)


class Matches:
    __slots__ = ("pt1", "pt2", "distance")

    def __init__(self, pt1=None, pt2=None, distance=0):
        self.pt1, self.pt2, self.distance = pt1, pt2, int(distance)


class Image:
    """Owns a deep copy of an 8-bit single-channel frame (src/Image.cc:8-13)."""

    def __init__(self, img):
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2:
            raise ValueError("Image expects an 8-bit single-channel array (the reference asserts CV_8UC1)")
        self.rawImage = img.copy()
        self.keypoints = []
        self.resetKeypoints = []

    def getW(self):
        return self.rawImage.shape[1]

    def getH(self):
        return self.rawImage.shape[0]

    def getPixelVal(self, i, j):
        # data[i*cols + j], unchecked linear indexing (src/Image.cc:15-17)
        return int(self.rawImage.reshape(-1)[i * self.rawImage.shape[1] + j])


class FastDetector:
    def __init__(self, _minDetectionThresold=12, _intensityThreshold=50, ctx=None):
        # include/FastDetector.hpp:32-38: both arguments are ignored except for bookkeeping
        self.minDetectionThreshold = _minDetectionThresold
        self.bresRadius = 3
        self.intensityThreshold = 40
        self.fastCornerNumThreshold = 2000
        self.harrisThreshold = 2
        self._ctx = ctx

    def _context(self, img):
        return self._ctx or default_context(img.getH(), img.getW())

    def getBresenhamCirclePoints(self, img, xc, yc):
        return [tuple(int(v) for v in p) for p in capi.ring_points(xc, yc)]

    def checkInBetween(self, centPixel, condPixel):
        return (centPixel > condPixel - self.intensityThreshold) and (centPixel < condPixel + self.intensityThreshold)

    def checkContiguousPixels(self, centPixel, circlePoints, img):
        # src/FastDetector.cc:135-153 — a scalar predicate on 16 pixels; evaluated on the host
        pix = 0
        for k in range(16):
            if self.checkInBetween(int(centPixel), img.getPixelVal(circlePoints[k][0], circlePoints[k][1])):
                pix = 0
            else:
                pix += 1
            if pix >= 12:
                return True
        return False

    def putPixel(self, img, pt, pixVal=255):
        # rawImage.at<uint8_t>(pt): cv::Point(x, y) addresses row y, col x (src/FastDetector.cc:118-133)
        img.rawImage[pt[1], pt[0]] = pixVal

    def getFastFeatures(self, img, return_scores=False):
        ctx = self._context(img)
        # rawImage is a public mutable member (tests write to it), so the pixels are uploaded per call
        ctx.upload(0, img.rawImage)
        rows, cols, scores, _ = ctx.fast_detect(0, self.fastCornerNumThreshold)
        pts = list(zip(rows.tolist(), cols.tolist()))
        return (pts, scores) if return_scores else pts


class Brief:
    def __init__(self, numTests=256, offsets=None, ctx=None):
        self.patchSize = numTests
        self.offsets = self.preComputeOffsets() if offsets is None else np.asarray(offsets, np.int32).reshape(256, 4)
        self._ctx = ctx

    def preComputeOffsets(self):
        # src/BriefDescriptor.cc:4-20: 256 x 4 uniform integers in [-8, 8] from a non-deterministic seed
        rng = random.SystemRandom()
        return np.array([[rng.randint(-8, 8) for _ in range(4)] for _ in range(256)], np.int32)

    def popCount(self, v):
        return bin(int(v) & 0xFF).count("1")

    def hammingDistance(self, a, b):
        return int(np.unpackbits(np.bitwise_xor(np.asarray(a, np.uint8), np.asarray(b, np.uint8))).sum())

    def checkBoundry(self, x, y, width, height):
        return not (x - 8 < 0 or x + 8 > width or y - 8 < 0 or y + 8 > height)

    def _context(self, img):
        return self._ctx or default_context(img.getH(), img.getW())

    def computeBrief(self, detectedCornerPoints, img):
        """Appends a KeyPoint (id = index in the input list) for every admitted point."""
        ctx = self._context(img)
        ctx.upload(0, img.rawImage)
        ctx.set_brief_offsets(self.offsets)
        rows = np.array([p[0] for p in detectedCornerPoints], np.int32)
        cols = np.array([p[1] for p in detectedCornerPoints], np.int32)
        desc, valid, self.last_oob = ctx.brief_describe(0, rows, cols)
        for i in range(rows.size):
            if valid[i]:
                kp = KeyPoint(rows[i], cols[i], i)
                kp.featVec = desc[i].copy()
                img.keypoints.append(kp)

    def matchFeatures(self, img1, img2):
        ctx = self._context(img1)
        d1 = np.array([k.featVec for k in img1.keypoints], np.uint8).reshape(-1, 32)
        d2 = np.array([k.featVec for k in img2.keypoints], np.uint8).reshape(-1, 32)
        idx, dist = ctx.match(d1, d2)
        out = []
        for i, kp1 in enumerate(img1.keypoints):
            kp2 = KeyPoint(0, 0, 0)
            if idx[i] >= 0:
                t = img2.keypoints[idx[i]]
                kp2 = KeyPoint(t.x, t.y, t.id)
            out.append(Matches(kp1, kp2, int(dist[i])))
        return out

    def removeOutliers(self, matches, newMatches, threshold=30):
        if not matches:
            return  # the reference dereferences end() here
        keep = capi.remove_outliers(np.array([m.distance for m in matches], np.int32), int(threshold))
        for m, k in zip(matches, keep):
            if k:
                m.pt1.matched = True
                m.pt2.matched = True
                newMatches.append(m)
