"""ctypes binding of oracle/_ref/libyavo_ref.so — the reference's own FastDetector.cc / BriefDescriptor.cc /
Image.cc compiled unmodified over the stub OpenCV tree (oracle/ref_shim/README.md).

TEST INFRASTRUCTURE ONLY.  Used to pin the oracle (tests/) and to generate tests/golden/ref_golden.npz.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libyavo_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_harris.restype = C.c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ring(xc, yc):
    out = np.zeros(32, np.int32)
    n = lib().ref_ring(int(xc), int(yc), _p(out))
    return n, out.reshape(16, 2)


def check_contiguous(img, xc, yc):
    img = np.ascontiguousarray(img, np.uint8)
    return bool(lib().ref_check_contiguous(_p(img), img.shape[0], img.shape[1], int(xc), int(yc)))


def harris(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return np.float32(lib().ref_harris(_p(img), img.shape[0], img.shape[1], int(x), int(y)))


def fast(img, cap=4096):
    img = np.ascontiguousarray(img, np.uint8)
    rows = np.zeros(cap, np.int32)
    cols = np.zeros(cap, np.int32)
    n = lib().ref_fast(_p(img), img.shape[0], img.shape[1], _p(rows), _p(cols), cap)
    return rows[:min(n, cap)].copy(), cols[:min(n, cap)].copy()


def brief(img, offsets, rows, cols):
    img = np.ascontiguousarray(img, np.uint8)
    off = np.ascontiguousarray(offsets, np.int32).reshape(-1)
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    n = rows.size
    x = np.zeros(n, np.int32)
    y = np.zeros(n, np.int32)
    ids = np.zeros(n, np.int32)
    desc = np.zeros((n, 32), np.uint8)
    m = lib().ref_brief(_p(img), img.shape[0], img.shape[1], _p(off), _p(rows), _p(cols), n, _p(x), _p(y), _p(ids),
                        _p(desc))
    return x[:m].copy(), y[:m].copy(), ids[:m].copy(), desc[:m].copy()


def match(d1, d2, threshold=20):
    d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
    d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
    n1, n2 = d1.shape[0], d2.shape[0]
    idx = np.zeros(n1, np.int32)
    dist = np.zeros(n1, np.int32)
    keep = np.zeros(n1, np.uint8)
    lib().ref_match(_p(d1), n1, _p(d2), n2, int(threshold), _p(idx), _p(dist), _p(keep))
    return idx, dist, keep.astype(bool)


GOLDEN_FRAMES = [("U", 101, 64, 96), ("G30", 102, 72, 120), ("B4", 103, 96, 128), ("B4", 104, 80, 200), ("U", 105, 40, 48),
                 ("U", 106, 160, 420), ("B4", 107, 200, 360)]  # the last two exceed 2000 candidates: exercise the cut


def write_golden(path):
    """Runs the shimmed reference on small frames and stores inputs' recipe + outputs."""
    import sys
    sys.path.insert(0, os.path.dirname(_HERE))
    from ya_vo_b200 import synth
    off = synth.brief_offsets()
    out = {"offsets": off}
    import cv2
    kitti = cv2.imread(os.path.join(os.path.dirname(_HERE), "tests", "golden", "kitti_frame.png"), 0)
    frames = [(k, s, h, w, synth.synth_frame(k, s, h, w)) for k, s, h, w in GOLDEN_FRAMES]
    frames.append(("K", 0, 120, 160, np.ascontiguousarray(kitti[150:270, 500:660])))
    names = []
    for kind, seed, h, w, img in frames:
        name = "%s_%d_%dx%d" % (kind, seed, h, w)
        names.append(name)
        r, c = fast(img)
        x, y, ids, desc = brief(img, off, r, c)
        out[name + "_img"] = img
        out[name + "_fast_rows"], out[name + "_fast_cols"] = r, c
        out[name + "_kp_x"], out[name + "_kp_y"], out[name + "_kp_id"], out[name + "_desc"] = x, y, ids, desc
        hs = np.array([harris(img, a, b) for a, b in zip(r[:64], c[:64])], np.float32)
        out[name + "_harris64"] = hs
    # matching: descriptors of consecutive golden frames + a planted set with duplicate minima
    d1, d2 = out[names[2] + "_desc"], out[names[3] + "_desc"]
    idx, dist, keep = match(d1, d2, 20)
    out["match_a_idx"], out["match_a_dist"], out["match_a_keep"] = idx, dist, keep
    p1 = synth.synth_descriptors(300, 5)
    p2 = synth.planted_descriptors(p1, 400, 6)
    p2[40] = p2[10]
    idx, dist, keep = match(p1, p2, 20)
    out["match_b_d1"], out["match_b_d2"] = p1, p2
    out["match_b_idx"], out["match_b_dist"], out["match_b_keep"] = idx, dist, keep
    out["names"] = np.array(names)
    np.savez_compressed(path, **out)
    print("ref_golden: %d frames, %s" % (len(names), ", ".join("%s:%d kp" % (n, out[n + "_fast_rows"].size) for n in names)))


class Stream:
    """Steady-state timing form: push(frame) = getFastFeatures + computeBrief on the frame and
    matchFeatures(previous, frame) + removeOutliers, all inside the reference's own object code."""

    def __init__(self, offsets):
        L = lib()
        L.ref_stream_new.restype = C.c_void_p
        L.ref_stream_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_stream_free.argtypes = [C.c_void_p]
        off = np.ascontiguousarray(offsets, np.int32).reshape(-1)
        self.h = L.ref_stream_new(_p(off))

    def push(self, frame, threshold=20):
        frame = np.ascontiguousarray(frame, np.uint8)
        nk = np.zeros(1, np.int32)
        kept = np.zeros(1, np.int32)
        lib().ref_stream_push(self.h, _p(frame), frame.shape[0], frame.shape[1], int(threshold), _p(nk), _p(kept))
        return int(nk[0]), int(kept[0])

    def close(self):
        if self.h:
            lib().ref_stream_free(self.h)
            self.h = None
