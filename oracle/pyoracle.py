"""ctypes binding of the CPU oracle (oracle/libyavo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under ya_vo_b200/
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libyavo_oracle.so")
INT_MAX = 2**31 - 1


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("yavo_oracle.cpp", "yavo_oracle_klt.cpp", "yavo_oracle_geom.cpp", "yavo_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libyavo_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.yavo_oracle_score_from_tensor.restype = C.c_float
        _lib.yavo_oracle_score_from_tensor.argtypes = [C.c_float] * 3
        _lib.yavo_oracle_harris.restype = C.c_float
        _lib.yavo_oracle_eigen2x2.argtypes = [C.c_float] * 3 + [C.c_void_p] * 2
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def ring(xc, yc):
    out = np.zeros(32, np.int32)
    lib().yavo_oracle_ring(int(xc), int(yc), _p(out))
    return out.reshape(16, 2)


def ring_literal(xc, yc):
    out = np.zeros(32, np.int32)
    n = lib().yavo_oracle_ring_literal(int(xc), int(yc), _p(out))
    return n, out.reshape(16, 2)


def check_in_between(c, p):
    return bool(lib().yavo_oracle_check_in_between(C.c_uint8(c), C.c_uint8(p)))


def check_contiguous(c, ring_vals):
    rv = _u8(ring_vals)
    assert rv.size == 16
    return bool(lib().yavo_oracle_check_contiguous(C.c_uint8(c), _p(rv)))


def sobel(img):
    img = _u8(img)
    H, W = img.shape
    ix = np.zeros((H, W), np.float32)
    iy = np.zeros((H, W), np.float32)
    lib().yavo_oracle_sobel(_p(img), H, W, _p(ix), _p(iy))
    return ix, iy


def eigen2x2(a, b, c):
    l1 = C.c_float()
    l2 = C.c_float()
    lib().yavo_oracle_eigen2x2(C.c_float(a), C.c_float(b), C.c_float(c), C.addressof(l1), C.addressof(l2))
    return np.float32(l1.value), np.float32(l2.value)


def score_from_tensor(a, b, c):
    return np.float32(lib().yavo_oracle_score_from_tensor(float(a), float(b), float(c)))


def harris(img, x, y):
    img = _u8(img)
    H, W = img.shape
    return np.float32(lib().yavo_oracle_harris(_p(img), H, W, int(x), int(y)))


def fast_candidates(img):
    img = _u8(img)
    H, W = img.shape
    cap = H * W
    rows = np.zeros(cap, np.int32)
    cols = np.zeros(cap, np.int32)
    sc = np.zeros(cap, np.float32)
    n = lib().yavo_oracle_fast_candidates(_p(img), H, W, cap, _p(rows), _p(cols), _p(sc))
    return rows[:n].copy(), cols[:n].copy(), sc[:n].copy()


def fast_detect(img, max_kp=2000):
    img = _u8(img)
    H, W = img.shape
    rows = np.zeros(max_kp, np.int32)
    cols = np.zeros(max_kp, np.int32)
    sc = np.zeros(max_kp, np.float32)
    nc = C.c_int()
    n = lib().yavo_oracle_fast_detect(_p(img), H, W, int(max_kp), _p(rows), _p(cols), _p(sc), C.byref(nc))
    return rows[:n].copy(), cols[:n].copy(), sc[:n].copy(), nc.value


def std_sort_desc(scores, payload):
    s = np.ascontiguousarray(scores, np.float32).copy()
    p = np.ascontiguousarray(payload, np.int32).copy()
    lib().yavo_oracle_std_sort_desc(_p(s), _p(p), s.size)
    return s, p


def introsort_topk(scores, payload, k, depth=-1):
    s = np.ascontiguousarray(scores, np.float32).copy()
    p = np.ascontiguousarray(payload, np.int32).copy()
    lib().yavo_oracle_introsort_topk_depth(_p(s), _p(p), s.size, int(k), int(depth))
    return s, p


def gaussian_blur(img):
    img = _u8(img)
    H, W = img.shape
    out = np.zeros((H, W), np.uint8)
    lib().yavo_oracle_gaussian_blur(_p(img), H, W, _p(out))
    return out


def brief(img, offsets, rows, cols, blurred=None):
    img = _u8(img)
    H, W = img.shape
    off = np.ascontiguousarray(offsets, np.int32).reshape(-1)
    assert off.size == 1024
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    n = rows.size
    desc = np.zeros((n, 32), np.uint8)
    valid = np.zeros(n, np.uint8)
    oob = C.c_int()
    bl = None if blurred is None else _u8(blurred)
    lib().yavo_oracle_brief(_p(img), _p(bl), H, W, _p(off), _p(rows), _p(cols), n, _p(desc), _p(valid),
                            C.byref(oob))
    return desc, valid.astype(bool), oob.value


def popcount(v):
    return lib().yavo_oracle_popcount(C.c_uint8(v))


def hamming(a, b):
    a = _u8(a)
    b = _u8(b)
    return lib().yavo_oracle_hamming(_p(a), _p(b))


def match(d1, d2, extensions=False):
    d1 = _u8(d1).reshape(-1, 32)
    d2 = _u8(d2).reshape(-1, 32)
    n1, n2 = d1.shape[0], d2.shape[0]
    idx = np.zeros(n1, np.int32)
    dist = np.zeros(n1, np.int32)
    if not extensions:
        lib().yavo_oracle_match(_p(d1), n1, _p(d2), n2, _p(idx), _p(dist), None, None)
        return idx, dist
    sec = np.zeros(n1, np.int32)
    rev = np.zeros(n2, np.int32)
    lib().yavo_oracle_match(_p(d1), n1, _p(d2), n2, _p(idx), _p(dist), _p(sec), _p(rev))
    return idx, dist, sec, rev


def remove_outliers(dist, threshold=20):
    dist = np.ascontiguousarray(dist, np.int32)
    keep = np.zeros(dist.size, np.uint8)
    lib().yavo_oracle_remove_outliers(_p(dist), dist.size, int(threshold), _p(keep))
    return keep.astype(bool)


def pipeline(frames, offsets, max_kp=2000, do_match=True, nthreads=1, outputs=True):
    frames = _u8(frames)
    F, H, W = frames.shape
    off = np.ascontiguousarray(offsets, np.int32).reshape(-1)
    if not outputs:
        lib().yavo_oracle_pipeline(_p(frames), F, H, W, _p(off), max_kp, int(do_match), int(nthreads),
                                   None, None, None, None, None, None, None)
        return None
    r = np.zeros((F, max_kp), np.int32)
    c = np.zeros((F, max_kp), np.int32)
    s = np.zeros((F, max_kp), np.float32)
    d = np.zeros((F, max_kp, 32), np.uint8)
    n = np.zeros(F, np.int32)
    mi = np.full((F, max_kp), -1, np.int32)
    md = np.full((F, max_kp), -1, np.int32)
    lib().yavo_oracle_pipeline(_p(frames), F, H, W, _p(off), max_kp, int(do_match), int(nthreads),
                               _p(r), _p(c), _p(s), _p(d), _p(n), _p(mi), _p(md))
    return dict(rows=r, cols=c, scores=s, desc=d, n_kp=n, match_idx=mi, match_dist=md)


# ---- sparse pyramidal Lucas-Kanade ----------------------------------------------------------
def pyr_down(img):
    img = _u8(img)
    H, W = img.shape
    out = np.zeros(((H + 1) // 2, (W + 1) // 2), np.uint8)
    lib().yavo_oracle_pyr_down(_p(img), H, W, _p(out))
    return out


def scharr(img):
    img = _u8(img)
    H, W = img.shape
    dx = np.zeros((H, W), np.int16)
    dy = np.zeros((H, W), np.int16)
    lib().yavo_oracle_scharr(_p(img), H, W, _p(dx), _p(dy))
    return dx, dy


def klt_levels(H, W, win=(11, 11), max_level=3):
    return lib().yavo_oracle_klt_levels(int(H), int(W), int(win[0]), int(win[1]), int(max_level))


def klt_track(prev, nxt, prev_pts, win=(11, 11), max_level=3, crit_type=3, max_count=30, epsilon=0.01, flags=0,
              min_eig=1e-3, init_pts=None):
    """cv::calcOpticalFlowPyrLK restated (src/LoopHandler.cc:372-375).  Points are (x = col, y = row) float32.
    Returns next_pts (n,2) f32, status (n,) u8, err (n,) f32."""
    prev, nxt = _u8(prev), _u8(nxt)
    assert prev.shape == nxt.shape
    H, W = prev.shape
    pp = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = pp.shape[0]
    nx = np.zeros((n, 2), np.float32)
    if init_pts is not None:
        nx[:] = np.asarray(init_pts, np.float32).reshape(-1, 2)
    st = np.zeros(n, np.uint8)
    er = np.zeros(n, np.float32)
    f = lib().yavo_oracle_klt
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
    f(_p(prev), _p(nxt), H, W, _p(pp), n, _p(nx), _p(st), _p(er), int(win[0]), int(win[1]), int(max_level),
      int(crit_type), int(max_count), float(epsilon), int(flags), float(min_eig))
    return nx, st, er


# ---- inlier count of the reference's F-matrix RANSAC -------------------------------------------
def epipolar_inliers(F, x1, y1, x2, y2, threshold=0.1, residuals=False):
    """src/3DHandler.cc:163-188.  F: (m, 3, 3) float64; x1, y1, x2, y2: int32 (n,), the Matches' pt1.x, pt1.y, pt2.x,
    pt2.y as the reference reads them.  Returns counts (m,), best index, best count[, residuals (m, n)]."""
    F = np.ascontiguousarray(F, np.float64).reshape(-1, 9)
    a = [np.ascontiguousarray(v, np.int32) for v in (x1, y1, x2, y2)]
    m, n = F.shape[0], a[0].size
    counts = np.zeros(m, np.int32)
    res = np.zeros((m, n), np.float64) if residuals else None
    best, bc = C.c_int32(), C.c_int32()
    f = lib().yavo_oracle_epipolar_inliers
    f.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(F), m, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), n, float(threshold), _p(counts), _p(res), C.byref(best), C.byref(bc))
    return (counts, best.value, bc.value, res) if residuals else (counts, best.value, bc.value)
