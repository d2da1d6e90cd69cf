"""ctypes binding of the C ABI in include/yavo_b200.h (ya_vo_b200/csrc/libyavo_b200.so).

This is the product boundary: plain pointers and sizes, no torch types.  The library is built
in-tree by build() (nvcc, sm_100a only) and there is no CPU path — on a machine without a CUDA
device `Context(...)` raises YavoError.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
# YAVO_LIB_PATH: tuning experiments only (A/B runs of differently built libraries in one GPU session)
LIB_PATH = os.environ.get("YAVO_LIB_PATH") or os.path.join(_CSRC, "libyavo_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "yavo_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-shared", "-Xcompiler", "-fPIC"]

EXPORTS = [
    "yavo_create", "yavo_destroy", "yavo_last_error", "yavo_kernel_launches", "yavo_sync",
    "yavo_get_stream", "yavo_set_profiling", "yavo_profile_collect",
    "yavo_upload", "yavo_upload_batch", "yavo_upload_from_device", "yavo_download",
    "yavo_ring_points", "yavo_fast_detect", "yavo_fast_candidates",
    "yavo_set_brief_offsets", "yavo_blurred", "yavo_brief_describe",
    "yavo_match", "yavo_remove_outliers",
    "yavo_frontend_batch", "yavo_fetch_batch", "yavo_process_host_batch", "yavo_submit_host_batch", "yavo_wait", "yavo_wait_batch", "yavo_filter_pairs", "yavo_set_pipeline_chunk", "yavo_set_sub_batch",
    "yavo_build_pyramid", "yavo_pyramid_level", "yavo_klt_track", "yavo_klt_track_batch", "yavo_klt_fetch", "yavo_epipolar_inliers", "yavo_stream_tracking", "yavo_stream_track_outputs", "yavo_pinned_alloc", "yavo_pinned_free", "yavo_set_matcher", "yavo_set_overlap", "yavo_frame_features", "yavo_slot_holds", "yavo_set_big_select",
]


class YavoError(RuntimeError):
    pass


def _sources():
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + [HEADER]


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a (cross-compiles without a GPU)."""
    if os.environ.get("YAVO_LIB_PATH"):
        return LIB_PATH  # an explicitly chosen build is used as it is
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in _sources())
    if not (force or stale):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("YAVO_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DYAVO_SEL_SMEM_ENTS=4096
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(_CSRC, "yavo_capi.cu")]
    subprocess.check_call(cmd, cwd=_CSRC)
    return LIB_PATH


_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YavoError("libyavo_b200.so is missing: run ya_vo_b200.capi.build() (nvcc, sm_100a). "
                            "There is no CPU implementation to fall back to.")
        L = C.CDLL(LIB_PATH)
        L.yavo_last_error.restype = C.c_char_p
        L.yavo_last_error.argtypes = [C.c_void_p]
        L.yavo_kernel_launches.restype = C.c_longlong
        L.yavo_kernel_launches.argtypes = [C.c_void_p]
        L.yavo_destroy.restype = None
        L.yavo_destroy.argtypes = [C.c_void_p]
        L.yavo_ring_points.restype = None
        L.yavo_pinned_alloc.restype = C.c_void_p
        L.yavo_pinned_alloc.argtypes = [C.c_size_t]
        L.yavo_pinned_free.restype = None
        L.yavo_pinned_free.argtypes = [C.c_void_p]
        L.yavo_get_stream.restype = C.c_void_p
        L.yavo_get_stream.argtypes = [C.c_void_p]
        L.yavo_upload_from_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One device, one stream, `n_slots` device-resident frames (include/yavo_b200.h: yavo_create)."""

    def __init__(self, device=0, n_slots=2, max_rows=376, max_cols=1241, max_kp=2000, max_cand=0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.yavo_create(int(device), int(n_slots), int(max_rows), int(max_cols), int(max_kp), int(max_cand),
                                 C.byref(h))
        if rc != 0:
            raise YavoError("yavo_create failed (%d): %s" % (rc, self._L.yavo_last_error(None).decode()))
        self._h = h
        self.device, self.n_slots, self.max_rows, self.max_cols, self.max_kp = device, n_slots, max_rows, max_cols, max_kp
        self._shape = {}

    def close(self):
        if getattr(self, "_h", None):
            self._L.yavo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc < 0:
            raise YavoError("yavo error %d: %s" % (rc, self._L.yavo_last_error(self._h).decode()))
        return rc

    @property
    def kernel_launches(self):
        return int(self._L.yavo_kernel_launches(self._h))

    def sync(self):
        self._ck(self._L.yavo_sync(self._h))

    @property
    def stream(self):
        """cudaStream_t of the context as an integer (wrap with torch.cuda.ExternalStream to record events)."""
        return int(self._L.yavo_get_stream(self._h) or 0)

    KERNEL_CLASSES = ("repitch", "detect_blur", "compact_score", "select_topk", "brief", "match_partial", "match_reduce",
                      "filter_pairs", "pyr_down", "klt_track", "epipolar_inliers", "match_tc", "select_big")

    def set_profiling(self, on):
        self._ck(self._L.yavo_set_profiling(self._h, int(bool(on))))

    def profile_collect(self):
        n = len(self.KERNEL_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int * n)()
        self._ck(self._L.yavo_profile_collect(self._h, ms, cnt, n))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    # ---- Image ----
    def upload(self, slot, img):
        img = np.ascontiguousarray(img, np.uint8)
        assert img.ndim == 2
        self._ck(self._L.yavo_upload(self._h, int(slot), _p(img), img.shape[0], img.shape[1], img.strides[0]))
        self._shape[slot] = img.shape

    def upload_batch(self, slot0, frames):
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.flags.c_contiguous
        n, H, W = frames.shape
        self._ck(self._L.yavo_upload_batch(self._h, int(slot0), n, _p(frames), H, W))
        for i in range(n):
            self._shape[slot0 + i] = (H, W)

    def upload_from_device(self, slot0, n, dev_ptr, rows, cols, pitch):
        self._ck(self._L.yavo_upload_from_device(self._h, int(slot0), int(n), C.c_void_p(int(dev_ptr)), int(rows),
                                                 int(cols), C.c_size_t(int(pitch))))
        for i in range(n):
            self._shape[slot0 + i] = (rows, cols)

    def download(self, slot):
        H, W = self._shape[slot]
        out = np.empty((H, W), np.uint8)
        self._ck(self._L.yavo_download(self._h, int(slot), _p(out), H, W))
        return out

    # ---- FastDetector ----
    def fast_detect(self, slot, max_kp=0):
        K = max_kp if max_kp > 0 else self.max_kp
        rows = np.empty(K, np.int32)
        cols = np.empty(K, np.int32)
        sc = np.empty(K, np.float32)
        n = C.c_int()
        nc = C.c_int()
        self._ck(self._L.yavo_fast_detect(self._h, int(slot), int(max_kp), _p(rows), _p(cols), _p(sc), C.byref(n),
                                          C.byref(nc)))
        return rows[:n.value].copy(), cols[:n.value].copy(), sc[:n.value].copy(), nc.value

    def frame_features(self, slot, img, max_kp=0):
        """getFastFeatures + computeBrief of one host frame in one call (include/yavo_b200.h: yavo_frame_features).
        Returns dict(rows, cols, scores [n_kp]; d_rows, d_cols, d_ids [n_desc]; desc [n_desc, 32]; n_cand)."""
        img = np.ascontiguousarray(img, np.uint8)
        K = max_kp if max_kp > 0 else self.max_kp
        r, c, s = np.empty(K, np.int32), np.empty(K, np.int32), np.empty(K, np.float32)
        dr, dc, di = np.empty(K, np.int32), np.empty(K, np.int32), np.empty(K, np.int32)
        desc = np.empty((K, 32), np.uint8)
        n, nd, nc = C.c_int32(), C.c_int32(), C.c_int()
        self._ck(self._L.yavo_frame_features(self._h, int(slot), _p(img), img.shape[0], img.shape[1], img.strides[0], int(max_kp),
                                             C.byref(n), _p(r), _p(c), _p(s), C.byref(nd), _p(dr), _p(dc), _p(di), _p(desc),
                                             C.byref(nc)))
        self._shape[slot] = img.shape
        n, nd = n.value, nd.value
        return dict(rows=r[:n].copy(), cols=c[:n].copy(), scores=s[:n].copy(), d_rows=dr[:nd].copy(), d_cols=dc[:nd].copy(),
                    d_ids=di[:nd].copy(), desc=desc[:nd].copy(), n_cand=nc.value)

    def fast_candidates(self, slot, cap=None):
        H, W = self._shape[slot]
        cap = cap or H * W
        rows = np.empty(cap, np.int32)
        cols = np.empty(cap, np.int32)
        sc = np.empty(cap, np.float32)
        nc = C.c_int()
        self._ck(self._L.yavo_fast_candidates(self._h, int(slot), int(cap), _p(rows), _p(cols), _p(sc), C.byref(nc)))
        n = min(nc.value, cap)
        return rows[:n].copy(), cols[:n].copy(), sc[:n].copy()

    # ---- Brief ----
    def set_brief_offsets(self, offsets):
        off = np.ascontiguousarray(offsets, np.int32).reshape(-1)
        assert off.size == 1024
        self._ck(self._L.yavo_set_brief_offsets(self._h, _p(off)))

    def blurred(self, slot):
        H, W = self._shape[slot]
        out = np.empty((H, W), np.uint8)
        self._ck(self._L.yavo_blurred(self._h, int(slot), _p(out), H, W))
        return out

    def brief_describe(self, slot, rows, cols):
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        n = rows.size
        desc = np.zeros((n, 32), np.uint8)
        valid = np.zeros(n, np.uint8)
        oob = C.c_int()
        self._ck(self._L.yavo_brief_describe(self._h, int(slot), _p(rows), _p(cols), n, _p(desc), _p(valid),
                                             C.byref(oob)))
        return desc, valid.astype(bool), oob.value

    def match(self, d1, d2, extensions=False):
        d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
        d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        n1, n2 = d1.shape[0], d2.shape[0]
        idx = np.zeros(n1, np.int32)
        dist = np.zeros(n1, np.int32)
        if not extensions:
            self._ck(self._L.yavo_match(self._h, _p(d1), n1, _p(d2), n2, _p(idx), _p(dist), None, None))
            return idx, dist
        sec = np.zeros(n1, np.int32)
        rev = np.zeros(n2, np.int32)
        self._ck(self._L.yavo_match(self._h, _p(d1), n1, _p(d2), n2, _p(idx), _p(dist), _p(sec), _p(rev)))
        return idx, dist, sec, rev

    # ---- batch ----
    def frontend_batch(self, slot0, n, do_match=True):
        self._ck(self._L.yavo_frontend_batch(self._h, int(slot0), int(n), int(do_match)))

    def alloc_batch_outputs(self, n, pinned=False):
        """Result arrays of the batch entry points; pinned=True puts them in page-locked memory, which is what lets
        submit_host_batch return before the device-to-host copies have run."""
        K = self.max_kp
        z = pinned_zeros if pinned else np.zeros
        return dict(n_kp=z(n, np.int32), rows=z((n, K), np.int32), cols=z((n, K), np.int32),
                    scores=z((n, K), np.float32), desc=z((n, K, 32), np.uint8),
                    match_idx=z((n, K), np.int32), match_dist=z((n, K), np.int32))

    def fetch_batch(self, slot0, n, out=None):
        out = out or self.alloc_batch_outputs(n)
        self._ck(self._L.yavo_fetch_batch(self._h, int(slot0), int(n), _p(out["n_kp"]), _p(out["rows"]), _p(out["cols"]),
                                          _p(out["scores"]), _p(out["desc"]), _p(out["match_idx"]),
                                          _p(out["match_dist"])))
        return out

    def fetch_batch_ptrs(self, slot0, n, ptrs):
        """fetch_batch into caller memory given as raw addresses (host, or device memory of this GPU — e.g. torch CUDA
        tensors' data_ptr()): dict with any of n_kp, rows, cols, scores, desc, match_idx, match_dist."""
        a = [C.c_void_p(int(ptrs[k])) if ptrs.get(k) else None
             for k in ("n_kp", "rows", "cols", "scores", "desc", "match_idx", "match_dist")]
        self._ck(self._L.yavo_fetch_batch(self._h, int(slot0), int(n), *a))

    def filter_pairs(self, slot0, n, threshold=20):
        """removeOutliers + point pairs on the device for the matches of the last frontend_batch."""
        n_pairs = np.zeros(n, np.int32)
        min_dist = np.zeros(n, np.int32)
        pairs = np.zeros((n, self.max_kp, 8), np.int32)
        self._ck(self._L.yavo_filter_pairs(self._h, int(slot0), int(n), int(threshold), _p(n_pairs), _p(min_dist),
                                           _p(pairs)))
        return n_pairs, min_dist, pairs

    def set_matcher(self, kind):
        """'tc' / 0 = tensor-core matcher on packed 4-bit operands (default), 'tc8' / 2 = on FP8 operands,
        'popc' / 1 = integer-pipe matcher; identical results."""
        self._ck(self._L.yavo_set_matcher(self._h, {"tc": 0, "popc": 1, "tc8": 2}.get(kind, kind)))

    def set_overlap(self, chunk_frames, n_streams=3):
        """Overlapped feature pipeline of the batch entry points (include/yavo_b200.h: yavo_set_overlap); 0 = serial."""
        self._ck(self._L.yavo_set_overlap(self._h, int(chunk_frames), int(n_streams)))

    def set_big_select(self, min_candidates):
        """Cluster pre-partition of large candidate lists (include/yavo_b200.h: yavo_set_big_select); -1 auto, 0 off."""
        self._ck(self._L.yavo_set_big_select(self._h, int(min_candidates)))

    def set_sub_batch(self, frames):
        self._ck(self._L.yavo_set_sub_batch(self._h, int(frames)))

    def set_pipeline_chunk(self, frames):
        self._ck(self._L.yavo_set_pipeline_chunk(self._h, int(frames)))

    def submit_host_batch(self, frames, do_match=True, out=None):
        """Asynchronous: frames must be pinned host memory; call wait() before reading `out`."""
        n, H, W = frames.shape
        out = out or self.alloc_batch_outputs(n)
        t = self._ck(self._L.yavo_submit_host_batch(self._h, _p(frames), n, H, W, int(do_match), _p(out["n_kp"]),
                                                    _p(out["rows"]), _p(out["cols"]), _p(out["scores"]), _p(out["desc"]),
                                                    _p(out["match_idx"]), _p(out["match_dist"])))
        for i in range(n):
            self._shape[i] = (H, W)
        return out, t

    def wait_batch(self, ticket):
        self._ck(self._L.yavo_wait_batch(self._h, int(ticket)))

    def wait(self):
        self._ck(self._L.yavo_wait(self._h))

    def process_host_batch(self, frames, do_match=True, out=None):
        n, H, W = frames.shape
        out = out or self.alloc_batch_outputs(n)
        self._ck(self._L.yavo_process_host_batch(self._h, _p(frames), n, H, W, int(do_match), _p(out["n_kp"]),
                                                 _p(out["rows"]), _p(out["cols"]), _p(out["scores"]), _p(out["desc"]),
                                                 _p(out["match_idx"]), _p(out["match_dist"])))
        for i in range(n):
            self._shape[i] = (H, W)
        return out

    # ---- tracking step: cv::calcOpticalFlowPyrLK (src/LoopHandler.cc:372-375) ----
    def build_pyramid(self, slot0, n, win=(11, 11), max_level=3):
        lv = C.c_int()
        self._ck(self._L.yavo_build_pyramid(self._h, int(slot0), int(n), int(win[0]), int(win[1]), int(max_level), C.byref(lv)))
        return lv.value

    def pyramid_level(self, slot, level):
        H, W = self._shape[slot]
        for _ in range(level):
            H, W = (H + 1) // 2, (W + 1) // 2
        out = np.empty((H, W), np.uint8)
        r, c = C.c_int(), C.c_int()
        self._ck(self._L.yavo_pyramid_level(self._h, int(slot), int(level), _p(out), C.c_size_t(out.size), C.byref(r), C.byref(c)))
        assert (r.value, c.value) == (H, W)
        return out

    def klt_track(self, slot_prev, slot_next, prev_pts, win=(11, 11), max_level=3, crit_type=3, max_count=30, epsilon=0.01,
                  flags=0, min_eig=1e-3, init_pts=None):
        """Points are OpenCV's (x = col, y = row) float32.  Returns next_pts (n,2), status (n,) u8, err (n,) f32."""
        pp = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
        n = pp.shape[0]
        nx = np.zeros((n, 2), np.float32)
        if init_pts is not None:
            nx[:] = np.asarray(init_pts, np.float32).reshape(-1, 2)
        st = np.zeros(n, np.uint8)
        er = np.zeros(n, np.float32)
        self._ck(self._L.yavo_klt_track(self._h, int(slot_prev), int(slot_next), _p(pp), n, _p(nx), _p(st), _p(er),
                                        int(win[0]), int(win[1]), int(max_level), int(crit_type), int(max_count),
                                        C.c_double(epsilon), int(flags), C.c_double(min_eig)))
        return nx, st, er

    def klt_track_batch(self, slot0, n, win=(11, 11), max_level=3, crit_type=3, max_count=30, epsilon=0.01, flags=0,
                        min_eig=1e-3):
        self._ck(self._L.yavo_klt_track_batch(self._h, int(slot0), int(n), int(win[0]), int(win[1]), int(max_level),
                                              int(crit_type), int(max_count), C.c_double(epsilon), int(flags),
                                              C.c_double(min_eig)))

    def klt_fetch(self, slot0, n):
        K = self.max_kp
        xy = np.zeros((n, K, 2), np.float32)
        st = np.zeros((n, K), np.uint8)
        er = np.zeros((n, K), np.float32)
        self._ck(self._L.yavo_klt_fetch(self._h, int(slot0), int(n), _p(xy), _p(st), _p(er)))
        return xy, st, er

    def stream_tracking(self, enable=True, win=(11, 11), max_level=3, crit_type=3, max_count=30, epsilon=0.01, flags=0,
                        min_eig=1e-3):
        """Tracking inside submit_host_batch / process_host_batch (pinned frames): see yavo_stream_tracking."""
        self._ck(self._L.yavo_stream_tracking(self._h, int(bool(enable)), int(win[0]), int(win[1]), int(max_level),
                                              int(crit_type), int(max_count), C.c_double(epsilon), int(flags),
                                              C.c_double(min_eig)))

    def alloc_track_outputs(self, n, pinned=False):
        K = self.max_kp
        z = pinned_zeros if pinned else np.zeros
        return dict(xy=z((n, K, 2), np.float32), status=z((n, K), np.uint8), err=z((n, K), np.float32))

    def stream_track_outputs(self, tracks):
        """Host arrays (alloc_track_outputs) the NEXT submit writes its tracks to; None detaches them."""
        if tracks is None:
            self._ck(self._L.yavo_stream_track_outputs(self._h, None, None, None))
        else:
            self._ck(self._L.yavo_stream_track_outputs(self._h, _p(tracks["xy"]), _p(tracks["status"]), _p(tracks["err"])))

    # ---- inlier count of the F-matrix RANSAC (src/3DHandler.cc:163-190) ----
    def epipolar_inliers(self, F, x1, y1, x2, y2, threshold=0.1, residuals=False):
        F = np.ascontiguousarray(F, np.float64).reshape(-1, 9)
        a = [np.ascontiguousarray(v, np.int32) for v in (x1, y1, x2, y2)]
        m, n = F.shape[0], a[0].size
        counts = np.zeros(m, np.int32)
        res = np.zeros((m, n), np.float64) if residuals else None
        best, bc = C.c_int32(), C.c_int32()
        self._ck(self._L.yavo_epipolar_inliers(self._h, _p(F), m, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), n,
                                               C.c_double(threshold), _p(counts), C.byref(best), C.byref(bc), _p(res)))
        return (counts, best.value, bc.value, res) if residuals else (counts, best.value, bc.value)


_PINNED_TYPES = {}


def _pinned_array_type(nbytes):
    """ctypes array type whose instances free their page-locked block when collected.  numpy arrays made with
    np.frombuffer keep the ctypes instance alive as their (transitive) base — including every slice and view derived
    later — so the block lives exactly as long as anything can still read it."""
    t = _PINNED_TYPES.get(nbytes)
    if t is None:
        class _Pinned(C.c_uint8 * nbytes):
            _free = None

            def __del__(self):
                free, self._free = self._free, None
                if free is not None:
                    try:
                        free(C.addressof(self))
                    except Exception:
                        pass
        t = _PINNED_TYPES[nbytes] = _Pinned
    return t


def _owned_view(ptr, nbytes, free):
    """uint8 numpy array over [ptr, ptr + nbytes) that calls free(ptr) when the last view of it is gone."""
    blk = _pinned_array_type(nbytes).from_address(ptr)
    blk._free = free
    return np.frombuffer(blk, dtype=np.uint8, count=nbytes)


def pinned_zeros(shape, dtype=np.uint8):
    """numpy array in page-locked host memory (cudaMallocHost through the C ABI), zero-filled."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    if n == 0:
        return np.zeros(shape, dt)
    L = lib()
    ptr = L.yavo_pinned_alloc(n)
    if not ptr:
        raise YavoError("yavo_pinned_alloc(%d) failed" % n)
    a = _owned_view(ptr, n, L.yavo_pinned_free).view(dt).reshape(shape)
    a[...] = 0
    return a


def ring_points(xc, yc):
    out = np.zeros(32, np.int32)
    lib().yavo_ring_points(int(xc), int(yc), _p(out))
    return out.reshape(16, 2)


def remove_outliers(dist, threshold=20):
    dist = np.ascontiguousarray(dist, np.int32)
    keep = np.zeros(dist.size, np.uint8)
    lib().yavo_remove_outliers(_p(dist), dist.size, int(threshold), _p(keep))
    return keep.astype(bool)
