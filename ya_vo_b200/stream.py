"""Streaming caller of the front end — the B200 counterpart of LoopHandler::getNextFrame + insertFrameFeatures
(reference src/LoopHandler.cc:917-927, 468-485).

The reference reads one frame (`cv::imread(path, 0)`), runs FAST + BRIEF on it, then reads the next.  Here a
background thread decodes the upcoming frames straight into pinned host batches while the GPU works on the current
batch, and consecutive batches overlap through yavo_submit_host_batch / yavo_wait_batch: decode, PCIe upload,
kernels and result download all run concurrently.  Frames keep their sequence order; batch seams re-use one frame so
every consecutive pair (f-1, f) is matched exactly once.
"""
import os
import queue
import threading

import numpy as np


def list_frames(directory, exts=(".png", ".jpg", ".jpeg", ".pgm", ".bmp")):
    """Sorted image paths of a KITTI-style `image_0/` directory (the reference sorts the listing: src/Utils.cc)."""
    return sorted(os.path.join(directory, f) for f in os.listdir(directory) if f.lower().endswith(exts))


def _imread_gray(path):
    import cv2
    img = cv2.imread(path, 0)  # the reference's cv::imread(path, 0): 8-bit single channel
    if img is None:
        raise IOError("cannot read " + path)
    return img


class PinnedBatches:
    """`depth` pinned (page-locked) uint8 buffers of shape [batch, H, W] handed round-robin to the decoder."""

    def __init__(self, depth, batch, H, W):
        from . import capi
        self.buffers = [capi.pinned_zeros((batch, H, W), np.uint8) for _ in range(depth)]


class FrameStream:
    """Iterates over a sequence in batches; decoding runs `prefetch` batches ahead on a background thread.

    read(i) -> HxW uint8 array for frame i (default: cv2.imread of paths[i]).  Batches after the first start with the
    last frame of the previous batch (the seam frame)."""

    def __init__(self, n_frames, read, batch, shape, prefetch=2, buffers=None):
        if batch < 2:
            # every batch after the first re-uses one frame as its seam: a 1-frame batch would never advance
            raise ValueError("FrameStream needs batch >= 2 (got %d)" % batch)
        self.n, self.read, self.batch, self.shape = n_frames, read, batch, shape
        self.buffers = buffers or [np.empty((batch,) + tuple(shape), np.uint8) for _ in range(prefetch + 2)]
        self._q = queue.Queue(maxsize=prefetch)
        self._free = queue.Queue()
        for i in range(len(self.buffers)):
            self._free.put(i)
        self._th = threading.Thread(target=self._worker, daemon=True)
        self._err = None
        self._th.start()

    @classmethod
    def from_directory(cls, directory, batch, prefetch=2, pinned=True):
        paths = list_frames(directory)
        if not paths:
            raise IOError("no frames in " + directory)
        H, W = _imread_gray(paths[0]).shape
        bufs = PinnedBatches(prefetch + 2, batch, H, W) if pinned else None
        s = cls(len(paths), lambda i: _imread_gray(paths[i]), batch, (H, W), prefetch, bufs.buffers if bufs else None)
        s._pinned_owner = bufs
        return s

    def _worker(self):
        try:
            a = 0
            while a < self.n:
                b = min(self.n, a + self.batch)
                bi = self._free.get()
                buf = self.buffers[bi]
                for k, f in enumerate(range(a, b)):
                    img = self.read(f)
                    if img.shape != tuple(self.shape) or img.dtype != np.uint8:
                        raise ValueError("frame %d is %s %s, expected %s uint8" % (f, img.shape, img.dtype, self.shape))
                    buf[k] = img
                self._q.put((a, b, bi))
                if b >= self.n:
                    break
                a = b - 1  # seam: the next batch starts with this batch's last frame
            self._q.put(None)
        except Exception as e:  # surfaced to the consumer
            self._err = e
            self._q.put(None)

    def __iter__(self):
        while True:
            item = self._q.get()
            if item is None:
                if self._err:
                    raise self._err
                return
            a, b, bi = item
            yield a, b, self.buffers[bi][: b - a], bi

    def release(self, bi):
        self._free.put(bi)


def run_sequence(ctx, stream, do_match=True, on_frame=None, on_tracks=None, track_params=None):
    """Pushes a FrameStream through the front end with one batch in flight ahead of the one being consumed.

    on_frame(f, n_kp, rows, cols, scores, desc, match_idx, match_dist) is called once per frame in order (match_* are
    None for frame 0; they describe the pair (f-1, f) and index frame f-1's / frame f's keypoints).
    on_tracks(f, next_xy, status, err), if given, switches on the tracking step of the reference's steady-state loop
    (cv::calcOpticalFlowPyrLK, src/LoopHandler.cc:372-375) inside the same pipeline: frame f-1's top-K FAST keypoints
    (before checkBoundry, (x, y) = (col, row)) tracked into frame f, called once per f >= 1 after on_frame(f, ...);
    track_params are keyword arguments of Context.stream_tracking (default: the reference's 11x11 / 3 levels / 30 / 0.01).
    Returns the number of frames delivered."""
    pending = None  # (a, b, out, tracks, ticket, buffer index)
    delivered = 0
    tracking = on_tracks is not None
    if tracking:
        ctx.stream_tracking(True, **(track_params or {}))

    def deliver(a, b, out, trk):
        nonlocal delivered
        for i in range(b - a):
            f = a + i
            if i == 0 and a > 0:
                continue  # the seam frame was delivered with the previous batch
            k = int(out["n_kp"][i])
            if on_frame:
                mi = md = None
                if f > 0 and do_match:
                    kq = int(out["n_kp"][i - 1])
                    mi, md = out["match_idx"][i, :kq], out["match_dist"][i, :kq]
                on_frame(f, k, out["rows"][i, :k], out["cols"][i, :k], out["scores"][i, :k], out["desc"][i, :k], mi, md)
            if tracking and i > 0:
                on_tracks(f, trk["xy"][i - 1], trk["status"][i - 1], trk["err"][i - 1])
            delivered += 1

    # page-locked result arrays: the device-to-host copies of a submit run behind the call, not inside it
    outs = [ctx.alloc_batch_outputs(stream.batch, pinned=True) for _ in range(2)]
    trks = [ctx.alloc_track_outputs(stream.batch, pinned=True) for _ in range(2)] if tracking else [None, None]
    turn = 0
    try:
        for a, b, frames, bi in stream:
            if tracking:
                ctx.stream_track_outputs(trks[turn])
            out, ticket = ctx.submit_host_batch(frames, do_match, outs[turn])
            if pending is not None:
                pa, pb, pout, ptrk, pt, pbi = pending
                ctx.wait_batch(pt)
                stream.release(pbi)
                deliver(pa, pb, pout, ptrk)
            pending = (a, b, out, trks[turn], ticket, bi)
            turn ^= 1
        if pending is not None:
            pa, pb, pout, ptrk, pt, pbi = pending
            ctx.wait()
            stream.release(pbi)
            deliver(pa, pb, pout, ptrk)
    finally:
        if tracking:
            ctx.wait()
            ctx.stream_track_outputs(None)
            ctx.stream_tracking(False)
    return delivered
