"""Tracking step (cv::calcOpticalFlowPyrLK at src/LoopHandler.cc:372-375) on a batch of device-resident
frames: every frame's top-2000 FAST keypoints are tracked into the next frame with the reference's parameters
(11x11 window, 3 pyramid levels, COUNT+EPS 30 / 0.01, minEig 0.001).  Frame f+1 is frame f moved by a small random
shift plus noise (synth.shifted_pair), so the tracker has real motion to recover.  Times the pyramid and tracking
kernels with CUDA events and, on the same inputs, OpenCV's own CPU implementation (cv2, if importable; timing only).
Prints one JSON object; commit it under profiles/."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ya_vo_b200 import capi, synth  # noqa: E402

H, W = 376, 1241


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kind", default="B4")
    ap.add_argument("--cpu-pairs", type=int, default=32)
    args = ap.parse_args()
    B = args.batch
    if os.environ.get("YAVO_REBUILD"):
        capi.build(force=True)  # tuning runs: YAVO_NVCC_EXTRA=-D... YAVO_REBUILD=1
    frames = np.empty((B, H, W), np.uint8)
    if args.kind == "K":  # the real KITTI-shaped frame of the reference's tests (tests/golden/kitti_frame.png)
        import cv2
        frames[0] = cv2.imread(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "kitti_frame.png"), 0)
    else:
        frames[0] = synth.synth_frame(args.kind, 500, H, W)
    for f in range(1, B):
        frames[f] = synth.shifted_pair(frames[f - 1], 500 + f)
    with capi.Context(device=0, n_slots=B, max_rows=H, max_cols=W, max_kp=2000, max_cand=65536) as ctx:
        ctx.set_brief_offsets(synth.brief_offsets())
        ctx.upload_batch(0, frames)
        ctx.frontend_batch(0, B, False)
        ctx.sync()
        for _ in range(2):  # warm-up (first call also allocates the pyramid)
            ctx.upload_batch(0, frames)  # re-upload invalidates the pyramids, so every step rebuilds them
            ctx.klt_track_batch(0, B)
        ctx.sync()
        t_tot = 0.0
        prof_sum = {}
        for _ in range(args.steps):
            ctx.upload_batch(0, frames)
            ctx.sync()
            ctx.set_profiling(True)
            t0 = time.perf_counter()
            ctx.klt_track_batch(0, B)
            ctx.sync()
            t_tot += time.perf_counter() - t0
            for k, v in ctx.profile_collect().items():
                prof_sum[k] = prof_sum.get(k, 0.0) + v[0]
            ctx.set_profiling(False)
        out = ctx.fetch_batch(0, B)
        xy, st, er = ctx.klt_fetch(0, B)
    nk = out["n_kp"][:-1].astype(np.int64)
    pts = int(nk.sum())
    tracked = int(sum(int(st[f, :nk[f]].sum()) for f in range(B - 1)))
    pyr_ms = prof_sum["pyr_down"] / args.steps
    klt_ms = prof_sum["klt_track"] / args.steps
    res = {"config": "%d consecutive %dx%d frames (%s, shifted pairs), top-2000 FAST keypoints of frame f tracked into f+1; "
                     "window 11x11, 3 levels, 30 iterations / 0.01, minEig 0.001" % (B, W, H, args.kind),
           "points_per_step": pts, "tracked_fraction": tracked / max(pts, 1),
           "pyramid_ms_per_step": pyr_ms, "track_ms_per_step": klt_ms, "wall_ms_per_step": 1e3 * t_tot / args.steps,
           "frames_per_s_kernels": (B - 1) / ((pyr_ms + klt_ms) * 1e-3),
           "points_per_s_track_kernel": pts / (klt_ms * 1e-3),
           "us_per_frame": {"pyramid": 1e3 * pyr_ms / B, "track": 1e3 * klt_ms / (B - 1)}}
    try:
        import cv2
        cv2.setNumThreads(1)
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
        n = min(args.cpu_pairs, B - 1)
        t0 = time.perf_counter()
        for f in range(n):
            p = np.stack([out["cols"][f, :nk[f]], out["rows"][f, :nk[f]]], 1).astype(np.float32)
            cv2.calcOpticalFlowPyrLK(frames[f], frames[f + 1], p, None, winSize=(11, 11), maxLevel=3, criteria=crit, flags=0,
                                     minEigThreshold=0.001)
        dt = time.perf_counter() - t0
        res["cpu_opencv"] = {"frames_per_s": n / dt, "threads": 1, "sample": "%d frame pairs, cv2 %s calcOpticalFlowPyrLK "
                             "(pyramids built inside the call)" % (n, cv2.__version__)}
    except ImportError:
        res["cpu_opencv"] = None
    print(json.dumps(res))


if __name__ == "__main__":
    main()
