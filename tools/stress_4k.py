"""BASELINE config 4: 3840x2160 uniform-noise frames (~308k FAST candidates each), cap raised to 20,000 keypoints.
Device-resident batch, per-kernel device times.  Prints one JSON object; commit it under profiles/."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ya_vo_b200 import capi, synth  # noqa: E402


def main():
    B, H, W, K = 16, 2160, 3840, 20000
    big_min = int(sys.argv[1]) if len(sys.argv) > 1 else None  # yavo_set_big_select threshold (default: the library's)
    frames = np.stack([synth.synth_frame("U", 3 + f, H, W) for f in range(B)])
    with capi.Context(device=0, n_slots=B, max_rows=H, max_cols=W, max_kp=K) as ctx:
        ctx.set_brief_offsets(synth.brief_offsets())
        if big_min is not None:
            ctx.set_big_select(big_min)
        ctx.upload_batch(0, frames)
        for _ in range(3):
            ctx.frontend_batch(0, B, True)
        ctx.sync()
        ctx.set_profiling(True)
        t0 = time.perf_counter()
        steps = 5
        for _ in range(steps):
            ctx.frontend_batch(0, B, True)
        ctx.sync()
        dt = time.perf_counter() - t0
        prof = ctx.profile_collect()
        out = ctx.fetch_batch(0, B)
    print(json.dumps({"config": "4K stress: %d frames 3840x2160 uniform noise per step, max_kp %d, match f-1->f" % (B, K),
                      "big_select_min": big_min, "frames_per_s": B * steps / dt, "ms_per_step": 1e3 * dt / steps,
                      "mean_keypoints": float(out["n_kp"].mean()),
                      "kernel_ms_per_step": {k: v[0] / steps for k, v in prof.items()},
                      "algorithmic_bytes_per_frame": W * H + 44 * K,
                      "achieved_gbs_detect_describe": B * (W * H + 44 * K) / 1e9 /
                      (sum(prof[k][0] for k in ("detect_blur", "compact_score", "select_topk", "select_big", "brief")) / steps / 1e3)}))


if __name__ == "__main__":
    main()
