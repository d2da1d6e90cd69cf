"""BASELINE config 3: a KITTI seq-00-length batch (4541 frames, 1241x376) device-resident on one GPU (or sharded over
WORLD_SIZE ranks with ya_vo_b200.sharding), FAST+BRIEF on each and the match f-1 -> f (timing only; the parity check of this configuration lives in tests/).
Prints one JSON object; commit it under profiles/."""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ya_vo_b200 import capi, synth  # noqa: E402

F, H, W = 4541, 376, 1241


def gen(rng):
    a, b = rng
    return synth.synth_batch(b - a, "G30", 1000 + a, H, W)


def main():
    t0 = time.perf_counter()
    cuts = np.linspace(0, F, 33).astype(int)
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        frames = np.concatenate(list(ex.map(gen, zip(cuts[:-1], cuts[1:]))))
    t_gen = time.perf_counter() - t0
    with capi.Context(device=0, n_slots=F, max_rows=H, max_cols=W, max_kp=2000, max_cand=32768) as ctx:
        ctx.set_brief_offsets(synth.brief_offsets())
        t0 = time.perf_counter()
        ctx.upload_batch(0, frames)
        ctx.sync()
        t_up = time.perf_counter() - t0
        ctx.frontend_batch(0, F, True)  # warm
        ctx.sync()
        ctx.set_profiling(True)
        t0 = time.perf_counter()
        ctx.frontend_batch(0, F, True)
        ctx.sync()
        dt = time.perf_counter() - t0
        prof = ctx.profile_collect()
        t0 = time.perf_counter()
        out = ctx.fetch_batch(0, F)
        t_fetch = time.perf_counter() - t0
    print(json.dumps({"config": "seq-00-length batch: %d frames %dx%d, device-resident, one launch set" % (F, W, H),
                      "frames_per_s": F / dt, "ms_total": 1e3 * dt, "upload_s": t_up, "fetch_s": t_fetch, "generate_s": t_gen,
                      "mean_keypoints": float(out["n_kp"].mean()), "kernel_ms": {k: v[0] for k, v in prof.items()},
                      "note": "parity of this configuration is covered by tests/test_gpu_fullsize.py::test_seq00_length_batch"}))


if __name__ == "__main__":
    main()
