"""Small end-to-end case for compute-sanitizer (one tool per GPU session): the batch path with the cluster pre-partition
forced on, the single-frame graph path, the matcher with both extensions, the tracking step.  Compared with the oracle."""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402
from ya_vo_b200 import capi, synth  # noqa: E402

off = synth.brief_offsets()
H, W, K = 136, 333, 400
frames = synth.synth_batch(4, "U", 11, H, W)
frames[2] = synth.shifted_pair(frames[1], 5)
exp = po.pipeline(frames, off, K, True, nthreads=2)
with capi.Context(device=0, n_slots=4, max_rows=H, max_cols=W, max_kp=K) as ctx:
    ctx.set_brief_offsets(off)
    for big in (0, 600):
        ctx.set_big_select(big)
        out = ctx.process_host_batch(frames, True)
        assert np.array_equal(out["n_kp"], exp["n_kp"]), big
        for f in range(4):
            k = exp["n_kp"][f]
            assert np.array_equal(out["desc"][f, :k], exp["desc"][f, :k]) and np.array_equal(out["rows"][f, :k], exp["rows"][f, :k])
            if f:
                kq = exp["n_kp"][f - 1]
                assert np.array_equal(out["match_idx"][f, :kq], exp["match_idx"][f, :kq])
    for rep in range(2):
        ff = ctx.frame_features(0, frames[3], K)
        k = exp["n_kp"][3]
        assert len(ff["desc"]) == k and np.array_equal(ff["desc"], exp["desc"][3, :k])
    d1, d2 = synth.synth_descriptors(300, 1), synth.synth_descriptors(500, 2)
    got = ctx.match(d1, d2, extensions=True)
    for a, b in zip(got, po.match(d1, d2, extensions=True)):
        assert np.array_equal(a, b)
    ctx.upload_batch(0, frames)
    ctx.frontend_batch(0, 4, True)
    ctx.klt_track_batch(0, 4)
    ctx.klt_fetch(0, 4)
print("sanitize case ok")
