import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from ya_vo_b200 import capi, synth
frames = synth.synth_batch(8, "G30", 1000)
off = synth.brief_offsets()
with capi.Context(device=0, n_slots=2, max_rows=376, max_cols=1241, max_kp=2000) as ctx:
    ctx.set_brief_offsets(off)
    ref = None
    for bm in (0, 4096, 2048, 1024, 512):
        ctx.set_big_select(bm)
        for i in range(20):
            ff = ctx.frame_features(i & 1, frames[i % 8])
        if ref is None:
            ref = [ctx.frame_features(0, frames[j])["desc"].copy() for j in range(8)]
        else:
            for j in range(8):
                assert np.array_equal(ctx.frame_features(0, frames[j])["desc"], ref[j])
        t = []
        for i in range(300):
            t0 = time.perf_counter()
            ctx.frame_features(i & 1, frames[i % 8])
            t.append(time.perf_counter() - t0)
        t = np.sort(np.array(t)) * 1e6
        print("big_min", bm, "p50 %.1f us  p10 %.1f  p90 %.1f" % (t[150], t[30], t[270]))
