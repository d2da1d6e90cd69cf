"""Per-kernel device time of the single-frame entry points (yavo_fast_detect / yavo_brief_describe / yavo_match)."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ya_vo_b200 import capi, synth

a = synth.synth_frame("G30", 77)
b = synth.shifted_pair(a, 78)
with capi.Context(device=0, n_slots=2, max_rows=376, max_cols=1241, max_kp=2000) as ctx:
    ctx.set_brief_offsets(synth.brief_offsets())
    for rep in range(3):
        ctx.set_profiling(True)
        t0 = time.perf_counter(); ctx.upload(0, a); r, c, s, nc = ctx.fast_detect(0); t1 = time.perf_counter()
        d1, v1, _ = ctx.brief_describe(0, r, c); t2 = time.perf_counter()
        ctx.upload(1, b); r2, c2, s2, _ = ctx.fast_detect(1); d2, v2, _ = ctx.brief_describe(1, r2, c2)
        t3 = time.perf_counter(); idx, dist = ctx.match(d1[v1], d2[v2]); t4 = time.perf_counter()
        prof = ctx.profile_collect()
        print("wall us: upload+fast %.0f brief %.0f match %.0f | device us per class (2 frames):" % ((t1-t0)*1e6, (t2-t1)*1e6, (t4-t3)*1e6),
              {k: round(v[0]*1e3, 1) for k, v in prof.items()})
