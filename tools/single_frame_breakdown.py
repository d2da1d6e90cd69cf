"""Where one yavo_frame_features call (the reference's per-frame shape, LoopHandler::insertFrameFeatures) spends its time:
the call's wall time, the device time of each kernel class (profiling mode: plain launches with event pairs), and the
host-side staging copy measured separately."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ya_vo_b200 import capi, synth

frames = synth.synth_batch(8, "G30", 1000)
with capi.Context(device=0, n_slots=2, max_rows=376, max_cols=1241, max_kp=2000) as ctx:
    ctx.set_brief_offsets(synth.brief_offsets())
    for i in range(30):
        ctx.frame_features(i & 1, frames[i % 8])
    t = []
    for i in range(500):
        t0 = time.perf_counter()
        ctx.frame_features(i & 1, frames[i % 8])
        t.append(time.perf_counter() - t0)
    t = np.sort(np.array(t)) * 1e6
    print("frame_features wall: p50 %.1f us  p10 %.1f  p90 %.1f" % (t[250], t[50], t[450]))
    ctx.set_profiling(True)
    for i in range(50):
        ctx.frame_features(i & 1, frames[i % 8])
    prof = ctx.profile_collect()
    print("device us per call by kernel class:", {k: round(v[0] * 1e3 / 50, 1) for k, v in prof.items() if v[0] > 0})
    ctx.set_profiling(False)
dst = capi.pinned_zeros(frames[0].shape, np.uint8)
t = []
for i in range(500):
    t0 = time.perf_counter()
    np.copyto(dst, frames[i % 8])
    t.append(time.perf_counter() - t0)
print("host copy of one frame into pinned memory: p50 %.1f us" % (np.sort(np.array(t))[250] * 1e6))
