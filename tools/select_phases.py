"""Debug: per-phase cycle counts of the select kernel (build with YAVO_NVCC_EXTRA=-DYAVO_SEL_TIMING)."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ya_vo_b200 import capi, synth
# usage: select_phases.py [kind [H W K frames]]
H, W, K, NF = (int(x) for x in sys.argv[2:6]) if len(sys.argv) > 5 else (376, 1241, 2000, 64)
frames = synth.synth_batch(NF, sys.argv[1] if len(sys.argv) > 1 else "G30", 1000, H, W)
with capi.Context(device=0, n_slots=NF, max_rows=H, max_cols=W, max_kp=K) as ctx:
    ctx.set_brief_offsets(synth.brief_offsets())
    ctx.upload_batch(0, frames)
    for n in (NF, 1):
        for _ in range(2):
            ctx.frontend_batch(0, n, False)
        out = np.zeros(64 * 8 + 8 * 16 * 4, np.int64)
        capi.lib().yavo_debug_select_timing(ctx._h, out.ctypes.data_as(C.c_void_p))
        w = out[64 * 8:].reshape(8, 16, 4)
        out = out[:64 * 8].reshape(64, 8)
        d = np.diff(out[:min(n, 64), :6], axis=1)
        print("frames in flight", n, "mean cycles per phase [load+score, phase1, switch, phase2, output]:", d.mean(axis=0).round(0), "total", d.sum(axis=1).mean().round(0))
        print("   frame 0 per warp: pop-wait cycles", w[0, :, 0].tolist(), "work cycles", w[0, :, 1].tolist(), "tasks", w[0, :, 2].tolist())
