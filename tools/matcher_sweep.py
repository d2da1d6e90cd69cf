"""BASELINE config 5: Hamming matcher sweep, N1 x N2 256-bit descriptors, device time of the match kernels
(CUDA events around each launch via yavo_set_profiling).  Prints one JSON object; commit it under profiles/."""
import json
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ya_vo_b200 import capi, synth  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1024, 2048, 4096, 8192, 16384, 32768, 65536]
    out = {"unit": "Gpairs/s", "note": "device time of the match kernels (match_tc, or match_partial + match_reduce), descriptors "
           "already on the device side of yavo_match's H2D; matcher = tc (tcgen05 kind::mxf4, default), tc8 (kind::f8f6f4) or "
           "popc (integer pipes); ext=1: second-best distance (from the tcgen05 epilogue since round 2) + cross-check (a second "
           "launch with the roles swapped: 2 x N1 x N2 pairs counted)",
           "rows": []}
    with capi.Context(device=0, n_slots=1, max_rows=64, max_cols=128, max_kp=16) as ctx:
        for n1 in sizes:
            for n2 in sizes:
                if n1 * n2 > 65536 * 65536:
                    continue
                d1 = synth.synth_descriptors(n1, n1 * 131 + n2)
                d2 = synth.synth_descriptors(n2, n1 * 131 + n2 + 1)
                for matcher, ext in (("tc", 0), ("tc8", 0), ("popc", 0), ("tc", 1)):
                    ctx.set_matcher(matcher)
                    ctx.match(d1, d2, extensions=False)  # warm
                    ctx.set_profiling(True)
                    reps = 3 if n1 * n2 < (1 << 30) else 1
                    for _ in range(reps):
                        if ext:
                            idx, dist, sec, rev = ctx.match(d1, d2, extensions=True)
                        else:
                            idx, dist = ctx.match(d1, d2)
                    prof = ctx.profile_collect()
                    ctx.set_profiling(False)
                    ms = (prof["match_tc"][0] + prof["match_partial"][0] + prof["match_reduce"][0]) / reps
                    pairs = n1 * n2 * (2 if ext else 1)  # the cross-check runs the kernel a second time, roles swapped
                    out["rows"].append({"n1": n1, "n2": n2, "matcher": matcher, "ext": ext, "ms": ms, "gpairs": pairs / ms / 1e6})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
