#!/bin/bash
# BASELINE configs[0]/[1]: one 1241x376 frame (pair) through the C++ drop-in classes, with the timing printouts of
# the reference's tests/BriefDescriptorTest.cc:21-32.  Usage: tools/single_frame_latency.sh  (needs a GPU)
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import numpy as np
from ya_vo_b200 import synth
a = synth.synth_frame("G30", 77); b = synth.shifted_pair(a, 78)
a.tofile("/tmp/yavo_a.bin"); b.tofile("/tmp/yavo_b.bin"); synth.brief_offsets().astype(np.int32).tofile("/tmp/yavo_off.bin")
PY
for i in 1 2 3; do ya_vo_b200/host/host_tests pipeline /tmp/yavo_a.bin /tmp/yavo_b.bin 376 1241 /tmp/yavo_off.bin /tmp/yavo_out.bin; done
# the tracking step of the steady-state loop (src/LoopHandler.cc:372-375) on the same pair
for i in 1 2 3; do ya_vo_b200/host/host_tests track /tmp/yavo_a.bin /tmp/yavo_b.bin 376 1241 /tmp/yavo_trk.bin; done
