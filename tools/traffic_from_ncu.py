#!/usr/bin/env python
"""profiles/traffic.json from an `ncu -i <rep> --page raw --csv` dump of one bench.py step (B frames per launch):
DRAM bytes and warp instructions per frame by kernel.   usage: traffic_from_ncu.py <raw.csv> <frames per launch> <source note>"""
import csv
import json
import sys

raw, B, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lines = open(raw).read().splitlines()
i = next(k for k, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.reader(lines[i:]))
hdr, units = rows[0], rows[1]
idx = {n: k for k, n in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram, insts, seen = {}, {}, set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("yavo::", "")
    if name in seen:
        continue
    seen.add(name)
    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    dram[name] = {"read": rd / B, "write": wr / B, "total": (rd + wr) / B}
    insts[name] = float(r[idx["smsp__inst_executed.sum"]]) / B
dd = sum(v["total"] for k, v in dram.items() if "match" not in k)
print(json.dumps({"source": note, "frames_per_launch": B, "dram_bytes_per_frame_by_kernel": dram,
                  "detect_describe_dram_bytes_per_frame": dd, "warp_insts_per_frame_by_kernel": insts}, indent=1))
