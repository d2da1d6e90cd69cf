timeout 300 python tools/matcher_sweep.py > gpurun_out/matcher_sweep_r1k.json 2> gpurun_out/sweep.err; tail -c 300 gpurun_out/sweep.err
timeout 200 python tools/stress_4k.py > gpurun_out/stress_4k_r1k.json 2> gpurun_out/stress.err; tail -c 300 gpurun_out/stress.err
timeout 300 python tools/seq00_batch.py > gpurun_out/seq00_batch_r1k.json 2> gpurun_out/seq.err; tail -c 300 gpurun_out/seq.err
python - <<PY
import json
d=json.load(open("gpurun_out/matcher_sweep_r1k.json"))
for r in d["rows"]:
    if r["n1"]==r["n2"]: print(r["n1"], r["matcher"], r["ext"], round(r["ms"],4), round(r["gpairs"],1))
print(open("gpurun_out/stress_4k_r1k.json").read()[:900])
print(open("gpurun_out/seq00_batch_r1k.json").read()[:900])
PY
