#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/run_gpu_checks.py --only side_streams 2>&1 | grep "^ok\|FAIL\|TOTAL\|EXC\|Error" | head
for dl in 3 6; do
  EVK_D_LANES=$dl timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-port --gpt 0 > gpurun_out/bench_r2o_dl$dl.json 2> gpurun_out/bench_r2o_dl$dl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2o_dl$dl.json"))
    r = d["roofline"]
    print("d_lanes=$dl s2", d["ms_per_step"], d["e2e"].get("ms_per_step"), d["gpu_launches"], "roofline", r["achieved"], r["frac"], r["ms_per_step"], r["share_of_step_time"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2o_dl$dl.err").read()[-2000:])
PY
done
