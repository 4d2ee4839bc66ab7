#!/bin/bash
mkdir -p gpurun_out
for ab in 0 1; do
  EVK_TMA_AUTO_BN=$ab timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-port > gpurun_out/bench_r2n_bn$ab.json 2> gpurun_out/bench_r2n_bn$ab.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2n_bn$ab.json"))
print("auto_bn=$ab s2", d["ms_per_step"], d["e2e"].get("ms_per_step"), d["gpu_launches"], d["roofline"]["achieved"], d["roofline"]["frac"])
g = d["gpt"]; print("auto_bn=$ab gpt", g.get("ms_per_step"), g.get("value"), g.get("roofline", {}).get("achieved"), g.get("error"))
PY
done
timeout 700 python tests/run_gpu_checks.py --out gpurun_out/checks_r2n.json > gpurun_out/checks_r2n.log 2>&1
grep -c "^ok" gpurun_out/checks_r2n.log
grep "FAIL\|EXCEPTION\|TOTAL\|dead\|Error" gpurun_out/checks_r2n.log | head -30
