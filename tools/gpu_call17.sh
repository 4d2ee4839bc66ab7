#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/exp/gt_profile.py --plain 2>&1 | grep "^==\|^##\|Error\|error" | head -30
timeout 300 python tools/exp/gt_profile.py > gpurun_out/gt_profile3.txt 2>&1; head -75 gpurun_out/gt_profile3.txt | grep -v "prologue\|emptyA"
timeout 700 python tests/run_gpu_checks.py --out gpurun_out/checks_r2m.json > gpurun_out/checks_r2m.log 2>&1
grep -c "^ok" gpurun_out/checks_r2m.log
grep "FAIL\|EXCEPTION\|TOTAL\|dead\|Error" gpurun_out/checks_r2m.log | head -30
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-port > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2m.json"))
print("s2", d["ms_per_step"], d["e2e"].get("ms_per_step"), d["gpu_launches"], d["roofline"]["achieved"], d["roofline"]["frac"])
g = d["gpt"]; print("gpt", g.get("ms_per_step"), g.get("value"), g.get("roofline", {}).get("achieved"), g.get("error"))
PY
tail -3 gpurun_out/bench_r2m.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_r2m.csv python tests/prof_s2.py > gpurun_out/ncu_s2.log 2>&1; tail -1 gpurun_out/ncu_s2.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gpt_launches_r2m.csv python tests/prof_gpt.py > gpurun_out/ncu_gpt.log 2>&1; tail -1 gpurun_out/ncu_gpt.log
