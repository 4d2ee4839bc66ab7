#!/bin/bash
# round-2 GPU call: parity with the multi-stream forward, side-stream A/B, GPT bench
mkdir -p gpurun_out
timeout 700 python tests/run_gpu_checks.py --out gpurun_out/checks_r2k.json > gpurun_out/checks_r2k.log 2>&1
grep -c "^ok" gpurun_out/checks_r2k.log
grep "FAIL\|EXCEPTION\|TOTAL\|dead\|Error" gpurun_out/checks_r2k.log | head -30
for ss in 0 1; do
  EVK_SIDE_STREAMS=$ss timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-port --gpt 0 > gpurun_out/bench_r2k_ss$ss.json 2> gpurun_out/bench_r2k_ss$ss.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2k_ss$ss.json"))
    print("side_streams=$ss", d["ms_per_step"], d["e2e"].get("ms_per_step"), d["gpu_launches"], d["roofline"]["achieved"], d["losses"])
except Exception as e:
    print("side_streams=$ss failed", e); print(open("gpurun_out/bench_r2k_ss$ss.err").read()[-1500:])
PY
done
timeout 400 python bench.py --only-gpt --steps 20 --warmup 5 > gpurun_out/bench_r2k_gpt.json 2> gpurun_out/bench_r2k_gpt.err
tail -c 1500 gpurun_out/bench_r2k_gpt.json; tail -5 gpurun_out/bench_r2k_gpt.err
