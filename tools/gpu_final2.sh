#!/bin/bash
# final re-run after the mel staging change: tests, default bench, ncu --set full (gemm_tma incl. GPT linear1, mel, flash)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/final2_pytest.log 2>&1; tail -4 gpurun_out/final2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; tail -1 gpurun_out/final2_smoke.log | cut -c1-200
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final2.json ) 2> gpurun_out/bench_r2_final2.err; tail -4 gpurun_out/bench_r2_final2.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2_final2.json"))
print("s2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
r = d["roofline"]; print("roofline", r["achieved"], r["frac"], r["traffic"], r["share_of_step_time"])
print("mel", {k: v for k, v in d["mel_roofline"].items() if k in ("achieved", "frac", "traffic", "ms")}, d["mel_roofline"]["mel_only"]["frac"], d["mel_roofline"]["batch16"]["frac"])
print("cpu", d.get("cpu_baseline")["value"]); print("port16", d["torch_gpu_port_fp16_autocast"]["ms_per_step"], "port32", d["torch_gpu_port"]["ms_per_step"])
g = d["gpt"]; print("gpt", g.get("value"), g.get("ms_per_step"), g["e2e"]["value"], g.get("error")); print("gpt roofline", g["roofline"]["achieved"], g["roofline"]["frac"], g["roofline"]["share_of_step_time"], g["roofline"]["attention_ms"])
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'gemm_tma|mel_fwd_warp|flash' -o gpurun_out/prof_final2 -f python tests/prof_kernels.py > gpurun_out/ncu_prof_final2.log 2>&1; tail -1 gpurun_out/ncu_prof_final2.log; ls -la gpurun_out/prof_final2.ncu-rep
