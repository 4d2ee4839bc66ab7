#!/bin/bash
# round-2 final GPU pass: tests, smoke, default bench, reference arm, launch lists, ncu --set full, config 5, 32 kHz label
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; tail -4 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json ) 2> gpurun_out/bench_r2_final.err; tail -4 gpurun_out/bench_r2_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2_final.json"))
print("s2", d["value"], d["ms_per_step"], d["e2e"], d["gpu_launches"], d["clocks"])
r = d["roofline"]; print("roofline", r["achieved"], r["frac"], r["traffic"], r["share_of_step_time"])
print("mel", {k: v for k, v in d["mel_roofline"].items() if k in ("achieved", "frac", "traffic")})
print("cpu", d.get("cpu_baseline")); print("port16", str(d.get("torch_gpu_port_fp16_autocast"))[:200]); print("port32", str(d.get("torch_gpu_port"))[:200])
g = d["gpt"]; print("gpt", {k: g.get(k) for k in ("value", "ms_per_step", "e2e", "cpu_baseline", "error")}); print("gpt roofline", str(g.get("roofline"))[:400])
PY
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 --cpu-budget 40 > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_r2_final_reference.err; cat gpurun_out/bench_r2_final_reference.json | cut -c1-600
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_final.csv python tests/prof_s2.py > gpurun_out/ncu_s2.log 2>&1; tail -1 gpurun_out/ncu_s2.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gpt_launches_final.csv python tests/prof_gpt.py > gpurun_out/ncu_gpt.log 2>&1; tail -1 gpurun_out/ncu_gpt.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'gemm_tma|mel_fwd_warp|flash' -o gpurun_out/prof_final -f python tests/prof_kernels.py > gpurun_out/ncu_prof_final.log 2>&1; tail -2 gpurun_out/ncu_prof_final.log; ls -la gpurun_out/prof_final.ncu-rep
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/bench_r2_final_cfg5.json 2> gpurun_out/bench_r2_final_cfg5.err; cut -c1-700 gpurun_out/bench_r2_final_cfg5.json
timeout 300 python bench.py --sr-label 32000 --steps 20 --warmup 5 --gpt 0 --no-cpu-baseline --no-torch-port > gpurun_out/bench_r2_final_32k.json 2> gpurun_out/bench_r2_final_32k.err; cut -c1-400 gpurun_out/bench_r2_final_32k.json
