#!/bin/bash
# final pass of the round: whole GPU suite, smoke, default bench line, the inference / preparation tools
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/final_pytest.log 2>&1; tail -5 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log | cut -c1-200
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b_final.json ) 2> gpurun_out/bench_r2b_final.err; tail -4 gpurun_out/bench_r2b_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2b_final.json"))
print("s2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
r = d["roofline"]; print("roofline", r["achieved"], r["frac"], r["traffic"], r.get("share_of_step_time"))
g = d["gpt"]; print("gpt", g.get("value"), g.get("ms_per_step"), g["e2e"]["value"], g.get("error")); print("gpt roofline", g["roofline"]["achieved"], g["roofline"]["frac"], g["roofline"].get("share_of_step_time"), g["roofline"].get("attention_ms"))
PY
timeout 200 python tools/bench_infer_panel.py > gpurun_out/bench_infer_panel_final.json 2> gpurun_out/bench_infer_panel.err; cut -c1-330 gpurun_out/bench_infer_panel_final.json; tail -2 gpurun_out/bench_infer_panel.err
timeout 200 python tools/bench_hubert.py > gpurun_out/bench_hubert_final.json 2> gpurun_out/bench_hubert.err; cut -c1-300 gpurun_out/bench_hubert_final.json; tail -2 gpurun_out/bench_hubert.err
