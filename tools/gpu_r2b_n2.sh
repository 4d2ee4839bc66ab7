#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_ranks" 2>&1 | tail -3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_r2b_n2.json 2> gpurun_out/bench_r2b_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2b_n2.err | cut -c1-300
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2b_n2.json"))
print("s2", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]); g = d.get("gpt"); print("gpt", g and (g.get("value"), g.get("ms_per_step"), g.get("error")))
PY
