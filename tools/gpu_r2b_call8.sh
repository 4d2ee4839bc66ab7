#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/run_gpu_checks.py --only infer_panel 2>&1 | grep "FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File\|attn_decode" | head -20
timeout 200 python tools/exp/infer_breakdown.py 2>&1 | tail -6 | tee gpurun_out/infer_breakdown2.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/infer_token_launches.csv python tools/exp/infer_breakdown.py --ncu > gpurun_out/infer_ncu.log 2>&1; tail -2 gpurun_out/infer_ncu.log; python tools/sum_launches.py gpurun_out/infer_token_launches.csv 14 | tee gpurun_out/infer_token_launches.txt
