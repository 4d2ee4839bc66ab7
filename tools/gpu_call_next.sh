#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tests/run_gpu_checks.py --only hubert 2>&1 | grep "^ok\|FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File" | head -20
timeout 300 python tools/bench_hubert.py > gpurun_out/bench_hubert.json 2> gpurun_out/bench_hubert.err; cat gpurun_out/bench_hubert.json; tail -3 gpurun_out/bench_hubert.err
for gr in 1 0; do
  echo "== EVK_INFER_GRAPH=$gr"
  EVK_INFER_GRAPH=$gr timeout 500 python tests/run_gpu_checks.py --only infer_panel 2>&1 | grep "^ok\|FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File" | head -20
  EVK_INFER_GRAPH=$gr timeout 300 python tools/bench_infer_panel.py > gpurun_out/bench_infer_panel_g$gr.json 2> gpurun_out/bench_infer_panel.err; cat gpurun_out/bench_infer_panel_g$gr.json; tail -3 gpurun_out/bench_infer_panel.err
done
