#!/bin/bash
# flash_tc A/B (parity + timing), then the check groups touched since call 1, then the decode / HuBERT tools
mkdir -p gpurun_out
timeout 240 python tools/flash_tc_ab.py > gpurun_out/flash_tc_ab.json 2> gpurun_out/flash_tc_ab.err; echo "flash_tc_ab rc=$?"; tail -12 gpurun_out/flash_tc_ab.err
for grp in gpt_kernels gpt_small gpt_ragged hubert infer_panel; do
  EVK_FLASH_TC=1 timeout 300 python tests/run_gpu_checks.py --only $grp 2>&1 | grep "FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File\|gemv" | head -30
done
timeout 200 python tools/bench_hubert.py > gpurun_out/bench_hubert.json 2> gpurun_out/bench_hubert.err; cat gpurun_out/bench_hubert.json; tail -3 gpurun_out/bench_hubert.err
timeout 200 python tools/bench_infer_panel.py > gpurun_out/bench_infer_panel_gemv.json 2> gpurun_out/bench_infer_panel.err; cat gpurun_out/bench_infer_panel_gemv.json; tail -3 gpurun_out/bench_infer_panel.err
