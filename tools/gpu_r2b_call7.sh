#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/exp/infer_breakdown.py 2>&1 | tail -6 | tee gpurun_out/infer_breakdown.txt
timeout 300 python tests/run_gpu_checks.py --only flash_tc 2>&1 | grep "FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File" | head -20
( time timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tests/run_gpu_checks.py --only flash_tc ) > gpurun_out/sanitizer_flash_tc.log 2>&1; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/sanitizer_flash_tc.log; grep "ERROR SUMMARY\|TOTAL\|real" gpurun_out/sanitizer_flash_tc.log
