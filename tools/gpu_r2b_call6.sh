#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/flash_tc_ab.py > gpurun_out/flash_tc_ab3.json 2> gpurun_out/flash_tc_ab3.err; echo "flash_tc_ab rc=$?"; tail -12 gpurun_out/flash_tc_ab3.err | cut -c1-400
for tc in 0 2; do
EVK_FLASH_TC=$tc timeout 300 python tests/run_gpu_checks.py --only gpt_kernels 2>&1 | grep "FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File\|drop" | head -30
done
EVK_FLASH_TC=0 timeout 300 python tests/run_gpu_checks.py --only gpt_small 2>&1 | grep "FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File" | head
