#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/flash_tc_ab.py > gpurun_out/flash_tc_ab2.json 2> gpurun_out/flash_tc_ab2.err; echo "flash_tc_ab rc=$?"; tail -14 gpurun_out/flash_tc_ab2.err
EVK_FLASH_TC=2 timeout 300 python tests/run_gpu_checks.py --only gpt_kernels 2>&1 | grep "FAIL\|TOTAL\|EXC\|Error\|error\|Traceback\|File" | head -30
