#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/final2_pytest.log 2>&1; tail -5 gpurun_out/final2_pytest.log
( time timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tests/run_gpu_checks.py --only flash_tc ) > gpurun_out/sanitizer_flash_tc.log 2>&1; grep "ERROR SUMMARY\|TOTAL\|real" gpurun_out/sanitizer_flash_tc.log
( time timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tests/run_gpu_checks.py --only infer_panel ) > gpurun_out/sanitizer_infer_panel.log 2>&1; grep "ERROR SUMMARY\|TOTAL\|real" gpurun_out/sanitizer_infer_panel.log
