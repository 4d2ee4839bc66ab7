#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/final3_pytest.log 2>&1; tail -5 gpurun_out/final3_pytest.log
timeout 200 python tools/flash_tc_ab.py > gpurun_out/flash_tc_ab5.json 2> gpurun_out/flash_tc_ab5.err; echo "ab rc=$?"; tail -6 gpurun_out/flash_tc_ab5.err | cut -c1-200
