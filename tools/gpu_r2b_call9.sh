#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/flash_tc_ab.py > gpurun_out/flash_tc_ab4.json 2> gpurun_out/flash_tc_ab4.err; echo "flash_tc_ab rc=$?"; tail -11 gpurun_out/flash_tc_ab4.err | cut -c1-330
