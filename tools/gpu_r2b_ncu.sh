#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash -f -o gpurun_out/prof_flash python tests/prof_flash.py > gpurun_out/ncu_prof_flash.log 2>&1; tail -2 gpurun_out/ncu_prof_flash.log; ls -la gpurun_out/prof_flash.ncu-rep
ncu -i gpurun_out/prof_flash.ncu-rep --page raw --csv > gpurun_out/prof_flash_raw.csv 2>/dev/null; wc -c gpurun_out/prof_flash_raw.csv
