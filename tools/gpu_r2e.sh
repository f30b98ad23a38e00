#!/bin/bash
mkdir -p gpurun_out
export PERF_N=10000 PERF_M=75776
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"finalize|xx_contract|StoreEpi" -c 3 -o gpurun_out/r2e_multitrait_full python tools/perf_multitrait.py > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
ncu -i gpurun_out/r2e_multitrait_full.ncu-rep --page raw --csv > gpurun_out/r2e_multitrait_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2e_*
