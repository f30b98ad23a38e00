#!/bin/bash
mkdir -p gpurun_out
python tools/perf_eigh_mg.py 4096 2 > gpurun_out/r2h_eigh_mg.log 2>&1
python tools/perf_eigh_mg.py 12000 2 >> gpurun_out/r2h_eigh_mg.log 2>&1
python tools/perf_eigh_mg.py 20000 2 30000 >> gpurun_out/r2h_eigh_mg.log 2>&1
grep -v Warning gpurun_out/r2h_eigh_mg.log | tail -20
