#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -12 gpurun_out/r2d_pytest.log
PERF_N=10000 PERF_M=151552 python tools/perf_multitrait.py > gpurun_out/r2d_multitrait_n10k.log 2>&1; tail -4 gpurun_out/r2d_multitrait_n10k.log
python tools/run_config.py --config c5 > gpurun_out/r2d_cfg_c5.json.log 2>&1; tail -c 1500 gpurun_out/r2d_cfg_c5.json.log
PERF_N=10000 PERF_M=151552 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"umma_i8|xx_contract|finalize|v_times|slice_v|vt_project" -c 40 --csv --log-file gpurun_out/r2d_launches_multitrait_n10k.csv python tools/perf_multitrait.py > gpurun_out/r2d_ncu.log 2>&1
