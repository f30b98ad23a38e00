#!/bin/bash
# round-2 second GPU pass (2 GPUs): full parity suite after the single-form refactor, multi-GPU tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
python tools/perf_multitrait.py > gpurun_out/r2b_multitrait.log 2>&1; tail -4 gpurun_out/r2b_multitrait.log
PERF_N=10000 PERF_M=151552 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"umma_i8|xx_contract|finalize|v_times|slice_v|vt_project" -c 60 --csv --log-file gpurun_out/r2b_launches_multitrait_n10k.csv python tools/perf_multitrait.py > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
