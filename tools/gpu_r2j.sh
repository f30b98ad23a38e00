#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -6 gpurun_out/r2j_pytest.log
CMD="python bench.py --steps 2 --warmup 3 --no-legs --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r2j_bench_plain.json.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"umma_i8|finalize|vt_project|colstats|gram_|vanraden|slice_v|colsum|unpack|convert|xx_contract|v_times" -c 200 --csv --log-file gpurun_out/r2_launches_c2.csv $CMD > gpurun_out/r2j_ncu1.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench_n1.json.log 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_reference.json.log 2>&1; echo "ref rc=$?"
