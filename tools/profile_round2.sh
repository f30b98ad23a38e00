#!/bin/bash
# Round-2 profile evidence on one GPU (run through gpurun; each capture only after the plain command exited 0).
# Outputs land in gpurun_out/; tools/ncu_traffic.py and the summaries under profiles/ are made from the raw CSVs.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-legs --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r2i_bench_plain.json.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/r2i_bench_plain.json.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(?!.*(sytrd|syherk|transpose|stedc|steqr|lansy|copy_info|xx_set|laed|larf|ormtr|gemm|cutlass|elementwise|reduce_kernel)).*$" -c 200 --csv --log-file gpurun_out/r2_launches_c2.csv $CMD > gpurun_out/r2i_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_i8 -s 3 -c 9 -o gpurun_out/r2_ncu_full_c2 $CMD > gpurun_out/r2i_ncu2.log 2>&1
ncu -i gpurun_out/r2_ncu_full_c2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_c2_raw.csv 2>/dev/null
PERF_N=50000 PERF_M=32768 python tools/perf_kinship.py > gpurun_out/r2i_kin50k_plain.log 2>&1 && \
PERF_N=50000 PERF_M=32768 ncu --set full --clock-control none -k regex:umma_i8 -s 1 -c 1 -o gpurun_out/r2_ncu_full_gram_n50000 python tools/perf_kinship.py > gpurun_out/r2i_ncu3.log 2>&1
ncu -i gpurun_out/r2_ncu_full_gram_n50000.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_gram_n50000_raw.csv 2>/dev/null
python bench.py --steps 5 --warmup 3 --slices 6 --no-legs --no-e2e > gpurun_out/r2i_bench_slices6.json.log 2>&1
ls -la gpurun_out/r2_* gpurun_out/r2i_*
