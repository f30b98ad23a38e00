#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -12 gpurun_out/r2f_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json.log 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2f_bench_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/r2f_bench_n1.json.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
        print('parity',d['parity']['pass'], d['parity']['batched_fp64'])
        c5=d['legs']['c5']; print('c5', {k:c5.get(k) for k in ('snp_trait_tests_per_s','end_to_end_s','scan_int8_tops_per_gpu_rotation_once')}, c5.get('timings_s'))
PY
PERF_N=10000 PERF_M=151552 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"umma_i8|xx_contract|finalize|v_times|slice_v|vt_project" -c 40 --csv --log-file gpurun_out/r2f_launches_multitrait_n10k.csv python tools/perf_multitrait.py > gpurun_out/r2f_ncu.log 2>&1
