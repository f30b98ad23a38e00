#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -6 gpurun_out/r2l_pytest.log
PERF_N=50000 PERF_M=32768 python tools/perf_kinship.py > gpurun_out/r2l_kin50k.log 2>&1; tail -6 gpurun_out/r2l_kin50k.log
python bench.py --steps 5 --warmup 3 --no-legs --no-e2e --no-cpu-baseline > gpurun_out/r2l_bench_quick.json.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r2l_bench_quick.json.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'kin',d['kinship']['tops'],d['kinship']['ms'],d['kinship']['gram_kernel_ms'])
PY
