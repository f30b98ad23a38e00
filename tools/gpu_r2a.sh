#!/bin/bash
# round-2 first GPU pass: parity suite, multi-trait A/B (shared rotation vs trait chunks), launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python tools/perf_multitrait.py > gpurun_out/r2a_multitrait_shared.log 2>&1; tail -6 gpurun_out/r2a_multitrait_shared.log
GAPIT_MULTI_CHUNKED=1 python tools/perf_multitrait.py > gpurun_out/r2a_multitrait_chunked.log 2>&1; tail -6 gpurun_out/r2a_multitrait_chunked.log
PERF_T=8 python tools/perf_multitrait.py > gpurun_out/r2a_multitrait_shared_T8.log 2>&1; tail -5 gpurun_out/r2a_multitrait_shared_T8.log
PERF_T=8 GAPIT_MULTI_CHUNKED=1 python tools/perf_multitrait.py > gpurun_out/r2a_multitrait_chunked_T8.log 2>&1; tail -5 gpurun_out/r2a_multitrait_chunked_T8.log
PERF_N=10000 PERF_M=151552 python tools/perf_multitrait.py > gpurun_out/r2a_multitrait_shared_n10k.log 2>&1; tail -6 gpurun_out/r2a_multitrait_shared_n10k.log
PERF_M=37888 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches_multitrait.csv python tools/perf_multitrait.py > gpurun_out/r2a_ncu.log 2>&1
