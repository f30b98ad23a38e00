#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vanraden or large_taxa or gram or sharded or one_process or zhang or c2_properties" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -5 gpurun_out/r2m_pytest.log
PERF_N=50000 PERF_M=32768 python tools/perf_kinship.py > gpurun_out/r2m_kin50k.log 2>&1; tail -5 gpurun_out/r2m_kin50k.log
