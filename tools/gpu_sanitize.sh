#!/bin/bash
# compute-sanitizer over the small-shape parity tests of the kernels added in round 2 (memcheck, then racecheck on one)
mkdir -p gpurun_out
K="multi_trait or shared_rotation or unmerged or plink or numeric_text or blup or vanraden_gram_bit_exact or emmaxp3d_with_several or hand_case"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2_sanitizer_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "numeric_text or shared_rotation" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_racecheck.log | tail -8
