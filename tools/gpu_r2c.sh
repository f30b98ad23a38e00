#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -2 gpurun_out/r2c_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_n1.json.log 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c_bench_n1.err
python tools/show_bench.py gpurun_out/r2c_bench_n1.json.log 2>/dev/null | head -60
