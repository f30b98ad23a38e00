#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -4 gpurun_out/r2n_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-legs --no-e2e > gpurun_out/r2n_bench_n2.json.log 2> gpurun_out/r2n_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2n_bench_n2.json.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'kin',d['kinship']['tops'],d['kinship']['ms'],d['kinship']['gram_kernel_ms'], d['kinship']['exchange_matches_nccl_bit_for_bit'])
PY
