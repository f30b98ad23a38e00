#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2k_bench_n8.json.log 2> gpurun_out/r2k_bench_n8.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2k_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r2k_bench_n8.json.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
        print('kinship', d['kinship'])
        for k,v in d['legs'].items():
            print(k, json.dumps(v)[:1600])
PY
python tools/perf_eigh_mg.py 20000 8 30000 > gpurun_out/r2k_eigh_mg_n8.log 2>&1; grep -v Warn gpurun_out/r2k_eigh_mg_n8.log | tail -4
