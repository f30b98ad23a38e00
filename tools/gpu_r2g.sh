#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "plink or numeric_text or one_process or sharded_pipeline" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -8 gpurun_out/r2g_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_n2.json.log 2> gpurun_out/r2g_bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2g_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2g_bench_n2.json.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
        print('kinship', d['kinship'])
        for k,v in d['legs'].items():
            print(k, json.dumps(v)[:1500])
PY
