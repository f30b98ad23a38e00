#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "gtindex or by_file or device_resident or plink" > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
SECONDS=0
python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench_n1.json.log 2> gpurun_out/r2r_bench_n1.err; echo "bench rc=$? in ${SECONDS}s"; tail -c 300 gpurun_out/r2r_bench_n1.err
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2r_bench_n2.json.log 2> gpurun_out/r2r_bench_n2.err; echo "bench2 rc=$? in ${SECONDS}s"; tail -c 300 gpurun_out/r2r_bench_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2r_bench_n1.json.log','gpurun_out/r2r_bench_n2.json.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kin',d['kinship']['tops'])
            for k,v in d['legs'].items():
                print(' ',k, json.dumps(v)[:1400])
PY
