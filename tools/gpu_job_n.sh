#!/usr/bin/env bash
N=$1
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err )
tail -3 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_n$N.json")); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'), d.get('host_affinity')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.3})
PY
