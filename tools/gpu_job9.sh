#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n8_smi.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err )
tail -4 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n8.json")); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.3})
PY
