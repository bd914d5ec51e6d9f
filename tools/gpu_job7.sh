#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_smi.txt
( timeout 600 python -m pytest tests/test_gpu_groth16.py -x -q -k "group" 2>&1 | tail -8 ) > gpurun_out/r02_t_group2.txt 2>&1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err )
( B2S_DIST_WITNESS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-verify > gpurun_out/r02_bench_n2_repl.json 2> gpurun_out/r02_bench_n2_repl.err )
cat gpurun_out/r02_t_group2.txt; tail -5 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
for f in ("r02_bench_n2","r02_bench_n2_repl"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.5})
    except Exception as e: print(f, "ERR", e)
PY
