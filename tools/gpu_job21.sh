#!/usr/bin/env bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q 2>&1 | tail -4 ) > gpurun_out/r02_t_tiny.txt 2>&1
( B2S_FULLSIZE_LOG=20,24 timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4 ) >> gpurun_out/r02_t_tiny.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err ); tail -2 gpurun_out/r02_bench_h.err
cat gpurun_out/r02_t_tiny.txt
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_h.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.8})"
