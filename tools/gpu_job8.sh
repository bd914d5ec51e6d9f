#!/usr/bin/env bash
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r02_pytest_gpu_a.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err ); tail -3 gpurun_out/r02_bench_b.err
( PROBE_TOP=10 PROBE_CFGS="9:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 21 1,2 2>&1 | tail -30 ) > gpurun_out/r02_probe21e.txt 2>&1
cat gpurun_out/r02_pytest_gpu_a.txt gpurun_out/r02_probe21e.txt
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_b.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'), d.get('extras')); print(d['kernel_ms_per_step']); print(d['roofline']); print(d.get('cpu_baseline'))"
