#!/usr/bin/env bash
mkdir -p gpurun_out
( B2S_MSM_AFFINE_ROUNDS=3 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q 2>&1 | tail -5 ) > gpurun_out/r02_t_msm_forced.txt 2>&1
( B2S_FULLSIZE_LOG=20 timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -8 ) > gpurun_out/r02_t_default.txt 2>&1
( PROBE_TOP=8 PROBE_CFGS="5:0" PROBE_PROFILE=1 timeout 900 python tools/msm_probe.py 24 1,2 2>&1 | tail -30 ) > gpurun_out/r02_probe24c.txt 2>&1
( PROBE_TOP=8 PROBE_CFGS="4:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 21 1,2 2>&1 | tail -30 ) > gpurun_out/r02_probe21c.txt 2>&1
( timeout 900 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err ); tail -3 gpurun_out/r02_bench_a.err
cat gpurun_out/r02_t_msm_forced.txt gpurun_out/r02_t_default.txt gpurun_out/r02_probe24c.txt gpurun_out/r02_probe21c.txt
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_a.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified')); print(d['kernel_ms_per_step'])"
