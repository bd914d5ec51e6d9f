#!/usr/bin/env bash
# round-2 first GPU call: everything round 1 left unrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_smi.txt 2>&1
nproc >> gpurun_out/r02_smi.txt
( B2S_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_zz_gpu_lcmap.py -q -x 2>&1 | tail -15 ) > gpurun_out/r02_lcmap.txt 2>&1
( B2S_FULLSIZE_LOG=24 timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k full_size 2>&1 | tail -15 ) > gpurun_out/r02_fullsize24.txt 2>&1
( timeout 600 python tools/spmv_probe.py 24 2>&1 | tail -5 ) > gpurun_out/r02_spmv_probe.txt 2>&1
( timeout 900 python bench.py --log-n 25 --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r02_bench_log25.json 2> gpurun_out/r02_bench_log25.err )
tail -3 gpurun_out/r02_lcmap.txt gpurun_out/r02_fullsize24.txt gpurun_out/r02_spmv_probe.txt
head -c 600 gpurun_out/r02_bench_log25.json; tail -3 gpurun_out/r02_bench_log25.err
