#!/usr/bin/env bash
mkdir -p gpurun_out
( PROBE_TOP=6 PROBE_CFGS="0:0,5:0" timeout 900 python tools/msm_probe.py 24 1,2 2>&1 | tail -30 ) > gpurun_out/r02_probe24d.txt 2>&1
( PROBE_TOP=6 PROBE_CFGS="0:0,5:0" timeout 900 python tools/msm_probe.py 22 1 2>&1 | tail -30 ) >> gpurun_out/r02_probe24d.txt 2>&1
( B2S_FULLSIZE_LOG=24 timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -k full_size 2>&1 | tail -15 ) > gpurun_out/r02_t_full24.txt 2>&1
( timeout 1200 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02_t_dist.txt 2>&1
cat gpurun_out/r02_probe24d.txt gpurun_out/r02_t_full24.txt gpurun_out/r02_t_dist.txt
