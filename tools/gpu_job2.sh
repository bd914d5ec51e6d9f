#!/usr/bin/env bash
mkdir -p gpurun_out
( B2S_MSM_AFFINE_ROUNDS=3 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02_msm_forced.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02_msm_default.txt 2>&1
( B2S_MSM_AFFINE_ROUNDS=2 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_msm.py -x -q -k "small or edge" 2>&1 | tail -25 ) > gpurun_out/r02_msm_sanitizer.txt 2>&1
( PROBE_CFGS="0:0,3:0,5:0,6:0" PROBE_PROFILE=1 timeout 900 python tools/msm_probe.py 24 1,2 2>&1 | tail -60 ) > gpurun_out/r02_msm_probe24.txt 2>&1
( PROBE_CFGS="0:0,4:0,5:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 21 1,2 2>&1 | tail -40 ) > gpurun_out/r02_msm_probe21.txt 2>&1
tail -4 gpurun_out/r02_msm_forced.txt gpurun_out/r02_msm_default.txt gpurun_out/r02_msm_sanitizer.txt
cat gpurun_out/r02_msm_probe24.txt
