#!/usr/bin/env bash
mkdir -p gpurun_out
( B2S_MSM_AFFINE_ROUNDS=3 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q 2>&1 | tail -4 ) > gpurun_out/r02_t_bulk.txt 2>&1
( timeout 1500 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q 2>&1 | tail -4 ) >> gpurun_out/r02_t_bulk.txt 2>&1
( B2S_FULLSIZE_LOG=20,24 timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4 ) >> gpurun_out/r02_t_bulk.txt 2>&1
( B2S_MSM_AFFINE_ROUNDS=2 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_msm.py -x -q -k "small or edge or repeated" 2>&1 | grep -v "Host Frame" | grep -v "^\s*$" | tail -12 ) > gpurun_out/r02_sanitizer_bulk.txt 2>&1
( PROBE_TOP=6 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 24 1,2 2>&1 | tail -12 ) > gpurun_out/r02_probe_bulk1.txt 2>&1
( B2S_MSM_BULK=0 B2S_MSM_DEDUP=0 PROBE_TOP=6 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 24 1,2 2>&1 | tail -12 ) > gpurun_out/r02_probe_bulk0.txt 2>&1
( B2S_MSM_DEDUP=0 PROBE_TOP=6 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 24 1,2 2>&1 | tail -12 ) > gpurun_out/r02_probe_bulk1_nodedup.txt 2>&1
cat gpurun_out/r02_t_bulk.txt gpurun_out/r02_sanitizer_bulk.txt; echo BULK1; cat gpurun_out/r02_probe_bulk1.txt; echo BULK0-nodedup; cat gpurun_out/r02_probe_bulk0.txt; echo BULK1-nodedup; cat gpurun_out/r02_probe_bulk1_nodedup.txt
