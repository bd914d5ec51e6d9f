#!/usr/bin/env bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q 2>&1 | tail -6 ) > gpurun_out/r02_t_stage.txt 2>&1
( B2S_FULLSIZE_LOG=20,24 timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6 ) >> gpurun_out/r02_t_stage.txt 2>&1
( B2S_MSM_STAGE=1 PROBE_KINDS=uniform PROBE_TOP=8 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 24 1,2 2>&1 | tail -12 ) > gpurun_out/r02_probe_stage1.txt 2>&1
( B2S_MSM_STAGE=0 PROBE_KINDS=uniform PROBE_TOP=8 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 24 1,2 2>&1 | tail -12 ) > gpurun_out/r02_probe_stage0.txt 2>&1
( B2S_MSM_STAGE=1 PROBE_KINDS=uniform PROBE_TOP=8 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 22 1 2>&1 | tail -12 ) >> gpurun_out/r02_probe_stage1.txt 2>&1
( B2S_MSM_STAGE=0 PROBE_KINDS=uniform PROBE_TOP=8 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 22 1 2>&1 | tail -12 ) >> gpurun_out/r02_probe_stage0.txt 2>&1
cat gpurun_out/r02_t_stage.txt; echo STAGE1; cat gpurun_out/r02_probe_stage1.txt; echo STAGE0; cat gpurun_out/r02_probe_stage0.txt
