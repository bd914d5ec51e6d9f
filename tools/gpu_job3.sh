#!/usr/bin/env bash
mkdir -p gpurun_out
( B2S_MSM_AFFINE_ROUNDS=2 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_msm.py -x -q -k "small" 2>&1 | grep -v "^\s*$" | head -90 ) > gpurun_out/r02_msm_sanitizer.txt 2>&1
( PROBE_TOP=40 PROBE_CFGS="5:0" PROBE_PROFILE=1 timeout 900 python tools/msm_probe.py 24 1,2 2>&1 | tail -60 ) > gpurun_out/r02_msm_probe24b.txt 2>&1
( B2S_L2_GRAN=32 PROBE_TOP=12 PROBE_CFGS="0:0,5:0" PROBE_PROFILE=1 timeout 900 python tools/msm_probe.py 24 1 2>&1 | tail -60 ) > gpurun_out/r02_msm_probe24_gran32.txt 2>&1
( PROBE_TOP=40 PROBE_CFGS="4:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 21 1 2>&1 | tail -40 ) > gpurun_out/r02_msm_probe21b.txt 2>&1
( PROBE_CFGS="1:0" timeout 1200 ncu --set full --import-source on --clock-control none -k regex:msm_ba_p -c 8 -o gpurun_out/r02_ba_g1 -f python tools/msm_probe.py 24 1 2>&1 | tail -5 ) > gpurun_out/r02_ncu_g1.log 2>&1
( PROBE_CFGS="1:0" timeout 1200 ncu --set full --import-source on --clock-control none -k regex:msm_ba_p -c 8 -o gpurun_out/r02_ba_g2 -f python tools/msm_probe.py 23 2 2>&1 | tail -5 ) > gpurun_out/r02_ncu_g2.log 2>&1
head -50 gpurun_out/r02_msm_sanitizer.txt
cat gpurun_out/r02_msm_probe24b.txt gpurun_out/r02_msm_probe24_gran32.txt gpurun_out/r02_msm_probe21b.txt
ls -la gpurun_out/*.ncu-rep
