#!/usr/bin/env bash
mkdir -p gpurun_out
( B2S_MSM_AFFINE_ROUNDS=2 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_msm.py -x -q -k "small or edge" 2>&1 | grep -v "Host Frame" | grep -v "^\s*$" | head -60 ) > gpurun_out/r02_msm_sanitizer.txt 2>&1
NCU="ncu --set full --clock-control none -k regex:msm_ba_p -c 2"
( PROBE_KINDS=uniform PROBE_CFGS="1:0" timeout 900 $NCU -o gpurun_out/r02_ba_g1_uniform -f python tools/msm_probe.py 24 1 2>&1 | tail -3 ) > gpurun_out/r02_ncu.log 2>&1
( PROBE_KINDS=equal PROBE_CFGS="1:0" timeout 900 $NCU -o gpurun_out/r02_ba_g1_equal -f python tools/msm_probe.py 24 1 2>&1 | tail -3 ) >> gpurun_out/r02_ncu.log 2>&1
( PROBE_KINDS=equal PROBE_CFGS="1:0" timeout 900 $NCU -o gpurun_out/r02_ba_g2_equal -f python tools/msm_probe.py 23 2 2>&1 | tail -3 ) >> gpurun_out/r02_ncu.log 2>&1
for f in g1_uniform g1_equal g2_equal; do ncu -i gpurun_out/r02_ba_$f.ncu-rep --page raw --csv > gpurun_out/r02_ba_${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/
du -sh gpurun_out
head -40 gpurun_out/r02_msm_sanitizer.txt
