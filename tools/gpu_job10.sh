#!/usr/bin/env bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py tests/test_gpu_serialize.py -x -q 2>&1 | tail -12 ) > gpurun_out/r02_t_dedup.txt 2>&1
( B2S_FULLSIZE_LOG=20,24 timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_limits.py -x -q 2>&1 | tail -8 ) >> gpurun_out/r02_t_dedup.txt 2>&1
( PROBE_TOP=12 PROBE_CFGS="auto:0" PROBE_PROFILE=1 timeout 600 python tools/msm_probe.py 24 1,2 2>&1 | tail -30 ) > gpurun_out/r02_probe24f.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err ); tail -3 gpurun_out/r02_bench_c.err
( B2S_MSM_DEDUP=0 timeout 900 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu > gpurun_out/r02_bench_c_nodedup.json 2> gpurun_out/r02_bench_c_nodedup.err ); tail -3 gpurun_out/r02_bench_c_nodedup.err
cat gpurun_out/r02_t_dedup.txt gpurun_out/r02_probe24f.txt
python - <<'PY'
import json
for f in ("r02_bench_c","r02_bench_c_nodedup"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.4})
    except Exception as e: print(f,"ERR",e)
PY
