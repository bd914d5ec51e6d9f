#!/usr/bin/env bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_groth16.py tests/test_gpu_setup.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r02_t_side.txt 2>&1
( B2S_FULLSIZE_LOG=20,24 timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6 ) >> gpurun_out/r02_t_side.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err ); tail -3 gpurun_out/r02_bench_e.err
( B2S_NO_SIDE_STREAM=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu --no-verify > gpurun_out/r02_bench_e_noside.json 2> gpurun_out/r02_bench_e_noside.err ); tail -3 gpurun_out/r02_bench_e_noside.err
cat gpurun_out/r02_t_side.txt
python - <<'PY'
import json
for f in ("r02_bench_e","r02_bench_e_noside"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'))
    except Exception as e: print(f,"ERR",e)
PY
