#!/usr/bin/env bash
mkdir -p gpurun_out
nproc > gpurun_out/r02_box.txt; lscpu | grep "Model name" >> gpurun_out/r02_box.txt
for i in 1 2; do
( timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu --no-verify > gpurun_out/r02_bench_f$i.json 2> gpurun_out/r02_bench_f$i.err ); tail -2 gpurun_out/r02_bench_f$i.err
done
( B2S_PROFILE_VERBOSE=1 timeout 900 python bench.py --steps 1 --warmup 2 --no-extras --no-cpu --no-verify > gpurun_out/r02_bench_f3.json 2> gpurun_out/r02_bench_f3.err )
cat gpurun_out/r02_box.txt
python - <<'PY'
import json
for f in ("r02_bench_f1","r02_bench_f2","r02_bench_f3"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['wall_ms_per_step'])
    except Exception as e: print(f,"ERR",e)
PY
grep -c "b2s-profile" gpurun_out/r02_bench_f3.err
