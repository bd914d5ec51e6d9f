#!/usr/bin/env bash
# round 2, session 3, call A: the 6-transform witness map + full-size NTT tables (and the composed fallback), then bench lines
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_groth16.py tests/test_gpu_dist.py -x -q 2>&1 | tail -6 ) > gpurun_out/r03_t_a.txt 2>&1
( B2S_NTT_FULL=0 timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_groth16.py -x -q 2>&1 | tail -4 ) > gpurun_out/r03_t_a_nofull.txt 2>&1
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r03_bench_a.json 2> gpurun_out/r03_bench_a.err ); tail -2 gpurun_out/r03_bench_a.err
( B2S_NTT_FULL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-verify --no-extras > gpurun_out/r03_bench_a_nofull.json 2> gpurun_out/r03_bench_a_nofull.err )
cat gpurun_out/r03_t_a.txt gpurun_out/r03_t_a_nofull.txt
python - <<'PY'
import json
for f in ('r03_bench_a', 'r03_bench_a_nofull'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'), (d.get('extras') or {}).get('ntt_2p24_forward'))
        print({k: v for k, v in d['kernel_ms_per_step'].items() if v > 1.5})
    except Exception as e:
        print(f, 'failed', e)
PY
