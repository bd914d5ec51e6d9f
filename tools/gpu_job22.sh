#!/usr/bin/env bash
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r02_pytest_gpu.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ); tail -2 gpurun_out/r02_bench_n1.err
( timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu --no-verify > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err )
( timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu --no-verify > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err )
( timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify > gpurun_out/r02_ncu_bench.log 2>&1 )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r02_smoke.txt 2>&1
cat gpurun_out/r02_pytest_gpu.txt gpurun_out/r02_smoke.txt
python -c "
import json
for f in ('r02_bench_n1','r02_bench_n1_b','r02_bench_n1_c'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'), d.get('extras'))
d=json.load(open('gpurun_out/r02_bench_n1.json')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.8}); print(d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'])"
