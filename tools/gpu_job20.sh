#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err ); tail -2 gpurun_out/r02_bench_g.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_g.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'), d.get('host_affinity'))"
head -12 gpurun_out/r02_topo.txt | cut -c1-200
