#!/usr/bin/env bash
# round 2, final evidence: whole GPU suite, default bench, smoke, ncu capture of the NTT passes with streamed twiddles
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r02_pytest_gpu_final.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err ); tail -2 gpurun_out/r02_bench_final_n1.err
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r02_smoke_final.txt 2>&1
( timeout 600 ncu --set full --clock-control none -k regex:ntt_pass --launch-skip 3 -c 3 -o gpurun_out/r02_ntt_full -f python tools/ntt_probe.py 24 > gpurun_out/r02_ncu_ntt.log 2>&1 )
ncu -i gpurun_out/r02_ntt_full.ncu-rep --page raw --csv > gpurun_out/r02_ntt_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ntt_full.ncu-rep
( timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify > gpurun_out/r02_ncu_bench.log 2>&1 )
cat gpurun_out/r02_pytest_gpu_final.txt gpurun_out/r02_smoke_final.txt; tail -3 gpurun_out/r02_ncu_ntt.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_final_n1.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'), d.get('extras'))
print({k: v for k, v in d['kernel_ms_per_step'].items() if v > 1.5}); print(d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'])
PY
