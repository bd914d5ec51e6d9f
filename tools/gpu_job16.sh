#!/usr/bin/env bash
# tests + bench + the ncu evidence of the round (launch list of bench, DRAM traffic of one MSM, full captures of the round kernels)
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r02_pytest_gpu.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ); tail -3 gpurun_out/r02_bench_n1.err
( B2S_MSM_DEDUP=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r02_bench_n1_nodedup.json 2> gpurun_out/r02_bench_n1_nodedup.err )
( timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err )
( timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify > gpurun_out/r02_ncu_bench.log 2>&1 )
( PROBE_KINDS=uniform PROBE_CFGS="auto:0" timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"msm_ba_|msm_accumulate" -c 46 --csv --log-file gpurun_out/r02_msm_traffic.csv python tools/msm_probe.py 24 1 > gpurun_out/r02_ncu_traffic.log 2>&1 )
( PROBE_KINDS=uniform PROBE_CFGS="auto:0" timeout 1200 ncu --set full --clock-control none -k regex:msm_ba_p -c 4 -o gpurun_out/r02_final_ba_g1 -f python tools/msm_probe.py 24 1 > gpurun_out/r02_ncu_full.log 2>&1 )
ncu -i gpurun_out/r02_final_ba_g1.ncu-rep --page raw --csv > gpurun_out/r02_final_ba_g1_raw.csv 2>/dev/null
rm -f gpurun_out/r02_final_ba_g1.ncu-rep
cat gpurun_out/r02_pytest_gpu.txt; du -sh gpurun_out
python -c "
import json
for f in ('r02_bench_n1','r02_bench_n1_nodedup'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified'), d.get('extras'))
d=json.load(open('gpurun_out/r02_bench_n1.json')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.4}); print(d['roofline']); print(d['cpu_baseline'])"
