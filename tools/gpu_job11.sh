#!/usr/bin/env bash
# tests + bench + the ncu evidence of the round (launch list of bench, DRAM traffic of one MSM, full captures of the two passes)
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r02_pytest_gpu_b.txt 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err ); tail -3 gpurun_out/r02_bench_d.err
( timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_d_ref.json 2> gpurun_out/r02_bench_d_ref.err )
# launch list of the same bench command (cold-cache, serialised: compare SHARES)
( timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify > gpurun_out/r02_ncu_bench.log 2>&1 )
# DRAM traffic + duration of every accumulation kernel of ONE uniform 2^24-point G1 MSM (the h-query MSM of the proof)
( PROBE_KINDS=uniform PROBE_CFGS="auto:0" timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"msm_ba_|msm_accumulate" -c 24 --csv --log-file gpurun_out/r02_msm_traffic.csv python tools/msm_probe.py 24 1 > gpurun_out/r02_ncu_traffic.log 2>&1 )
# full captures: first-round pass 1 / pass 2, uniform scalars
( PROBE_KINDS=uniform PROBE_CFGS="auto:0" timeout 1200 ncu --set full --import-source on --clock-control none -k regex:msm_ba_p -c 2 -o gpurun_out/r02_final_ba_g1 -f python tools/msm_probe.py 24 1 > gpurun_out/r02_ncu_full.log 2>&1 )
ncu -i gpurun_out/r02_final_ba_g1.ncu-rep --page raw --csv > gpurun_out/r02_final_ba_g1_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_final_ba_g1.ncu-rep --page details > gpurun_out/r02_final_ba_g1_details.txt 2>/dev/null
cat gpurun_out/r02_pytest_gpu_b.txt; du -sh gpurun_out; ls -la gpurun_out | head -30
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_d.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('verified')); print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.4}); print(d['roofline']); print(open('gpurun_out/r02_bench_d_ref.json').read()[:600])"
