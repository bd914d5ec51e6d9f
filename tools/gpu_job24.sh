#!/usr/bin/env bash
# round 2, session 3, call B: universal-setup (Marlin-style) path -- polynomial kernels, GPU prover vs big-int prover, 2^20 bench line
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_marlin.py -x -q 2>&1 | tail -25 ) > gpurun_out/r03_t_marlin.txt 2>&1
cat gpurun_out/r03_t_marlin.txt
( timeout 600 python tools/marlin_bench.py --log-n 20 --steps 3 > gpurun_out/r03_marlin_bench.json 2> gpurun_out/r03_marlin_bench.err ); tail -5 gpurun_out/r03_marlin_bench.err | cut -c1-400
cat gpurun_out/r03_marlin_bench.json | cut -c1-1500
