#!/usr/bin/env bash
mkdir -p gpurun_out
( B2S_HOST_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 2 --no-extras --no-cpu --no-verify > gpurun_out/r02_bench_host.json 2> gpurun_out/r02_bench_host.err )
grep "b2s-host" gpurun_out/r02_bench_host.err | tail -24
