#!/usr/bin/env bash
mkdir -p gpurun_out
( B2S_PROFILE_VERBOSE=1 timeout 900 python bench.py --steps 1 --warmup 2 --no-extras --no-cpu --no-verify > gpurun_out/r02_bench_gap.json 2> gpurun_out/r02_bench_gap.err )
grep "b2s-profile" gpurun_out/r02_bench_gap.err | awk '{ if ($NF+0 > 0.15 || $(NF-1)=="gap" && $(NF-2)+0>0) print }' | awk '$(NF-1)+0 > 0.15' | head -80
python - <<'PY'
import re
tot=0; big=[]
for i,l in enumerate(open("gpurun_out/r02_bench_gap.err")):
    m=re.match(r"\[b2s-profile\] (\S+)\s+([\d.]+) ms\s+gap (-?[\d.]+) ms", l)
    if m:
        g=float(m.group(3))
        if g>0.03: tot+=g; big.append((g,i,m.group(1)))
print("sum of positive gaps > 0.03 ms:", round(tot,2))
for g,i,n in sorted(big, reverse=True)[:40]: print(f"{g:8.3f} ms before #{i} {n}")
PY
