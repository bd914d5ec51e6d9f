"""How the C++ oracle's MSM scales with host threads on this box (sanity check of the CPU baseline)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import cnative

n = 1 << 18
g1 = np.array(bench._G1_GEN_MONT, dtype=np.uint32)
print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "hw", cnative.threads_default())
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cpu.max n/a", e)
t0 = time.perf_counter(); bases = cnative.multiples(0, 1, g1, 11, n, cnative.threads_default()); print("multiples", time.perf_counter() - t0)
rng = np.random.default_rng(1)
x = np.frombuffer(rng.bytes(32 * n), dtype=np.uint32).copy(); x[7::8] &= 0x3fffffff
for t in (1, 4, 16, 32, 64, 128):
    t0 = time.perf_counter(); cnative.msm(0, 1, bases, x, n, True, t); print("threads", t, "msm 2^18:", round(time.perf_counter() - t0, 3), "s")
