"""CPU time of the transforms and commitments of ONE universal-setup (Marlin-style) proof of tools/marlin_bench.py's shape,
with the C++ oracle (oracle/c/oracle.cpp: the restated ark-poly radix-2 FFT and ark-ec Pippenger, all host threads).

Not a CPU prover: it issues the list of `ntt` / `commit` calls `snark_b200/marlin.py::prove_assigned` makes for |H| = |K| =
2^log_n (sizes below are read off that function) on random data and leaves out the element-wise work -- a LOWER bound for a
CPU prover built on these algorithms, reported as context next to the GPU line (a baseline, not a target).
usage: python tools/marlin_cpu_shape.py [log_n] [curve: bn254 | bls12_381]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import cnative
from tests.util import pack_points

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cname = sys.argv[2] if len(sys.argv) > 2 else "bn254"
from oracle.params import BLS12_381, BN254

curve = BN254 if cname == "bn254" else BLS12_381
cid = curve.curve_id
n = 1 << log_n
D = 3 * n
thr = cnative.threads_default()
rng = np.random.default_rng(5)


def rand_fr(k):
    a = rng.integers(0, 1 << 32, size=(k, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x0FFFFFFF            # < r for both curves (values are arbitrary Montgomery residues)
    return np.ascontiguousarray(a.reshape(-1))


t0 = time.perf_counter()
gen = pack_points(curve, 1, [curve.g1])
srs = cnative.multiples(cid, 1, gen, 1, D + 1)        # (i + 1) G: arbitrary distinct bases, same MSM cost as tau^i G
t_bases = time.perf_counter() - t0

# transforms of one proof: (log2 size, count)   -- prove_assigned: rounds 1-3 and the two opening quotients
ntts = [(log_n, 7), (log_n + 2, 13)]
# commitments: sizes in units of 2^log_n  -- w, zA, zB, t, g1, g1 shifted, h1 (2), g2, g2 shifted, h2 (3), W1 (3), W2 (3)
msms = [1, 1, 1, 1, 1, 1, 2, 1, 1, 3, 3, 3]
t_ntt = 0.0
for lg, cnt in ntts:
    x = rand_fr(1 << lg)
    for _ in range(cnt):
        t = time.perf_counter()
        cnative.ntt(cid, x, lg)
        t_ntt += time.perf_counter() - t
t_msm = 0.0
for k in msms:
    sc = rand_fr(k * n)
    t = time.perf_counter()
    cnative.msm(cid, 1, srs, sc, k * n, mont=True)
    t_msm += time.perf_counter() - t
print(json.dumps({"what": "CPU time of the transforms + commitments of one universal-setup proof (C++ oracle, lower bound for a CPU prover)",
                  "curve": cname, "log_n": log_n, "threads": thr, "s_ntt": round(t_ntt, 3), "s_msm": round(t_msm, 3),
                  "s_total": round(t_ntt + t_msm, 3), "ntt_calls": sum(c for _, c in ntts), "msm_points": sum(msms) * n,
                  "s_base_generation": round(t_bases, 2), "host": "build container (not the GPU box)"}))
