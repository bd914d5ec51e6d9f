"""Quick device check of b2s_r1cs_upload_lcmap against b2s_r1cs_upload (SpMV equality on small circuits).
usage: python tools/lcmap_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
from oracle import r1cs as orc  # noqa: E402
from oracle.params import BLS12_381, BN254  # noqa: E402
from snark_b200 import Backend  # noqa: E402
from tests.test_host_lcmap import circuits  # noqa: E402
from tests.util import csr_from_rows, pack_fr, unpack_fr  # noqa: E402

ok = True
for cid, curve in enumerate((BLS12_381, BN254)):
    be = Backend(curve=cid)
    for name, cs in circuits(curve).items():
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        n_rows = len(mats[0])
        lm = cs.to_lcmap()
        args = [np.array(a, dtype=np.uint64) for a in lm["args"]]
        m_ref = be.r1cs_upload(n_rows, len(inst), len(wit), [csr_from_rows(curve, M) for M in mats])
        m_lc = be.r1cs_upload_lcmap(n_rows, len(inst), len(wit), args, np.array(lm["offsets"], dtype=np.uint64),
                                    np.array(lm["vars"] or [0], dtype=np.uint64), np.array(lm["coeffs"] or [0], dtype=np.uint32),
                                    pack_fr(curve, lm["pool"]))
        z = pack_fr(curve, inst + wit)
        ref, got = be.spmv(m_ref, z, n_rows), be.spmv(m_lc, z, n_rows)
        same = all(np.array_equal(ref[k], got[k]) for k in range(3))
        exp = all(unpack_fr(curve, got[k]) == orc.mat_vec_mul(curve.r, mats[k], inst + wit) for k in range(3))
        print(f"{curve.name:10s} {name:18s} rows={n_rows:3d} same_as_matrix_path={same} equals_oracle={exp}", flush=True)
        ok &= same and exp
    be.close()
print("LCMAP PROBE", "OK" if ok else "FAILED", f"{time.time() - t0:.1f}s")
sys.exit(0 if ok else 1)
