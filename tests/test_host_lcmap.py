"""CSR built straight from the constraint system's LcMap.  The per-row functions the GPU kernels call
(snark_b200/csrc/lcmap.cuh) are compiled for the host (tests/native/host_ff.cpp: ht_lcmap_csr) and must reproduce
`to_matrices()` (constraint_system.rs:768-804) entry for entry, including its filters and its error cases."""
import ctypes

import numpy as np
import pytest

from oracle import r1cs as orc
from oracle.params import BLS12_381, BN254

ERR_BAD_TAG, ERR_LC_INDEX, ERR_NESTED_LC, ERR_COLUMN, ERR_COEFF = 1, 2, 4, 8, 16


def run_host(lib, lm, n_instance, n_vars, cap=None):
    n_rows = len(lm["args"][0])
    args = [np.array(a, dtype=np.uint64) for a in lm["args"]]
    off = np.array(lm["offsets"], dtype=np.uint64)
    vars_ = np.array(lm["vars"], dtype=np.uint64)
    coeffs = np.array(lm["coeffs"], dtype=np.uint32)
    is_zero = np.array([1 if v == 0 else 0 for v in lm["pool"]], dtype=np.uint8)
    cap = cap if cap is not None else 3 * (len(vars_) + n_rows) + 1
    rp = [np.zeros(n_rows + 1, dtype=np.uint64) for _ in range(3)]
    col = [np.zeros(cap, dtype=np.uint32) for _ in range(3)]
    ids = [np.zeros(cap, dtype=np.uint32) for _ in range(3)]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.ht_lcmap_csr.restype = ctypes.c_uint32
    lib.ht_lcmap_csr.argtypes = [ctypes.c_uint64] * 3 + [ctypes.c_void_p] * 3 + [ctypes.c_uint64] + [ctypes.c_void_p] * 4 + [
        ctypes.c_uint32, ctypes.c_uint64] + [ctypes.c_void_p] * 9
    err = lib.ht_lcmap_csr(n_rows, n_instance, n_vars, p(args[0]), p(args[1]), p(args[2]), len(off) - 1, p(off), p(vars_), p(coeffs),
                           p(is_zero), len(lm["pool"]), cap, *[p(x) for x in rp], *[p(x) for x in col], *[p(x) for x in ids])
    mats = []
    for k in range(3):
        rows = []
        for r in range(n_rows):
            lo, hi = int(rp[k][r]), int(rp[k][r + 1])
            rows.append([(lm["pool"][int(ids[k][e])], int(col[k][e])) for e in range(lo, hi)])
        mats.append(rows)
    return err, mats


def circuits(curve):
    c2 = orc.circuit2(curve, 1, 1, 2); c2.finalize()
    out = orc.circuit2(curve, 1, 1, 2); out.set_instance_outliner("R1CS", orc.outline_r1cs); out.finalize()
    bc = orc.bench_circuit(curve, 40, seed=3); bc.finalize()
    return {"circuit2": c2, "circuit2-outlined": out, "dummy": orc.dummy_circuit(curve, 3, 5, 16, 16), "bench40": bc,
            "quirks": quirks_circuit(curve)}


def quirks_circuit(curve):
    """Everything make_row filters: zero coefficients, the Zero variable, duplicate columns, -1 and other pooled
    coefficients, the constant column, an empty row."""
    r = curve.r
    cs = orc.ConstraintSystem(curve)
    x = cs.new_input_variable(lambda: 3)
    w = [cs.new_witness_variable(lambda i=i: 5 + i) for i in range(4)]
    L = lambda *t: orc.LinearCombination(r, list(t))
    cs.enforce_r1cs_constraint(L((0, w[0]), (r - 1, w[1]), (7, orc.V_ZERO), (2, orc.V_ONE)), L((1, x), (1, x)), L((9, w[2]), (0, x)))
    cs.enforce_r1cs_constraint(L(), L((1, orc.V_ZERO)), L((1, w[3])))
    cs.enforce_r1cs_constraint(L((r - 1, orc.V_ONE), (12345, w[3]), (12345, w[0])), L((1, orc.V_ONE)), L((2, w[1]), (2, w[1]), (2, w[1])))
    return cs


@pytest.mark.parametrize("curve", [BLS12_381, BN254], ids=lambda c: c.name)
def test_lcmap_csr_equals_to_matrices(hosttest_lib, curve):
    for name, cs in circuits(curve).items():
        lm = cs.to_lcmap()
        assert lm["pool"][:2] == [1, curve.r - 1] and lm["offsets"][0] == 0 and lm["offsets"][1] == 0, name   # LC 0 is empty
        err, mats = run_host(hosttest_lib, lm, cs.num_instance_variables, cs.num_instance_variables + cs.num_witness_variables)
        assert err == 0, name
        assert mats == cs.to_matrices(), name
    q = quirks_circuit(curve).to_matrices()
    assert q[0][0] == [(curve.r - 1, 3), (2, 0)] and q[1][0] == [(1, 1), (1, 1)] and q[2][0] == [(9, 4)]      # what the filters leave
    assert q[0][1] == [] and q[1][1] == []


def test_lcmap_rejects_what_make_row_would_panic_on(hosttest_lib):
    curve = BLS12_381
    cs = orc.circuit2(curve, 1, 1, 2)                      # NOT finalized: LC e = d + d refers to LC d
    lm = cs.to_lcmap()
    n_inst, n_vars = cs.num_instance_variables, cs.num_instance_variables + cs.num_witness_variables
    err, _ = run_host(hosttest_lib, lm, n_inst, n_vars)
    assert err == ERR_NESTED_LC
    cs.finalize()
    good = cs.to_lcmap()
    assert run_host(hosttest_lib, good, n_inst, n_vars)[0] == 0
    bad = dict(good, args=[list(a) for a in good["args"]])
    bad["args"][1][0] = (4 << 61) | 99                    # SymbolicLc(99): no such LC
    assert run_host(hosttest_lib, bad, n_inst, n_vars)[0] == ERR_LC_INDEX
    bad["args"][1][0] = (5 << 61) | 1                     # tag 5 does not exist
    assert run_host(hosttest_lib, bad, n_inst, n_vars)[0] == ERR_BAD_TAG
    bad["args"][1][0] = (3 << 61) | 50                    # Witness(50) of 2
    assert run_host(hosttest_lib, bad, n_inst, n_vars)[0] == ERR_COLUMN
    bad["args"][1][0] = (2 << 61) | n_inst                # Instance(l): would alias the first witness column
    assert run_host(hosttest_lib, bad, n_inst, n_vars)[0] == ERR_COLUMN
    bad = dict(good, coeffs=[len(good["pool"])] + good["coeffs"][1:])
    assert run_host(hosttest_lib, bad, n_inst, n_vars)[0] == ERR_COEFF
    # erroneous terms are skipped by BOTH passes, so the fill never writes past what the count reserved
    err, _ = run_host(hosttest_lib, lm, n_inst, n_vars, cap=len(lm["vars"]) + 3)
    assert err == ERR_NESTED_LC
