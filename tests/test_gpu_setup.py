"""`CircuitSpecificSetupSNARK::setup` on the GPU (b2s_groth16_setup) against the oracle's
known-trapdoor generator (oracle/groth16.py: setup), element by element, and a proof made with the GPU-built key."""
import random
import time

import numpy as np
import pytest

from oracle import groth16 as og
from oracle import r1cs as orc
from oracle.ec import groups
from oracle.params import BLS12_381, BN254
from tests.util import csr_from_rows, pack_fr, pairing_verify_packed, unpack_points

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


def test_setup_matches_oracle_and_proves(be):
    curve = CURVES[be.curve]
    rng = random.Random(0xB2000003)
    cases = []
    cs = orc.circuit2(curve, 1, 1, 2); cs.finalize(); cases.append(("circuit2", cs))
    cases.append(("dummy", orc.dummy_circuit(curve, 3, 5, 16, 16)))
    bc = orc.bench_circuit(curve, 25, seed=9); bc.finalize(); cases.append(("bench25", bc))
    # instance-outlined system (instance_outliner.rs:40-60): two extra rows, unsorted columns in the rewritten LCs
    oc = orc.circuit2(curve, 1, 1, 2); oc.set_instance_outliner("R1CS", orc.outline_r1cs); oc.finalize(); cases.append(("circuit2-outlined", oc))
    for name, cs in cases:
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
        pk = og.setup(curve, mats, len(inst), len(wit), td)
        m = be.r1cs_upload(len(mats[0]), len(inst), len(wit), [csr_from_rows(curve, M) for M in mats])
        pkh, vk = be.groth16_setup(m, pack_fr(curve, [td.tau, td.alpha, td.beta, td.gamma, td.delta]), len(inst))
        n_vars = len(inst) + len(wit)
        assert unpack_points(curve, 1, be.pk_query(pkh, 0, n_vars)) == pk.a_query, name
        assert unpack_points(curve, 1, be.pk_query(pkh, 1, n_vars)) == pk.b_g1_query, name
        assert unpack_points(curve, 2, be.pk_query(pkh, 2, n_vars)) == pk.b_g2_query, name
        assert unpack_points(curve, 1, be.pk_query(pkh, 3, pk.domain - 1)) == pk.h_query, name
        assert unpack_points(curve, 1, be.pk_query(pkh, 4, len(wit))) == pk.l_query, name
        assert unpack_points(curve, 1, be.pk_query(pkh, 5, 3)) == [pk.alpha_g1, pk.beta_g1, pk.delta_g1]
        assert unpack_points(curve, 2, be.pk_query(pkh, 6, 2)) == [pk.beta_g2, pk.delta_g2]
        assert unpack_points(curve, 1, vk["alpha_g1"]) == [pk.alpha_g1]
        assert unpack_points(curve, 2, vk["beta_g2"]) == [pk.beta_g2]
        assert unpack_points(curve, 2, vk["gamma_g2"]) == [pk.gamma_g2]
        assert unpack_points(curve, 2, vk["delta_g2"]) == [pk.delta_g2]
        assert unpack_points(curve, 1, vk["gamma_abc_g1"]) == pk.gamma_abc_g1
        rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
        A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
        a, b, c = be.groth16_prove(pkh, m, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [rr]), pack_fr(curve, [ss]))
        assert (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b)[0], unpack_points(curve, 1, c)[0]) == (A, B, C), name
        # SNARK::verify with real pairings on the GPU's own key and proof (no trapdoor involved)
        assert pairing_verify_packed(curve, vk, len(inst), inst, (a, b, c)), name
        be.pk_free(pkh); be.r1cs_free(m)


def test_setup_dummy_2p18_spot_checks(be):
    """DummyCircuit shape at domain 2^18 (one column holds every row): a few key elements against the closed form."""
    curve = CURVES[be.curve]
    r = curve.r
    rng = random.Random(5)
    N = 1 << 18
    n_rows, n_inst, n_wit = N - 2, 2, N - 3
    one = pack_fr(curve, [1])
    nnz = n_rows - 1
    row_ptr = np.minimum(np.arange(n_rows + 1, dtype=np.uint64), np.uint64(nnz))
    csr = [(row_ptr, np.full(nnz, col, dtype=np.uint32), np.tile(one, nnz)) for col in (2, 3, 1)]
    m = be.r1cs_upload(n_rows, n_inst, n_wit, csr)
    td = og.Trapdoor(*[rng.randrange(1, r) for _ in range(5)])
    t0 = time.time()
    pkh, vk = be.groth16_setup(m, pack_fr(curve, [td.tau, td.alpha, td.beta, td.gamma, td.delta]), n_inst)
    dt = time.time() - t0
    print(f"\nGPU setup, DummyCircuit shape, domain 2^18: {dt:.3f} s")
    G1, G2 = groups(curve)
    # S = sum_{i < n_rows - 1} L_i(tau) = 1 - sum of the remaining three Lagrange coefficients
    log_n = 18
    w = curve.omega(log_n)
    zt = (pow(td.tau, N, r) - 1) % r
    L = lambda i: zt * pow(w, i, r) % r * pow(N, -1, r) % r * pow((td.tau - pow(w, i, r)) % r, -1, r) % r
    S = (1 - L(n_rows - 1) - L(n_rows) - L(n_rows + 1)) % r
    a_q = unpack_points(curve, 1, be.pk_query(pkh, 0, n_inst + n_wit)[: 5 * be.g1_bytes // 4])
    assert a_q[0] == G1.mul(G1.gen, L(n_rows)) and a_q[1] == G1.mul(G1.gen, L(n_rows + 1))     # input-consistency rows
    assert a_q[2] == G1.mul(G1.gen, S) and a_q[3] is None and a_q[4] is None
    b2 = unpack_points(curve, 2, be.pk_query(pkh, 2, n_inst + n_wit)[: 5 * be.g2_bytes // 4])
    assert b2[3] == G2.mul(G2.gen, S) and b2[2] is None
    hq = unpack_points(curve, 1, be.pk_query(pkh, 3, N - 1)[-2 * be.g1_bytes // 4:])
    dinv = pow(td.delta, -1, r)
    assert hq[-1] == G1.mul(G1.gen, pow(td.tau, N - 2, r) * zt % r * dinv % r)
    abc = unpack_points(curve, 1, vk["gamma_abc_g1"])
    assert abc[1] == G1.mul(G1.gen, (td.beta * L(n_rows + 1) + S) * pow(td.gamma, -1, r) % r)   # c is instance 1: C column
    be.pk_free(pkh); be.r1cs_free(m)
