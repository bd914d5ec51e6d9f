"""Pins the pure-Python oracle: against the reference's golden vectors where the reference has them
(R1CS matrices / satisfiability: gr1cs/tests/{circuit1,circuit2}.rs, tests/mod.rs) and against
implementation-independent definitions everywhere else (parity for MSM / NTT / proofs
is otherwise unpinned because the reference tree holds none of that arithmetic)."""
import random

import pytest

from oracle import groth16 as og
from oracle import msm as omsm
from oracle import ntt as ontt
from oracle import r1cs as orc
from oracle.ec import groups
from oracle.params import BLS12_381, BN254

CURVES = [BLS12_381, BN254]


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_curve_constants(curve):
    G1, G2 = groups(curve)
    assert G1.on_curve(G1.gen) and G2.on_curve(G2.gen)
    assert G1.mul(G1.gen, curve.r) is None and G2.mul(G2.gen, curve.r) is None   # generator order
    assert G1.mul(G1.gen, curve.r - 1) == G1.neg(G1.gen)
    S = curve.fr_two_adicity
    assert (curve.r - 1) % (1 << S) == 0 and ((curve.r - 1) >> S) % 2 == 1      # two-adicity
    w = curve.fr_root_of_unity
    assert pow(w, 1 << (S - 1), curve.r) == curve.r - 1                           # exact order 2^S
    assert pow(curve.fr_generator, (curve.r - 1) // 2, curve.r) == curve.r - 1   # generator is a non-residue
    if curve is BLS12_381:  # SURVEY Appendix B value
        assert w == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    else:                   # the 2^28-th root of unity published with ark-bn254 / snarkjs (5^((r-1)/2^28))
        assert w == 19103219067921713944291392827692070036145651957329286315305642004821462161904


def test_published_bn254_doubling_vector():
    """A known answer from outside this repository: 2 * (1, 2) on alt_bn128, the EIP-196 ecAdd / ecMul test vector."""
    G1 = groups(BN254)[0]
    expect = (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
              0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)
    assert G1.mul(G1.gen, 2) == expect and G1.add(G1.gen, G1.gen) == expect
    assert expect == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                      9918110051302171585080402603319702774565515993150576347155970296011118125764)


def test_published_bls12_381_doubling_vector():
    """2 * G1 in the zcash compressed form, as it appears in the BLS12-381 test suites of other libraries."""
    from oracle import serialize as oser

    G1 = groups(BLS12_381)[0]
    assert oser.point_compressed(BLS12_381, 1, G1.mul(G1.gen, 2)).hex() == (
        "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e")


def test_reference_golden_circuit2():
    """test_circuit2_matrices (gr1cs/tests/mod.rs:136-147): matrices after finalize == circuit2.rs:21-43."""
    cs = orc.circuit2(BLS12_381, 1, 1, 2)
    cs.finalize()
    assert cs.to_matrices() == orc.CIRCUIT2_GOLDEN
    assert cs.is_satisfied()
    z = cs.z()
    A, B, C = cs.to_matrices()
    r = BLS12_381.r
    az, bz, cz = (orc.mat_vec_mul(r, M, z) for M in (A, B, C))
    assert (az, bz, cz) == ([1, 1, 1], [2, 2, 4], [2, 2, 4])                      # 1*2=2, 1*2=2, 1*4=4
    assert [orc.evaluate_constraint(r, row, z) for row in B] == bz
    # a wrong witness is rejected
    bad = orc.circuit2(BLS12_381, 1, 1, 3)
    assert bad.which_is_unsatisfied() == ("R1CS", 0)


def test_reference_golden_circuit1():
    """test_circuit1_matrices / _sat / _non_sat (gr1cs/tests/mod.rs:17-103)."""
    cs = orc.circuit1(BLS12_381, (0,) * 5, (0,) * 8)
    assert cs.to_matrices_all() == orc.CIRCUIT1_GOLDEN          # before finalize, as the reference asserts
    sat = orc.circuit1(BLS12_381, *orc.CIRCUIT1_SAT)
    sat.finalize()
    assert sat.is_satisfied()
    assert not orc.circuit1(BLS12_381, *orc.CIRCUIT1_UNSAT).is_satisfied()


def test_circuit1_instance_outlined():
    """test_circuit1_instance_outlined (gr1cs/tests/mod.rs:105-131) + what outlining must preserve."""
    cs = orc.circuit1(BLS12_381, (0,) * 5, (0,) * 8)
    num_instance, prev_num_witness = cs.num_instance_variables, cs.num_witness_variables
    cs.set_instance_outliner("R1CS", orc.outline_r1cs)
    cs.finalize()
    assert num_instance == cs.num_witness_variables - prev_num_witness
    assert cs.instance_outliner is None                                  # taken by finalize (constraint_system.rs:699)
    # an outliner naming an unregistered predicate is dropped silently (constraint_system.rs:701)
    cs = orc.circuit2(BLS12_381, 1, 1, 2)
    cs.set_instance_outliner("no-such-predicate", orc.outline_r1cs)
    cs.finalize()
    assert cs.num_witness_variables == 2 and cs.to_matrices() == orc.CIRCUIT2_GOLDEN


def test_outlined_circuit2_semantics():
    """After outline_r1cs the system has l more witnesses and l more constraints (one*one = One; one*w_i = x_i), stays
    satisfied by the same instance, rejects a wrong one, and stored LCs no longer mention instance columns."""
    r = BLS12_381.r
    cs = orc.circuit2(BLS12_381, 1, 1, 2)
    cs.set_instance_outliner("R1CS", orc.outline_r1cs)
    cs.finalize()
    ell = cs.num_instance_variables
    assert (ell, cs.num_witness_variables, cs.num_constraints()) == (2, 4, 5)
    assert cs.witness_assignment == [1, 2, 1, 1] and cs.is_satisfied()   # copies: one_w = 1, w(x1) = 1
    A, B, C = cs.to_matrices()
    # rows 0..2: LCs that were stored in lc_map lost their instance / One columns; bare variables (row 0: A = x1) keep them
    assert A[0] == [(1, 1)] and B[1] == [(1, 5), (1, 2)] and C[2] == [(2, 5), (2, 2)]   # replaced in place: rows unsorted
    assert A[2] == [(1, 0)]                                             # `lc!() + One` is a bare variable, not an LC
    # the new rows: (one_w, one_w, One) and (one_w, w_x1, x1)
    assert (A[3], B[3], C[3]) == ([(1, 4)], [(1, 4)], [(1, 0)])
    assert (A[4], B[4], C[4]) == ([(1, 4)], [(1, 5)], [(1, 1)])
    z = cs.z()
    az, bz, cz = (orc.mat_vec_mul(r, M, z) for M in (A, B, C))
    assert all(a * b % r == c for a, b, c in zip(az, bz, cz))
    cs.instance_assignment[1] = 7                                        # the tie one*w = x now fails
    assert cs.which_is_unsatisfied() == ("R1CS", 0)
    # setup mode: same shape, closures never evaluated
    st = orc.ConstraintSystem(BLS12_381, setup_mode=True)
    v = st.new_input_variable(lambda: 1 / 0)
    w = st.new_witness_variable(lambda: 1 / 0)
    st.enforce_r1cs_constraint(orc.lc(r, v, w), orc.lc(r, w), orc.lc(r, v))
    st.set_instance_outliner("R1CS", orc.outline_r1cs)
    st.finalize()
    assert (st.num_instance_variables, st.num_witness_variables, st.num_constraints()) == (2, 3, 3)
    assert st.to_matrices()[0][0] == [(1, 4), (1, 2)]                   # x1 + w0 became w(x1) + w0, order kept


def test_dummy_circuit_shapes():
    """DummyCircuit (sr1cs/mod.rs:296-317): builder output == the direct generator used at scale."""
    for curve in CURVES:
        cs = orc.dummy_circuit(curve, 3, 5, 16, 16)
        mats, inst, wit = orc.dummy_circuit_direct(curve, 3, 5, 16, 16)
        assert cs.to_matrices() == mats and cs.instance_assignment == inst and cs.witness_assignment == wit
        assert cs.is_satisfied() and cs.num_instance_variables == 2 and cs.num_witness_variables == 15
        assert mats[0][-1] == [] and mats[1][-1] == [] and mats[2][-1] == []


def test_lc_quirks():
    """`lc + var` on a short LC inserts a duplicate; compactify merges on inlining."""
    r = BLS12_381.r
    v = orc.witness(0)
    l = orc.LinearCombination(r) + v + v
    assert l.t == [(1, v), (1, v)]
    l.compactify()
    assert l.t == [(2, v)]
    assert orc.variable_index(orc.V_ONE, 5) == 0 and orc.variable_index(orc.instance(3), 5) == 3
    assert orc.variable_index(orc.witness(2), 5) == 7 and orc.variable_index(orc.symbolic_lc(1), 5) is None
    assert sorted([orc.symbolic_lc(0), orc.witness(9), orc.instance(1), orc.V_ONE, orc.V_ZERO]) == [
        orc.V_ZERO, orc.V_ONE, orc.instance(1), orc.witness(9), orc.symbolic_lc(0)]  # variable.rs:206-266


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_ntt_against_definition(curve):
    rng = random.Random(1)
    for log_n in range(0, 8):
        x = [rng.randrange(curve.r) for _ in range(1 << log_n)]
        assert ontt.ntt(curve, x) == ontt.dft_naive(curve, x)
        assert ontt.ntt(curve, x, inverse=True) == ontt.dft_naive(curve, x, inverse=True)
        assert ontt.ntt(curve, ontt.ntt(curve, x), inverse=True) == x
        assert ontt.coset_intt(curve, ontt.coset_ntt(curve, x)) == x
        g, w = curve.fr_generator, curve.omega(log_n)
        direct = [sum(x[j] * pow(g * pow(w, i, curve.r), j, curve.r) for j in range(len(x))) % curve.r for i in range(len(x))]
        assert ontt.coset_ntt(curve, x) == direct


@pytest.mark.parametrize("curve,group", [(BLS12_381, 1), (BLS12_381, 2), (BN254, 1), (BN254, 2)], ids=str)
def test_msm_against_definition(curve, group):
    G = groups(curve)[group - 1]
    rng = random.Random(2)
    bases = [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(40)]
    s = [rng.randrange(curve.r) for _ in range(40)]
    t = [rng.randrange(curve.r) for _ in range(40)]
    assert omsm.msm_pippenger(G, bases, s) == omsm.msm_naive(G, bases, s)
    lhs = omsm.msm_pippenger(G, bases, [(a + b) % curve.r for a, b in zip(s, t)])
    assert lhs == G.add(omsm.msm_pippenger(G, bases, s), omsm.msm_pippenger(G, bases, t))   # linearity
    assert omsm.msm_pippenger(G, bases[:1], [0]) is None


def test_window_rule_and_add_counts():
    """SURVEY App. A.4 / 8d figures."""
    assert [omsm.ark_window_bits(1 << k) for k in (16, 20, 22, 24, 26)] == [13, 15, 17, 18, 19]
    assert omsm.ark_window_bits(31) == 3
    assert omsm.reference_add_count(1 << 22, 255) == (1 << 22) * 15 + 15 * (1 << 17) == 64880640
    assert abs(omsm.reference_add_count(1 << 24, 255) - 2.56e8) < 0.01e8
    assert abs(omsm.reference_add_count(1 << 26, 255) - 9.47e8) < 0.01e8
    for k in (0, 1, (1 << 255) - 19, BLS12_381.r - 1):
        for c in (3, 13, 16, 17):
            d = omsm.signed_digits(k, c, 255)
            assert sum(v << (c * i) for i, v in enumerate(d)) == k


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_groth16_known_trapdoor(curve):
    rng = random.Random(0xB2000003)
    outlined = orc.circuit2(curve, 1, 1, 2)
    outlined.set_instance_outliner("R1CS", orc.outline_r1cs)
    for cs in (orc.circuit2(curve, 1, 1, 2), orc.dummy_circuit(curve, 3, 5, 8, 8), orc.bench_circuit(curve, 6, seed=1), outlined):
        cs.finalize()
        assert cs.is_satisfied()
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
        pk = og.setup(curve, mats, len(inst), len(wit), td)
        rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
        A, B, C, h = og.prove(pk, mats, inst, wit, rr, ss)
        assert h[-1] == 0                                           # deg h <= N - 2
        assert og.check_in_exponent(pk, (A, B, C), inst, wit, h, rr, ss)
        exps = og.expected_proof_exponents(pk, inst, wit, h, rr, ss)
        assert og.verify_equation_in_exponent(pk, inst, *exps)     # e(A,B) = e(alpha,beta) e(IC,gamma) e(C,delta)
        # an unsatisfying witness does not verify
        bad = list(wit)
        bad[0] = (bad[0] + 1) % curve.r
        _, _, _, hb = og.prove(pk, mats, inst, bad, rr, ss)
        eb = og.expected_proof_exponents(pk, inst, bad, hb, rr, ss)
        assert not og.verify_equation_in_exponent(pk, inst, *eb)


def test_compressed_encoding_known_vectors():
    """the standard compressed BLS12-381 generators pin the zcash-form encoder."""
    from oracle import serialize as oser

    G1, G2 = groups(BLS12_381)
    assert oser.point_compressed(BLS12_381, 1, G1.gen).hex() == (
        "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
    assert oser.point_compressed(BLS12_381, 2, G2.gen).hex() == (
        "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
        "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")
    assert oser.point_compressed(BLS12_381, 1, None) == bytes([0xC0]) + bytes(47)
    assert oser.point_compressed(BLS12_381, 1, G1.neg(G1.gen))[0] == 0xB7
    g = groups(BN254)[0]
    assert oser.point_compressed(BN254, 1, g.gen) == (1).to_bytes(32, "little")
    assert oser.point_compressed(BN254, 1, g.neg(g.gen))[-1] & 0x80
    assert len(oser.proof_compressed(BN254, g.gen, groups(BN254)[1].gen, None)) == 128


@pytest.mark.parametrize("curve", [BLS12_381, BN254], ids=["bls12_381", "bn254"])
def test_witness_map_six_transform_identity(curve):
    """The product computes h with 6 transforms instead of ark-groth16's 7 (snark_b200/csrc/r1cs.cu, witness_map_t):
    Z is the constant g^N - 1 on the coset and deg C < N, so the c term needs no coset round trip.  Pin the identity, in
    the composed form and in the merged-scaling form the kernels use, against the oracle's 7-transform witness_map -- for a
    satisfying AND a non-satisfying assignment (the identity does not depend on a*b = c holding on the domain)."""
    r, g = curve.r, curve.fr_generator
    for cs in (orc.circuit2(curve, 1, 1, 2), orc.dummy_circuit(curve, 3, 5, 8, 8), orc.bench_circuit(curve, 11, seed=3)):
        cs.finalize()
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        for bump in (0, 1):
            z = list(inst) + list(wit)
            z[-1] = (z[-1] + bump) % r
            want = og.witness_map(curve, mats, z, len(inst))
            A, B, C = mats
            n, N = len(A), og.domain_size(len(A), len(inst))
            a = orc.mat_vec_mul(r, A, z) + [0] * (N - n)
            b = orc.mat_vec_mul(r, B, z) + [0] * (N - n)
            c = orc.mat_vec_mul(r, C, z) + [0] * (N - n)
            for i in range(len(inst)):
                a[n + i] = z[i]
            zinv = pow((pow(g, N, r) - 1) % r, -1, r)
            n_inv = pow(N, -1, r)
            # composed: standard transforms, h = (cosetiNTT(a_coset * b_coset) - iNTT(c)) * Zinv
            ac, bc, cc = (ontt.ntt(curve, v, inverse=True) for v in (a, b, c))
            ae, be = ontt.coset_ntt(curve, ac), ontt.coset_ntt(curve, bc)
            q = ontt.coset_intt(curve, [x * y % r for x, y in zip(ae, be)])
            assert [(x - y) * zinv % r for x, y in zip(q, cc)] == want
            # merged: unscaled inverse transforms (N * coefficients), input scaling g^j / N, output scaling g^-j Zinv / N,
            # c's 1/N and Zinv folded into the last subtraction
            raw = [[x * N % r for x in v] for v in (ac, bc, cc)]
            ae2 = ontt.ntt(curve, [x * pow(g, j, r) * n_inv % r for j, x in enumerate(raw[0])])
            be2 = ontt.ntt(curve, [x * pow(g, j, r) * n_inv % r for j, x in enumerate(raw[1])])
            assert ae2 == ae and be2 == be
            q_raw = [x * N % r for x in ontt.ntt(curve, [x * y % r for x, y in zip(ae2, be2)], inverse=True)]
            g_inv = pow(g, -1, r)
            qz = [x * pow(g_inv, j, r) * n_inv * zinv % r for j, x in enumerate(q_raw)]
            beta = zinv * n_inv % r
            assert [(x - y * beta) % r for x, y in zip(qz, raw[2])] == want
