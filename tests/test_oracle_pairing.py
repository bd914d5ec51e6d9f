"""the Groth16 verification equation with REAL pairings (oracle/pairing.py, both curves), i.e.
`SNARK::verify` (snark/src/lib.rs:57-66) -- independent of the setup trapdoor that the in-the-exponent check uses."""
import random

import pytest

from oracle import groth16 as og
from oracle import pairing as pr
from oracle import r1cs as orc
from oracle.ec import groups
from oracle.params import BLS12_381, BN254
from tests.util import pack_points, pairing_verify_packed


CURVES = [BLS12_381, BN254]


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_pairing_is_bilinear_and_non_degenerate(curve):
    G1, G2 = groups(curve)
    E = pr.engine(curve)
    one = E.Fq12.one()
    rng = random.Random(3)
    e = E.pairing(G1.gen, G2.gen)
    assert not e == one and e.pow(curve.r) == one
    a, b = rng.randrange(1, curve.r), rng.randrange(1, curve.r)
    assert E.pairing(G1.mul(G1.gen, a), G2.mul(G2.gen, b)) == e.pow(a * b % curve.r)
    assert E.pairing(G1.mul(G1.gen, a), G2.gen) * E.pairing(G1.mul(G1.gen, curve.r - a), G2.gen) == one
    # additive in each argument separately
    Pa, Pb = G1.mul(G1.gen, a), G1.mul(G1.gen, b)
    assert E.pairing(G1.add(Pa, Pb), G2.gen) == E.pairing(Pa, G2.gen) * E.pairing(Pb, G2.gen)
    Qa, Qb = G2.mul(G2.gen, a), G2.mul(G2.gen, b)
    assert E.pairing(G1.gen, G2.add(Qa, Qb)) == E.pairing(G1.gen, Qa) * E.pairing(G1.gen, Qb)
    assert E.pairing(None, G2.gen) == one and E.pairing(G1.gen, None) == one
    # the twisted generator really lies on y^2 = x^3 + b over Fq12
    x, y = E.twist(G2.gen)
    assert y * y == x * x * x + E.embed_fq(curve.b)


def test_module_shorthands_are_bls12_381():
    G1, G2 = groups(BLS12_381)
    assert pr.pairing(G1.gen, G2.gen) == pr.engine(BLS12_381).pairing(G1.gen, G2.gen) and pr.Fq12 is pr.engine(BLS12_381).Fq12


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_groth16_proofs_verify_under_the_pairing_equation(curve):
    rng = random.Random(0xB2000003)
    for cs in (orc.circuit2(curve, 1, 1, 2), orc.dummy_circuit(curve, 3, 5, 8, 8), orc.bench_circuit(curve, 5, seed=4)):
        cs.finalize()
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
        pk = og.setup(curve, mats, len(inst), len(wit), td)
        rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
        A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
        vk = {"alpha_g1": pk.alpha_g1, "beta_g2": pk.beta_g2, "gamma_g2": pk.gamma_g2, "delta_g2": pk.delta_g2, "gamma_abc_g1": pk.gamma_abc_g1}
        assert pr.groth16_verify(vk, inst[1:], (A, B, C), curve)
        G1, _ = groups(curve)
        assert not pr.groth16_verify(vk, inst[1:], (A, B, G1.add(C, G1.gen)), curve)                 # tampered proof
        if len(inst) > 1:
            assert not pr.groth16_verify(vk, [(inst[1] + 1) % curve.r] + inst[2:], (A, B, C), curve)  # wrong public input
        # the same check through the C-ABI array layout (what the GPU tests feed it)
        packed_vk = {"alpha_g1": pack_points(curve, 1, [pk.alpha_g1]), "beta_g2": pack_points(curve, 2, [pk.beta_g2]),
                     "gamma_g2": pack_points(curve, 2, [pk.gamma_g2]), "delta_g2": pack_points(curve, 2, [pk.delta_g2]),
                     "gamma_abc_g1": pack_points(curve, 1, pk.gamma_abc_g1)}
        proof = (pack_points(curve, 1, [A]), pack_points(curve, 2, [B]), pack_points(curve, 1, [C]))
        assert pairing_verify_packed(curve, packed_vk, len(inst), inst, proof)
