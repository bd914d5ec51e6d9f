"""the Groth16 verification equation with REAL pairings (oracle/pairing.py, BLS12-381), i.e.
`SNARK::verify` (snark/src/lib.rs:57-66) -- independent of the setup trapdoor that the in-the-exponent check uses."""
import random

import numpy as np

from oracle import groth16 as og
from oracle import pairing as pr
from oracle import r1cs as orc
from oracle.ec import groups
from oracle.params import BLS12_381 as curve
from tests.util import pack_points, pairing_verify_packed


def test_pairing_is_bilinear_and_non_degenerate():
    G1, G2 = groups(curve)
    rng = random.Random(3)
    e = pr.pairing(G1.gen, G2.gen)
    assert not e == pr.Fq12.one() and e.pow(curve.r) == pr.Fq12.one()
    a, b = rng.randrange(1, curve.r), rng.randrange(1, curve.r)
    assert pr.pairing(G1.mul(G1.gen, a), G2.mul(G2.gen, b)) == e.pow(a * b % curve.r)
    assert pr.pairing(G1.mul(G1.gen, a), G2.gen) * pr.pairing(G1.mul(G1.gen, curve.r - a), G2.gen) == pr.Fq12.one()
    assert pr.pairing(None, G2.gen) == pr.Fq12.one()


def test_groth16_proofs_verify_under_the_pairing_equation():
    rng = random.Random(0xB2000003)
    for cs in (orc.circuit2(curve, 1, 1, 2), orc.dummy_circuit(curve, 3, 5, 8, 8), orc.bench_circuit(curve, 5, seed=4)):
        cs.finalize()
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
        pk = og.setup(curve, mats, len(inst), len(wit), td)
        rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
        A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
        vk = {"alpha_g1": pk.alpha_g1, "beta_g2": pk.beta_g2, "gamma_g2": pk.gamma_g2, "delta_g2": pk.delta_g2, "gamma_abc_g1": pk.gamma_abc_g1}
        assert pr.groth16_verify(vk, inst[1:], (A, B, C))
        G1, _ = groups(curve)
        assert not pr.groth16_verify(vk, inst[1:], (A, B, G1.add(C, G1.gen)))                 # tampered proof
        if len(inst) > 1:
            assert not pr.groth16_verify(vk, [(inst[1] + 1) % curve.r] + inst[2:], (A, B, C))  # wrong public input
        # the same check through the C-ABI array layout (what the GPU tests feed it)
        packed_vk = {"alpha_g1": pack_points(curve, 1, [pk.alpha_g1]), "beta_g2": pack_points(curve, 2, [pk.beta_g2]),
                     "gamma_g2": pack_points(curve, 2, [pk.gamma_g2]), "delta_g2": pack_points(curve, 2, [pk.delta_g2]),
                     "gamma_abc_g1": pack_points(curve, 1, pk.gamma_abc_g1)}
        proof = (pack_points(curve, 1, [A]), pack_points(curve, 2, [B]), pack_points(curve, 1, [C]))
        assert pairing_verify_packed(curve, packed_vk, len(inst), inst, proof)
