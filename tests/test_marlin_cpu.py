"""Universal-setup (Marlin-style) path on the CPU: the protocol of snark_b200/marlin.py over the big-int backend against the
independent verifier of oracle/marlin.py -- completeness, soundness smoke tests, and one real-pairing check."""
import copy
import random

import pytest

from oracle import marlin as om
from oracle import r1cs as orc
from oracle.params import BLS12_381, BN254
from snark_b200 import marlin as M


def circuits(curve):
    out = []
    for cs in (orc.circuit2(curve, 1, 1, 2), orc.dummy_circuit(curve, 3, 5, 8, 8), orc.bench_circuit(curve, 9, seed=2),
               orc.dummy_circuit(curve, 2, 7, 20, 13)):
        cs.finalize()
        assert cs.is_satisfied()
        out.append((cs.to_matrices(), list(cs.instance_assignment), list(cs.witness_assignment)))
    return out


def test_variable_positions_are_a_bijection_off_the_instance_subgroup():
    for n_inst, n_wit, n, l in ((1, 6, 8, 1), (2, 5, 8, 2), (3, 11, 16, 4), (2, 14, 16, 2), (5, 40, 64, 8)):
        pos = M.variable_positions(n_inst, n_inst + n_wit, n, l)
        assert len(set(pos)) == len(pos) and max(pos) < n
        s = n // l
        assert pos[:n_inst] == [j * s for j in range(n_inst)]
        assert all(p % s != 0 for p in pos[n_inst:])
        assert pos[n_inst:] == sorted(pos[n_inst:])


@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
def test_prove_verify_roundtrip(curve):
    rng = random.Random(0xB2000005)
    be = om.IntBackend(curve)
    for mats, x, w in circuits(curve):
        info = M.index_shape(mats, len(x), len(x) + len(w))
        tau = rng.randrange(2, curve.r)
        srs = be.setup(info.D + 1, tau)
        pk, vk = M.index(be, srs, mats, len(x), len(x) + len(w))
        proof = M.prove(be, pk, x, w, check=True)
        assert om.verify(curve, vk, x, proof, tau=tau)
        # a different instance does not verify
        x_bad = list(x)
        x_bad[-1] = (x_bad[-1] + 1) % curve.r
        if len(x) > 1:
            assert not om.verify(curve, vk, x_bad, proof, tau=tau)
        # tampering with any single evaluation / commitment / opening is caught
        for fld, i in (("evals1", 0), ("evals1", 3), ("evals1", 5), ("evals2", 0), ("evals2", 1), ("evals2", 7)):
            bad = copy.deepcopy(proof)
            getattr(bad, fld)[i] = (getattr(bad, fld)[i] + 1) % curve.r
            assert not om.verify(curve, vk, x, bad, tau=tau)
        bad = copy.deepcopy(proof)
        bad.comms[4], bad.comms[7] = bad.comms[7], bad.comms[4]
        assert not om.verify(curve, vk, x, bad, tau=tau)
        bad = copy.deepcopy(proof)
        bad.openings[0] = be.G1.dbl(bad.openings[0])
        assert not om.verify(curve, vk, x, bad, tau=tau)


def test_unsatisfied_witness_is_rejected():
    curve = BN254
    be = om.IntBackend(curve)
    mats, x, w = circuits(curve)[1]
    info = M.index_shape(mats, len(x), len(x) + len(w))
    srs = M.universal_setup(be, info.D, 0x1234567)
    pk, vk = M.index(be, srs, mats, len(x), len(x) + len(w))
    with pytest.raises(AssertionError):
        M.index(be, M.universal_setup(be, info.D - 1, 0x1234567), mats, len(x), len(x) + len(w))     # compute bound too small
    w_bad = list(w)
    w_bad[0] = (w_bad[0] + 1) % curve.r
    with pytest.raises(AssertionError):
        M.prove(be, pk, x, w_bad, check=True)
    # without the prover's own checks a proof comes out, and the verifier refuses it
    proof = M.prove(be, pk, x, w_bad, check=False)
    assert not om.verify(curve, vk, x, proof, tau=0x1234567)


def test_degree_bound_is_enforced():
    """g1 must have degree < |H| - 1: a prover that commits the shifted polynomial with the wrong shift is caught."""
    curve = BN254
    be = om.IntBackend(curve)
    mats, x, w = circuits(curve)[0]
    info = M.index_shape(mats, len(x), len(x) + len(w))
    tau = 0xABCDEF
    srs = be.setup(info.D + 1, tau)
    pk, vk = M.index(be, srs, mats, len(x), len(x) + len(w))
    proof = M.prove(be, pk, x, w, check=True)
    bad = copy.deepcopy(proof)
    bad.comms[5] = be.G1.mul(bad.comms[5], tau)       # = commitment of X^(shift + 1) g1
    assert not om.verify(curve, vk, x, bad, tau=tau)


def test_real_pairing_opening_check():
    from oracle import pairing as opair
    from oracle.ec import groups

    curve = BN254
    be = om.IntBackend(curve)
    mats, x, w = circuits(curve)[0]
    info = M.index_shape(mats, len(x), len(x) + len(w))
    tau = 0x5EED5EED5EED
    srs = be.setup(info.D + 1, tau)
    pk, vk = M.index(be, srs, mats, len(x), len(x) + len(w))
    proof = M.prove(be, pk, x, w, check=True)
    G2 = groups(curve)[1]
    tau_g2 = G2.mul(G2.gen, tau)
    eng = opair.engine(curve)
    assert om.verify(curve, vk, x, proof, tau_g2=tau_g2, engine=eng)
    bad = copy.deepcopy(proof)
    bad.evals2[3] = (bad.evals2[3] + 1) % curve.r
    assert not om.verify(curve, vk, x, bad, tau_g2=tau_g2, engine=eng)


def test_wire_form_round_trip():
    curve = BN254
    be = om.IntBackend(curve)
    mats, x, w = circuits(curve)[2]
    info = M.index_shape(mats, len(x), len(x) + len(w))
    tau = 0x77777
    pk, vk = M.index(be, M.universal_setup(be, info.D, tau), mats, len(x), len(x) + len(w))
    proof = M.prove(be, pk, x, w)
    raw = M.proof_to_bytes(proof, curve.r, be.fq_bytes)
    assert len(raw) == 4 + 10 * 64 + 4 + 2 * 64 + 8 + 20 * 32
    back = M.proof_from_bytes(raw, curve.r, be.fq_bytes)
    assert (back.comms, back.evals1, back.evals2, back.openings) == (proof.comms, proof.evals1, proof.evals2, proof.openings)
    vk2 = M.vk_from_bytes(M.vk_to_bytes(vk, be.fq_bytes), be.fq_bytes)
    assert vk2.info == vk.info and vk2.index_comms == vk.index_comms
    assert om.verify(curve, vk2, x, back, tau=tau)
    with pytest.raises(ValueError):
        M.proof_from_bytes(raw[:-1], curve.r, be.fq_bytes)
    bad = bytearray(raw)
    bad[-32:] = (curve.r).to_bytes(32, "little")                      # a non-canonical evaluation
    with pytest.raises(ValueError):
        M.proof_from_bytes(bytes(bad), curve.r, be.fq_bytes)
