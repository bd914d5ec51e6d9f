"""Wire format on the oracle side: encode -> decode round trips for both curves, malformed inputs
rejected, uncompressed form, and the VerifyingKey / ProvingKey framing.  (The GPU encoder is compared with
`point_compressed` in tests/test_gpu_serialize.py; the standard BLS12-381 generator encodings pin it.)"""
import random

import pytest

from oracle import groth16 as og
from oracle import r1cs as orc
from oracle import serialize as oser
from oracle.ec import groups
from oracle.params import BLS12_381, BN254

CURVES = [BLS12_381, BN254]


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_compress_decompress_round_trip(curve):
    rng = random.Random(77)
    G1, G2 = groups(curve)
    for group, G in ((1, G1), (2, G2)):
        pts = [G.gen, G.neg(G.gen), None] + [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(12)]
        for P in pts:
            blob = oser.point_compressed(curve, group, P)
            assert oser.point_decompress(curve, group, blob) == P
            if P is not None:                                   # the sign flag alone selects between P and -P
                other = oser.point_compressed(curve, group, G.neg(P))
                assert other != blob and oser.point_decompress(curve, group, other) == G.neg(P)
                diff = [a ^ b for a, b in zip(blob, other)]
                assert sum(1 for d in diff if d) == 1 and max(diff) in (0x20, 0x80)
    A, B, C = G1.mul(G1.gen, 5), G2.mul(G2.gen, 7), G1.mul(G1.gen, curve.r - 3)
    assert oser.proof_decompress(curve, oser.proof_compressed(curve, A, B, C)) == (A, B, C)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_malformed_encodings_are_rejected(curve):
    G1, _ = groups(curve)
    good = bytearray(oser.point_compressed(curve, 1, G1.gen))
    with pytest.raises(ValueError):
        oser.point_decompress(curve, 1, bytes(good[:-1]))                       # short
    # an x with no point above it: walk x upwards from the generator's until x^3 + b is a non-residue
    x = G1.gen[0]
    while oser._sqrt_fq(curve.p, (x ** 3 + curve.b) % curve.p) is not None:
        x += 1
    fq = len(good)
    raw = x.to_bytes(fq, "big") if curve is BLS12_381 else x.to_bytes(fq, "little")
    bad = bytearray(raw)
    if curve is BLS12_381:
        bad[0] |= 0x80
        with pytest.raises(ValueError):
            oser.point_decompress(curve, 1, bytes([good[0] & 0x7F]) + bytes(good[1:]))   # compression bit missing
        inf = bytearray(oser.point_compressed(curve, 1, None)); inf[-1] = 1
    else:
        with pytest.raises(ValueError):
            oser.point_decompress(curve, 1, bytes(good[:-1]) + bytes([good[-1] | 0xC0]))  # both flags
        inf = bytearray(oser.point_compressed(curve, 1, None)); inf[0] = 1
    with pytest.raises(ValueError):
        oser.point_decompress(curve, 1, bytes(bad))                             # not on the curve
    with pytest.raises(ValueError):
        oser.point_decompress(curve, 1, bytes(inf))                             # infinity with payload
    over = (curve.p).to_bytes(fq, "big" if curve is BLS12_381 else "little")    # x = p: not canonical
    over = bytes([over[0] | 0x80]) + over[1:] if curve is BLS12_381 else over
    if curve is BLS12_381 or not (over[-1] & 0xC0):
        with pytest.raises(ValueError):
            oser.point_decompress(curve, 1, over)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_uncompressed_and_key_framing(curve):
    G1, G2 = groups(curve)
    fq = 48 if curve is BLS12_381 else 32
    P, Q = G1.mul(G1.gen, 11), G2.mul(G2.gen, 13)
    u1, u2 = oser.point_uncompressed(curve, 1, P), oser.point_uncompressed(curve, 2, Q)
    assert len(u1) == 2 * fq and len(u2) == 4 * fq
    if curve is BLS12_381:
        assert u1 == P[0].to_bytes(fq, "big") + P[1].to_bytes(fq, "big") and not u1[0] & 0xE0
        assert u2[:fq] == Q[0][1].to_bytes(fq, "big") and u2[3 * fq:] == Q[1][0].to_bytes(fq, "big")
        assert oser.point_uncompressed(curve, 1, None) == b"\x40" + bytes(2 * fq - 1)
    else:
        assert u1[:fq] == P[0].to_bytes(fq, "little") and u1[fq:-1] == P[1].to_bytes(fq, "little")[:-1]
        assert (u1[-1] & 0x80 != 0) == (P[1] > curve.p - P[1]) and u1[-1] & 0x3F == P[1].to_bytes(fq, "little")[-1]
        assert oser.point_uncompressed(curve, 1, None) == bytes(2 * fq - 1) + b"\x40"
    # the compressed x bytes are the uncompressed x bytes up to the flag bits
    c1 = oser.point_compressed(curve, 1, P)
    assert (c1[1:] == u1[1:fq]) if curve is BLS12_381 else (c1[:-1] == u1[:fq - 1])
    # keys: sizes follow from the framing (8-byte little-endian Vec lengths)
    cs = orc.circuit2(curve, 1, 1, 2)
    cs.finalize()
    mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
    pk = og.setup(curve, mats, len(inst), len(wit), og.Trapdoor(3, 5, 7, 11, 13))
    vk = {"alpha_g1": pk.alpha_g1, "beta_g2": pk.beta_g2, "gamma_g2": pk.gamma_g2, "delta_g2": pk.delta_g2, "gamma_abc_g1": pk.gamma_abc_g1}
    for comp, g1, g2 in ((True, fq, 2 * fq), (False, 2 * fq, 4 * fq)):
        vb = oser.verifying_key_bytes(curve, vk, comp)
        assert len(vb) == g1 + 3 * g2 + 8 + len(inst) * g1
        assert vb[g1 + 3 * g2: g1 + 3 * g2 + 8] == len(inst).to_bytes(8, "little")
        pb = oser.proving_key_bytes(curve, pk, comp)
        n_vars, N = len(inst) + len(wit), pk.domain
        assert pb[:len(vb)] == vb
        assert len(pb) == len(vb) + 2 * g1 + 5 * 8 + (2 * n_vars + (N - 1) + len(wit)) * g1 + n_vars * g2
    # every compressed element of the key decodes back
    vb = oser.verifying_key_bytes(curve, vk, True)
    assert oser.point_decompress(curve, 1, vb[:fq]) == pk.alpha_g1 and oser.point_decompress(curve, 2, vb[fq:3 * fq]) == pk.beta_g2
