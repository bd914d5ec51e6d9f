"""compressed wire format through the C ABI against the oracle's restatement and the standard
compressed BLS12-381 generators (the only externally known byte vectors available offline)."""
import random

import pytest

from oracle import serialize as oser
from oracle.ec import groups
from oracle.params import BLS12_381, BN254
from tests.util import pack_points

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]
BLS_G1_HEX = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
BLS_G2_HEX = ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
              "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


def test_compressed_encodings(be):
    curve = CURVES[be.curve]
    G1, G2 = groups(curve)
    rng = random.Random(31)
    for group, G in ((1, G1), (2, G2)):
        pts = [G.gen, G.neg(G.gen), None] + [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(20)]
        got = be.serialize_points(group, pack_points(curve, group, pts), len(pts))
        exp = b"".join(oser.point_compressed(curve, group, P) for P in pts)
        assert got == exp
        per = len(exp) // len(pts)
        if curve is BLS12_381:
            assert got[:per].hex() == (BLS_G1_HEX if group == 1 else BLS_G2_HEX)          # standard generator encodings
            assert got[per] == got[0] ^ 0x20 and got[per + 1: 2 * per] == got[1:per]       # -G flips only the sign bit
            assert got[2 * per] == 0xC0 and not any(got[2 * per + 1: 3 * per])              # infinity
        else:
            assert got[2 * per: 3 * per] == bytes(per - 1) + b"\x40"
            assert got[:32] == (1).to_bytes(32, "little") if group == 1 else True             # BN254 G1 generator x = 1, y = 2 <= -y
    A, B, C = G1.mul(G1.gen, 5), G2.mul(G2.gen, 7), G1.mul(G1.gen, curve.r - 3)
    blob = be.proof_bytes(pack_points(curve, 1, [A]), pack_points(curve, 2, [B]), pack_points(curve, 1, [C]))
    assert blob == oser.proof_compressed(curve, A, B, C)
    assert len(blob) == (192 if curve is BLS12_381 else 128)


def test_uncompressed_encodings(be):
    curve = CURVES[be.curve]
    G1, G2 = groups(curve)
    rng = random.Random(32)
    for group, G in ((1, G1), (2, G2)):
        pts = [G.gen, G.neg(G.gen), None] + [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(12)]
        got = be.serialize_points(group, pack_points(curve, group, pts), len(pts), compressed=False)
        assert got == b"".join(oser.point_uncompressed(curve, group, P) for P in pts)
    A, B, C = G1.mul(G1.gen, 11), None, G1.mul(G1.gen, curve.r - 9)
    blob = be.proof_bytes(pack_points(curve, 1, [A]), pack_points(curve, 2, [B]), pack_points(curve, 1, [C]), compressed=False)
    assert blob == oser.point_uncompressed(curve, 1, A) + oser.point_uncompressed(curve, 2, B) + oser.point_uncompressed(curve, 1, C)
    assert len(blob) == (384 if curve is BLS12_381 else 256)


@pytest.mark.parametrize("compressed", [True, False])
def test_key_serialization_of_a_gpu_setup(be, compressed):
    """ProvingKey / VerifyingKey bytes (ark-groth16 framing) of a key made by b2s_groth16_setup, the proving key streamed
    from its device-resident form, against the oracle's encoding of the oracle's own setup with the same trapdoor."""
    from oracle import groth16 as og
    from oracle import r1cs as orc
    from tests.util import csr_from_rows, pack_fr

    curve = CURVES[be.curve]
    rng = random.Random(33)
    bc = orc.bench_circuit(curve, 12, seed=4)
    bc.finalize()
    mats, inst, wit = bc.to_matrices(), bc.instance_assignment, bc.witness_assignment
    td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
    pk = og.setup(curve, mats, len(inst), len(wit), td)
    m = be.r1cs_upload(len(mats[0]), len(inst), len(wit), [csr_from_rows(curve, M) for M in mats])
    pkh, vk = be.groth16_setup(m, pack_fr(curve, [td.tau, td.alpha, td.beta, td.gamma, td.delta]), len(inst))
    vkb = be.vk_bytes(vk["alpha_g1"], vk["beta_g2"], vk["gamma_g2"], vk["delta_g2"], vk["gamma_abc_g1"], len(inst), compressed)
    exp_vk = oser.verifying_key_bytes(curve, {"alpha_g1": pk.alpha_g1, "beta_g2": pk.beta_g2, "gamma_g2": pk.gamma_g2, "delta_g2": pk.delta_g2,
                                              "gamma_abc_g1": pk.gamma_abc_g1}, compressed)
    assert vkb == exp_vk
    assert be.pk_bytes(pkh, vkb, compressed) == oser.proving_key_bytes(curve, pk, compressed)
    be.pk_free(pkh)
    be.r1cs_free(m)
