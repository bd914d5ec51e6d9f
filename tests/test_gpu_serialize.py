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
