"""Device field / group arithmetic (the templates every kernel is built from) against the oracle.
GPU only: these call b2s_field_op / b2s_group_op through the C ABI."""
import random

import numpy as np
import pytest

from oracle.ec import groups
from oracle.params import BLS12_381, BN254
from tests.util import fq_limbs, pack_points, pack_u32, unpack_points, unpack_u32

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


def test_field_ops(be):
    curve = CURVES[be.curve]
    for field, p in ((0, curve.p), (1, curve.r)):
        n = (be.fq_bytes if field == 0 else be.fr_bytes) // 4
        R = 1 << (32 * n)
        Rinv = pow(R, -1, p)
        rng = random.Random(42 + field)
        edge = [0, 1, 2, p - 1, p - 2, R % p, R * R % p, (p - 1) // 2]
        a = edge + [rng.randrange(p) for _ in range(4000)]
        b = [rng.randrange(p) for _ in range(4000)] + edge
        A, B = pack_u32(a, n), pack_u32(b, n)
        cases = {
            0: lambda x, y: x * y * Rinv % p, 1: lambda x, y: (x + y) % p, 2: lambda x, y: (x - y) % p,
            4: lambda x, y: (-x) % p, 5: lambda x, y: x * R % p, 6: lambda x, y: x * Rinv % p,
            7: lambda x, y: x * x * Rinv % p,
        }
        for op, fn in cases.items():
            got = unpack_u32(be.field_op(field, op, A, B), n)
            assert got == [fn(x, y) for x, y in zip(a, b)], (field, op)
        got = unpack_u32(be.field_op(field, 3, A[: 64 * n], B[: 64 * n]), n)
        assert got == [(pow(x * Rinv % p, -1, p) * R % p if x else 0) for x in a[:64]]


@pytest.mark.parametrize("group", [1, 2])
def test_group_ops(be, group):
    curve = CURVES[be.curve]
    G = groups(curve)[group - 1]
    rng = random.Random(5 + group)
    P = [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(6)]
    a = [P[0], P[1], P[2], None, P[3], None, P[4], P[5]]
    b = [P[1], P[1], G.neg(P[2]), P[0], None, None, P[5], P[4]]
    ks = [0, 1, 2, curve.r - 1, curve.r, rng.randrange(curve.r), rng.randrange(curve.r), 3]
    A, B = pack_points(curve, group, a), pack_points(curve, group, b)
    K = pack_u32(ks, 8)
    exp_add = [G.add(x, y) for x, y in zip(a, b)]
    assert unpack_points(curve, group, be.group_op(group, 0, A, B, K)) == exp_add      # mixed add incl. P+P, P-P, inf
    assert unpack_points(curve, group, be.group_op(group, 1, A, B, K)) == exp_add      # general add
    assert unpack_points(curve, group, be.group_op(group, 2, A, B, K)) == [G.dbl(x) for x in a]
    assert unpack_points(curve, group, be.group_op(group, 3, A, B, K)) == [G.mul(x, k) for x, k in zip(a, ks)]


@pytest.mark.parametrize("group", [1, 2])
def test_fixed_base(be, group):
    curve = CURVES[be.curve]
    G = groups(curve)[group - 1]
    rng = random.Random(9)
    ks = [0, 1, 2, 255, 256, curve.r - 1] + [rng.randrange(curve.r) for _ in range(10)]
    out = be.fixed_base(group, pack_u32(ks, 8), len(ks), mont=False)
    assert unpack_points(curve, group, out) == [G.mul(G.gen, k) for k in ks]
    R = 1 << 256
    out = be.fixed_base(group, pack_u32([k * R % curve.r for k in ks], 8), len(ks), mont=True)
    assert unpack_points(curve, group, out) == [G.mul(G.gen, k) for k in ks]
