"""The kernels' group-law templates (snark_b200/csrc/ec.cuh) on the host against the oracle's
affine/Jacobian arithmetic, including every special case of the addition law."""
import random

import numpy as np
import pytest

from oracle.ec import groups
from oracle.params import BLS12_381, BN254
from tests.util import fq_limbs, pack_points, pack_u32, ptr, unpack_points

CASES = [(BLS12_381, 1), (BLS12_381, 2), (BN254, 1), (BN254, 2)]


def call(lib, curve, group, op, a, b, k=0):
    n = fq_limbs(curve)
    out = np.zeros(2 * group * n, dtype=np.uint32)
    A = pack_points(curve, group, [a])
    B = pack_points(curve, group, [b])
    K = pack_u32([k], 8)
    lib.ht_ec_op(curve.curve_id, group, op, ptr(A), ptr(B), ptr(K), 8, ptr(out))
    return unpack_points(curve, group, out)[0]


@pytest.mark.parametrize("curve,group", CASES)
def test_group_law(hosttest_lib, curve, group):
    G = groups(curve)[group - 1]
    rng = random.Random(7 + group)
    n = fq_limbs(curve)
    gen = np.zeros(2 * group * n, dtype=np.uint32)
    hosttest_lib.ht_generator(curve.curve_id, group, ptr(gen))
    assert unpack_points(curve, group, gen)[0] == G.gen
    P = G.mul(G.gen, rng.randrange(1, curve.r))
    Q = G.mul(G.gen, rng.randrange(1, curve.r))
    for op in (0, 1):
        assert call(hosttest_lib, curve, group, op, P, Q) == G.add(P, Q)
        assert call(hosttest_lib, curve, group, op, P, P) == G.dbl(P)          # P == Q -> doubling
        assert call(hosttest_lib, curve, group, op, P, G.neg(P)) is None       # P == -Q -> identity
        assert call(hosttest_lib, curve, group, op, None, Q) == Q              # identity + Q
        assert call(hosttest_lib, curve, group, op, P, None) == P              # P + identity
        assert call(hosttest_lib, curve, group, op, None, None) is None
    assert call(hosttest_lib, curve, group, 2, P, Q) == G.dbl(P)
    assert call(hosttest_lib, curve, group, 2, None, Q) is None
    assert call(hosttest_lib, curve, group, 4, P, Q) == G.add(G.mul(P, 3), G.mul(Q, 2))
    assert call(hosttest_lib, curve, group, 4, P, P) == G.mul(P, 5)
    assert call(hosttest_lib, curve, group, 5, P, Q) == P
    assert call(hosttest_lib, curve, group, 6, P, Q) == G.mul(P, 4)
    for k in (0, 1, 2, 3, curve.r - 1, curve.r, rng.randrange(curve.r)):
        assert call(hosttest_lib, curve, group, 3, P, Q, k) == G.mul(P, k), k
