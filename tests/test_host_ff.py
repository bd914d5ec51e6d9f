"""The kernels' field templates (snark_b200/csrc/ff.cuh), compiled for the host with the PTX carry
flag emulated, against Python big-int arithmetic.  Runs without a GPU; the same cases run on the
device in tests/test_gpu_field.py."""
import random

import numpy as np
import pytest

from oracle.params import BLS12_381 as BLS, BN254 as BN
from tests.util import pack_u32, ptr, unpack_u32

FIELDS = [("bls_fq", BLS.p, 12), ("bls_fr", BLS.r, 8), ("bn_fq", BN.p, 8), ("bn_fr", BN.r, 8)]


def run(lib, f, op, a, b, n):
    out = np.zeros(len(a) * n, dtype=np.uint32)
    A, B = pack_u32(a, n), pack_u32(b, n)
    lib.ht_field_op(f, op, ptr(A), ptr(B), ptr(out), len(a))
    return unpack_u32(out, n)


@pytest.mark.parametrize("f", range(4))
def test_field_ops_match_bigint(hosttest_lib, f):
    name, p, n = FIELDS[f]
    R = 1 << (32 * n)
    Rinv = pow(R, -1, p)
    rng = random.Random(1000 + f)
    edge = [0, 1, 2, p - 1, p - 2, R % p, R * R % p, (p - 1) // 2, (1 << (32 * n - 32)) % p]
    a = edge + [rng.randrange(p) for _ in range(3000)]
    b = [rng.randrange(p) for _ in range(3000)] + edge
    cases = {
        0: lambda x, y: x * y * Rinv % p,
        1: lambda x, y: (x + y) % p,
        2: lambda x, y: (x - y) % p,
        4: lambda x, y: (-x) % p,
        5: lambda x, y: x * R % p,
        6: lambda x, y: x * Rinv % p,
        7: lambda x, y: x * x * Rinv % p,
    }
    for op, fn in cases.items():
        assert run(hosttest_lib, f, op, a, b, n) == [fn(x, y) for x, y in zip(a, b)], (name, op)
    # edge x edge products
    ea = [x for x in edge for _ in edge]
    eb = [y for _ in edge for y in edge]
    assert run(hosttest_lib, f, 0, ea, eb, n) == [x * y * Rinv % p for x, y in zip(ea, eb)]
    inv = run(hosttest_lib, f, 3, a[:40], b[:40], n)
    assert inv == [(pow(x * Rinv % p, -1, p) * R % p if x else 0) for x in a[:40]]


@pytest.mark.parametrize("f", range(4))
def test_generated_constants(hosttest_lib, f):
    name, p, n = FIELDS[f]
    R = 1 << (32 * n)
    out = np.zeros(n, dtype=np.uint32)
    exp = [p, R % p, R * R % p]
    if f in (1, 3):
        cur = BLS if f == 1 else BN
        exp += [cur.fr_generator * R % p, cur.fr_root_of_unity * R % p]
    for which, e in enumerate(exp):
        hosttest_lib.ht_field_const(f, which, ptr(out))
        assert unpack_u32(out, n)[0] == e, (name, which)
