"""K4 parity: b2s_msm_g1 / b2s_msm_g2 against the oracle (naive double-and-add definition and the
restated ark-ec Pippenger) on small inputs with every edge case the domain has, and
at scale through a size-independent identity: bases k_i*G with known k_i give
MSM(bases, s) == (sum_i s_i k_i mod r) * G."""
import random

import numpy as np
import pytest

from oracle import msm as omsm
from oracle.ec import groups
from oracle.params import BLS12_381, BN254
from tests.util import limbs_to_ints, pack_fr, pack_points, pack_u32, random_fr_limbs, unpack_points

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


@pytest.fixture(params=[None, 3], ids=["rounds_auto", "rounds_3"])
def affine_rounds(request, monkeypatch):
    """The batched-affine halving rounds (csrc/msm_affine.cuh) switch themselves on only where they pay (large MSMs);
    `rounds_3` forces three rounds so that the small inputs and every edge case below run through them as well."""
    if request.param is not None:
        monkeypatch.setenv("B2S_MSM_AFFINE_ROUNDS", str(request.param))
    return request.param


def gpu_msm(be, curve, group, bases, scalars, mont=True):
    B = pack_points(curve, group, bases)
    S = pack_fr(curve, scalars, mont=mont)
    fn = be.msm_g1 if group == 1 else be.msm_g2
    return unpack_points(curve, group, fn(B, S, len(bases), mont=mont))[0]


@pytest.mark.parametrize("group", [1, 2])
def test_msm_small_vs_oracle(be, group, affine_rounds):
    curve = CURVES[be.curve]
    G = groups(curve)[group - 1]
    rng = random.Random(11 * group)
    pts = [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(24)]
    for n in (1, 2, 3, 7, 24, 100, 300):
        bases = [pts[rng.randrange(len(pts))] for _ in range(n)]          # repeats -> doubling path in buckets
        scalars = [rng.randrange(curve.r) for _ in range(n)]
        exp = omsm.msm_naive(G, bases, scalars)
        assert omsm.msm_pippenger(G, bases, scalars) == exp               # oracle self-consistency
        assert gpu_msm(be, curve, group, bases, scalars) == exp
        assert gpu_msm(be, curve, group, bases, scalars, mont=False) == exp


@pytest.mark.parametrize("group", [1, 2])
def test_msm_edge_cases(be, group, affine_rounds):
    curve = CURVES[be.curve]
    G = groups(curve)[group - 1]
    rng = random.Random(3)
    P = [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(8)]
    r = curve.r
    fn = be.msm_g1 if group == 1 else be.msm_g2
    # empty input -> identity
    z = np.zeros(8, dtype=np.uint32)
    assert unpack_points(curve, group, fn(z, z, 0))[0] is None
    cases = [
        ([P[0]], [0]),                                  # zero scalar
        ([None], [5]),                                  # base at infinity
        ([P[0], P[0]], [1, r - 1]),                     # cancels to identity
        ([P[0], G.neg(P[0])], [7, 7]),                  # P and -P in the same bucket
        ([P[1]] * 50, [3] * 50),                        # all-equal scalars and bases (DummyCircuit-like)
        (P, [1] * 8),                                   # unit scalars
        (P, [r - 1] * 8),                               # all -1
        (P, [(1 << 255) % r, (1 << 254), (1 << 16) - 1, 1 << 15, (1 << 15) + 1, (1 << 16), r - 2, 2]),  # window edges
        ([P[2], None, P[3], None], [4, 5, 0, 0]),
    ]
    for bases, scalars in cases:
        assert gpu_msm(be, curve, group, bases, scalars) == omsm.msm_naive(G, bases, scalars), (bases, scalars)


@pytest.mark.parametrize("group", [1, 2])
def test_msm_repeated_scalars(be, group, monkeypatch):
    """The multiplicity-aware front end (csrc/msm.cu: sample -> candidates -> heavy lists + rest) on inputs small enough for
    the oracle: all-equal scalars (the reference's DummyCircuit witness, sr1cs/mod.rs:306-309), a few heavy values mixed with
    zeros and random scalars, heavy values 1 and r - 1, Montgomery and canonical scalar representations."""
    monkeypatch.setenv("B2S_MSM_DEDUP_MIN", "1")
    curve = CURVES[be.curve]
    G = groups(curve)[group - 1]
    r = curve.r
    rng = random.Random(41 + group)
    pool = [G.mul(G.gen, rng.randrange(1, r)) for _ in range(16)] + [None]
    n = 160
    bases = [pool[rng.randrange(len(pool))] for _ in range(n)]
    v1, v2 = rng.randrange(r), rng.randrange(r)
    mixes = [
        [v1] * n,
        [v1 if i % 10 < 7 else (v2 if i % 10 < 9 else rng.randrange(r)) for i in range(n)],
        [0 if i % 3 == 0 else (1 if i % 3 == 1 else r - 1) for i in range(n)],
        [v1 if i % 2 else 0 for i in range(n)],
    ]
    for scalars in mixes:
        exp = omsm.msm_pippenger(G, bases, scalars)
        assert gpu_msm(be, curve, group, bases, scalars) == exp
        assert gpu_msm(be, curve, group, bases, scalars, mont=False) == exp
    monkeypatch.setenv("B2S_MSM_AFFINE_ROUNDS", "2")          # the heavy lists through the batched-affine rounds as well
    assert gpu_msm(be, curve, group, bases, mixes[1]) == omsm.msm_pippenger(G, bases, mixes[1])


def test_msm_window_sizes(be, monkeypatch):
    """The result must not depend on the window size c or the task length L."""
    curve = CURVES[be.curve]
    G = groups(curve)[0]
    rng = random.Random(77)
    bases = [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(40)]
    scalars = [rng.randrange(curve.r) for _ in range(40)]
    exp = omsm.msm_naive(G, bases, scalars)
    for c, L in ((2, 1), (3, 2), (5, 64), (8, 3), (11, 64), (13, 1), (16, 64), (17, 5)):
        monkeypatch.setenv("B2S_MSM_C", str(c))
        monkeypatch.setenv("B2S_MSM_L", str(L))
        assert gpu_msm(be, curve, 1, bases, scalars) == exp, (c, L)


@pytest.mark.parametrize("group,log_n", [(1, 14), (1, 18), (2, 14)])
def test_msm_known_discrete_logs(be, group, log_n, affine_rounds):
    """Random scalars; bases (i+1)*G built on the GPU by the fixed-base kernel."""
    import torch

    curve = CURVES[be.curve]
    G = groups(curve)[group - 1]
    n = 1 << log_n
    ks = np.zeros((n, 8), dtype=np.uint32)
    ks[:, 0] = np.arange(1, n + 1, dtype=np.uint32)
    ks_t = torch.from_numpy(ks.view(np.int32)).cuda()
    pt_bytes = be.g1_bytes if group == 1 else be.g2_bytes
    bases = torch.empty(n * pt_bytes // 4, dtype=torch.int32, device="cuda")
    be.fixed_base(group, ks_t, n, mont=False, out=bases)
    rng = np.random.default_rng(0xB2000001)
    raw = random_fr_limbs(rng, n, bits=curve.r.bit_length() - 1)
    s_t = torch.from_numpy(raw.view(np.int32)).cuda()
    fn = be.msm_g1 if group == 1 else be.msm_g2
    got = unpack_points(curve, group, fn(bases, s_t, n, mont=True))[0]
    Rinv = pow(1 << 256, -1, curve.r)
    total = sum(s * (i + 1) for i, s in enumerate(limbs_to_ints(raw))) * Rinv % curve.r
    assert got == G.mul(G.gen, total)
    # degenerate distribution: every scalar equal (one bucket per window holds all points)
    same = np.tile(raw[:8], n)
    s_t = torch.from_numpy(same.view(np.int32)).cuda()
    got = unpack_points(curve, group, fn(bases, s_t, n, mont=True))[0]
    s0 = limbs_to_ints(raw[:8])[0] * Rinv % curve.r
    assert got == G.mul(G.gen, s0 * (n * (n + 1) // 2) % curve.r)


def test_msm_partial_and_sum(be):
    """Shard form: two base-range shards' XYZZ partials joined by b2s_g1_sum equal the full MSM."""
    curve = CURVES[be.curve]
    G = groups(curve)[0]
    rng = random.Random(21)
    bases = [G.mul(G.gen, rng.randrange(1, curve.r)) for _ in range(30)]
    scalars = [rng.randrange(curve.r) for _ in range(30)]
    B, S = pack_points(curve, 1, bases), pack_fr(curve, scalars)
    n1 = 13
    pb = be.g1_bytes // 4
    p1 = be.msm_g1_partial(B[: n1 * pb], S[: n1 * 8], n1)
    p2 = be.msm_g1_partial(B[n1 * pb :], S[n1 * 8 :], 30 - n1)
    got = unpack_points(curve, 1, be.g1_sum(np.concatenate([p1, p2]), 2))[0]
    assert got == omsm.msm_naive(G, bases, scalars)
