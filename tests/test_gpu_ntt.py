"""K2 parity: b2s_ntt against the oracle restatement of ark-poly's radix-2 domain,
for every size 2^0..2^12 (covers the 1-, 2-pass schedules), forward / inverse / coset, both curves; and
the size-independent round-trip property at 2^20 (3 passes) and BASELINE config 3's 2^24."""
import random

import numpy as np
import pytest

from oracle import ntt as ontt
from oracle.params import BLS12_381, BN254
from tests.util import pack_fr, random_fr_limbs, unpack_fr

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


@pytest.mark.parametrize("log_n", list(range(0, 13)))
def test_ntt_matches_oracle(be, log_n):
    curve = CURVES[be.curve]
    rng = random.Random(100 + log_n)
    n = 1 << log_n
    x = [rng.randrange(curve.r) for _ in range(n)]
    if n >= 4:
        x[0], x[1], x[2] = 0, 1, curve.r - 1
    if n <= 64:
        assert ontt.ntt(curve, x) == ontt.dft_naive(curve, x)  # oracle self-check against the definition
    assert unpack_fr(curve, be.ntt(pack_fr(curve, x), log_n)) == ontt.ntt(curve, x)
    assert unpack_fr(curve, be.ntt(pack_fr(curve, x), log_n, inverse=True)) == ontt.ntt(curve, x, inverse=True)
    assert unpack_fr(curve, be.ntt(pack_fr(curve, x), log_n, coset=True)) == ontt.coset_ntt(curve, x)
    assert unpack_fr(curve, be.ntt(pack_fr(curve, x), log_n, inverse=True, coset=True)) == ontt.coset_intt(curve, x)


@pytest.mark.parametrize("log_n", [13, 16, 19])
def test_ntt_matches_oracle_multi_pass(be, log_n):
    """2-pass (13, 16) and 3-pass (19) schedules against the oracle NTT."""
    curve = CURVES[be.curve]
    if be.curve == 1 and log_n == 19:
        pytest.skip("one 2^19 oracle transform is enough")
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    raw = random_fr_limbs(rng, n, bits=curve.r.bit_length() - 1)
    x = unpack_fr(curve, raw)
    assert unpack_fr(curve, be.ntt(raw.copy(), log_n)) == ontt.ntt(curve, x)
    if log_n <= 16:
        assert unpack_fr(curve, be.ntt(raw.copy(), log_n, inverse=True, coset=True)) == ontt.coset_intt(curve, x)


@pytest.mark.parametrize("log_n", [20, 24])
def test_ntt_round_trip_large(be, log_n):
    """intt(ntt(x)) == x and coset_intt(coset_ntt(x)) == x bit-exactly (device-resident data)."""
    import torch

    if be.curve == 1 and log_n == 24:
        pytest.skip("2^24 round trip is run on BLS12-381 (BASELINE config 3)")
    n = 1 << log_n
    curve = CURVES[be.curve]
    rng = np.random.default_rng(0xB2000002)
    raw = random_fr_limbs(rng, n, bits=curve.r.bit_length() - 1)
    x = torch.from_numpy(raw.view(np.int32)).cuda()
    y = x.clone()
    be.ntt(y, log_n)
    be.sync()
    assert not torch.equal(x, y)
    # the first 2^12-point sub-problem against the oracle: X[k * n/4096], k < 4096, is the
    # 4096-point NTT of x folded to length 4096 (y[j] = sum_m x[j + 4096 m]) -- 4096 outputs, each a sum over all inputs
    from tests.util import limbs_to_ints

    sub = 4096
    Rinv = pow(1 << 256, -1, curve.r)
    xs = limbs_to_ints(raw)
    folded = [sum(xs[j::sub]) * Rinv % curve.r for j in range(sub)]
    got_sub = y.view(n, 8)[:: n // sub].contiguous().cpu().numpy().view(np.uint32).reshape(-1)
    assert unpack_fr(curve, got_sub) == ontt.ntt(curve, folded)
    be.ntt(y, log_n, inverse=True)
    be.sync()
    assert torch.equal(x, y)
    be.ntt(y, log_n, coset=True)
    be.ntt(y, log_n, inverse=True, coset=True)
    be.sync()
    assert torch.equal(x, y)
    # linearity spot-check against the oracle: NTT(x)[0] = sum x, NTT(x)[n/2] = sum (-1)^j x_j
    be.ntt(y, log_n)
    be.sync()
    s0 = sum(xs) * Rinv % curve.r
    s1 = (sum(xs[0::2]) - sum(xs[1::2])) * Rinv % curve.r
    got = y.cpu().numpy().view(np.uint32)
    assert unpack_fr(curve, got[:8])[0] == s0
    assert unpack_fr(curve, got[(n // 2) * 8 : (n // 2) * 8 + 8])[0] == s1
