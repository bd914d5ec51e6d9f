"""Upper sizes of BASELINE.json's range (2^20-2^26) and the error behaviour of the C ABI
(status codes mirror SynthesisError, relations/src/utils/error.rs:5-21)."""
import ctypes

import numpy as np
import pytest

from oracle.ec import groups
from oracle.params import BLS12_381 as curve
from tests.util import limbs_to_ints, pack_fr, unpack_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from snark_b200 import Backend

    b = Backend(curve=0)
    yield b
    b.close()


def rnd(torch, gen, n, dev):
    t = torch.randint(-(1 << 31), (1 << 31) - 1, (n, 8), dtype=torch.int32, device=dev, generator=gen)
    t[:, 7] &= 0x1FFFFFFF
    return t


def test_ntt_2p26_round_trip(be):
    import torch

    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(26)
    x = rnd(torch, gen, 1 << 26, dev)
    y = x.clone()
    be.ntt(y, 26)
    be.ntt(y, 26, inverse=True)
    be.sync()
    assert torch.equal(x, y)
    be.ntt(y, 26, coset=True)
    be.ntt(y, 26, inverse=True, coset=True)
    be.sync()
    assert torch.equal(x, y)


def test_msm_2p26_known_discrete_logs(be):
    """2^26 bases (i+1)*G... too slow to build one by one on the CPU, fine on the GPU; scalars small-weight so the
    expected sum is cheap: s_i = 1 for i in a sparse set, 0 elsewhere, plus full-width scalars on the first 2^16."""
    import torch

    dev = torch.device("cuda", 0)
    n = 1 << 26
    G1 = groups(curve)[0]
    ks = torch.zeros((n, 8), dtype=torch.int32, device=dev)
    ks[:, 0] = torch.arange(1, n + 1, dtype=torch.int64, device=dev).to(torch.int32)   # k_i = i + 1  (< 2^31)
    bases = torch.empty(n * be.g1_bytes // 4, dtype=torch.int32, device=dev)
    be.fixed_base(1, ks, n, mont=False, out=bases)
    be.sync()          # the library runs on its own stream: inputs must outlive the call until sync
    del ks
    gen = torch.Generator(device=dev); gen.manual_seed(7)
    s = torch.zeros((n, 8), dtype=torch.int32, device=dev)
    head = rnd(torch, gen, 1 << 16, dev)
    s[: 1 << 16] = head
    s[(1 << 16)::4097, 0] = 1                 # canonical 1 at a sparse set of positions
    got = unpack_points(curve, 1, be.msm_g1(bases, s, n, mont=False))[0]
    head_i = limbs_to_ints(head.cpu().numpy().view(np.uint32).reshape(-1))
    total = sum(v * (i + 1) for i, v in enumerate(head_i))
    total += sum(i + 1 for i in range(1 << 16, n, 4097))
    assert got == G1.mul(G1.gen, total % curve.r)


def test_error_codes(be):
    from snark_b200 import B2SError

    lib, h = be.lib, be.h
    buf = np.zeros(64, dtype=np.uint32)
    assert lib.b2s_ntt(h, None, 4, 0, 0, 0) == 16                       # null data -> INVALID_ARG
    assert lib.b2s_ntt(h, buf.ctypes.data, 28, 0, 0, 0) == 5             # beyond the backend limit -> PolynomialDegreeTooLarge
    assert b"exceeds" in lib.b2s_last_error(h)
    assert lib.b2s_msm_g1(h, None, None, 3, 1, 0, buf.ctypes.data) == 16
    assert lib.b2s_msm_g1(h, buf.ctypes.data, buf.ctypes.data, 1, 1, 0, None) == 16
    assert lib.b2s_spmv(h, None, buf.ctypes.data, 0, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data) == 1   # MissingCS
    assert lib.b2s_groth16_prove(h, None, None, *([buf.ctypes.data] * 7)) == 1
    assert lib.b2s_g1_sum(h, buf.ctypes.data, 0, buf.ctypes.data) == 16
    # n_instance == 0 is rejected (the constant One is always variable 0, constraint_system.rs:121)
    rp = np.zeros(2, dtype=np.uint64)
    empty = np.zeros(0, dtype=np.uint32)
    with pytest.raises(B2SError) as e:
        be.r1cs_upload(1, 0, 1, [(rp, empty, empty)] * 3)
    assert e.value.code == 16
    # key / matrices mismatch -> AssignmentMissing
    from snark_b200.lib import PkDesc

    m = be.r1cs_upload(1, 1, 1, [(rp, empty, empty)] * 3)
    g1 = np.zeros(be.g1_bytes // 4 * 4, dtype=np.uint32)
    g2 = np.zeros(be.g2_bytes // 4 * 4, dtype=np.uint32)
    d = PkDesc()
    d.n_instance, d.n_witness, d.domain_size = 1, 5, 8
    for f in ("alpha_g1", "beta_g1", "delta_g1", "a_query", "b_g1_query", "h_query", "l_query"):
        setattr(d, f, g1.ctypes.data)
    d.beta_g2 = d.delta_g2 = d.b_g2_query = g2.ctypes.data
    pk = be.pk_upload(d)
    one = pack_fr(curve, [1])
    with pytest.raises(B2SError) as e:
        be.groth16_prove(pk, m, one, one, one, one)
    assert e.value.code == 7                                             # not a full key -> MalformedVerifyingKey
    d.a_len = d.b1_len = d.b2_len = 4   # exceeds n_vars? no: 6 variables; ranges beyond the key are rejected at upload
    d.a_off = 5
    with pytest.raises(B2SError) as e:
        be.pk_upload(d)
    assert e.value.code == 7
    be.pk_free(pk); be.r1cs_free(m)
    # the context still works after errors
    x = pack_fr(curve, [1, 2, 3, 4])
    from oracle import ntt as ontt
    from tests.util import unpack_fr

    assert unpack_fr(curve, be.ntt(x.copy(), 2)) == ontt.ntt(curve, [1, 2, 3, 4])
