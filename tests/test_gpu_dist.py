"""Distributed witness map (csrc/dntt.cu): the four-step transforms, SpMV and quotient on column
slabs that b2s_groth16_prove_group runs over 2^lg GPUs must give, bit for bit, the h of the single-GPU witness_map.
b2s_witness_map_sim plays all 2^lg ranks on one GPU (exchange = device copies), so the whole index algebra -- bit-field
address maps, per-destination packing, twiddles keyed by global indices, coset scalings, the final slab redistribution --
is checked here without a multi-GPU box; the NCCL exchange itself is covered by test_group_of_two_devices_one_process."""
import random

import numpy as np
import pytest

from oracle import groth16 as og
from oracle import r1cs as orc
from oracle.params import BLS12_381, BN254
from tests.util import csr_from_rows, pack_fr, random_fr_limbs, unpack_fr

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


@pytest.mark.parametrize("n_rows,log_ranks", [(250, 1), (250, 2), (1000, 1), (1000, 2), (4000, 3)])
def test_sim_equals_single_gpu_and_oracle(be, n_rows, log_ranks):
    """DummyCircuit-shaped and BenchCircuit systems whose domain is 2^8 / 2^10 / 2^12, over 2 / 4 / 8 virtual ranks."""
    curve = CURVES[be.curve]
    rng = random.Random(n_rows + log_ranks)
    systems = [orc.dummy_circuit_direct(curve, rng.randrange(curve.r), rng.randrange(curve.r), n_rows, n_rows - 1)]
    if n_rows <= 1000:
        bc = orc.bench_circuit(curve, n_rows - 6, seed=3)
        bc.finalize()
        systems.append((bc.to_matrices(), bc.instance_assignment, bc.witness_assignment))
    for mats, inst, wit in systems:
        m = be.r1cs_upload(len(mats[0]), len(inst), len(wit), [csr_from_rows(curve, M) for M in mats])
        assert be.domain_size(m) in (256, 1024, 4096)
        z = pack_fr(curve, inst + wit)
        ref = be.witness_map(m, z)
        got = be.witness_map_sim(m, z, log_ranks)
        assert np.array_equal(ref, got)
        if n_rows <= 250:
            assert unpack_fr(curve, got) == og.witness_map(curve, mats, inst + wit, len(inst))
        be.r1cs_free(m)


def test_sim_rejects_odd_log_domain(be):
    from snark_b200 import B2SError

    curve = CURVES[be.curve]
    mats, inst, wit = orc.dummy_circuit_direct(curve, 3, 5, 100, 99)       # domain 128 = 2^7
    m = be.r1cs_upload(len(mats[0]), len(inst), len(wit), [csr_from_rows(curve, M) for M in mats])
    with pytest.raises(B2SError) as e:
        be.witness_map_sim(m, pack_fr(curve, inst + wit), 1)
    assert e.value.code == 16
    be.r1cs_free(m)


@pytest.mark.parametrize("log_ranks", [1, 3])
def test_sim_2p20_random_assignment(log_ranks):
    """Domain 2^20 (the smallest of BASELINE's range), BenchCircuit-shaped rows with real gathers, a random (not even
    satisfying) assignment: the two schedules are the same function of z."""
    from snark_b200 import Backend
    from tools.spmv_probe import bench_shaped_csr

    curve = BLS12_381
    be = Backend(curve=0)
    n_rows = (1 << 20) - 1
    mats, n_wit = bench_shaped_csr(n_rows, seed=7)
    one = pack_fr(curve, [1])
    csr = [(rp, col, np.tile(one, len(col))) for rp, col in mats]
    m = be.r1cs_upload(n_rows, 1, n_wit, csr)
    assert be.domain_size(m) == 1 << 20
    z = random_fr_limbs(np.random.default_rng(20), 1 + n_wit, bits=254)
    ref = be.witness_map(m, z)
    got = be.witness_map_sim(m, z, log_ranks)
    assert np.array_equal(ref, got)
    be.r1cs_free(m)
    be.close()
