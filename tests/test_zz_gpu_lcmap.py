"""SURVEY 8(f) row 1 on the device: b2s_r1cs_upload_lcmap must give the same handle as b2s_r1cs_upload fed with
to_matrices() -- same SpMV, same proof.

All of it has run green on a B200 (profiles/r02_gpu_job1.txt): the handle, setup + prove from it, the error path and the
C++ host through B2S_HOST_LCMAP=1."""
import os
import random

import numpy as np
import pytest

from oracle import groth16 as og
from oracle import r1cs as orc
from oracle.params import BLS12_381, BN254
from tests.test_host_lcmap import circuits
from tests.util import csr_from_rows, pack_fr, unpack_fr, unpack_points

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


def upload_lcmap(be, curve, cs):
    lm = cs.to_lcmap()
    args = [np.array(a, dtype=np.uint64) for a in lm["args"]]
    return be.r1cs_upload_lcmap(len(args[0]), cs.num_instance_variables, cs.num_witness_variables, args,
                                np.array(lm["offsets"], dtype=np.uint64), np.array(lm["vars"] or [0], dtype=np.uint64),
                                np.array(lm["coeffs"] or [0], dtype=np.uint32), pack_fr(curve, lm["pool"]))


def test_lcmap_spmv_equals_matrix_path(be):
    """CSR built on the device from the LcMap == CSR uploaded from to_matrices(): same SpMV, equal to the oracle's."""
    curve = CURVES[be.curve]
    for name, cs in circuits(curve).items():
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        n_rows = len(mats[0])
        m_ref = be.r1cs_upload(n_rows, len(inst), len(wit), [csr_from_rows(curve, M) for M in mats])
        m_lc = upload_lcmap(be, curve, cs)
        z = pack_fr(curve, inst + wit)
        ref, got = be.spmv(m_ref, z, n_rows), be.spmv(m_lc, z, n_rows)
        for k in range(3):
            assert np.array_equal(ref[k], got[k]), (name, k)
            assert unpack_fr(curve, got[k]) == orc.mat_vec_mul(curve.r, mats[k], inst + wit), (name, k)
        be.r1cs_free(m_ref); be.r1cs_free(m_lc)


def test_lcmap_handle_equals_matrix_handle(be):
    curve = CURVES[be.curve]
    rng = random.Random(0xB2000005)
    for name, cs in circuits(curve).items():
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        n_rows = len(mats[0])
        m_ref = be.r1cs_upload(n_rows, len(inst), len(wit), [csr_from_rows(curve, M) for M in mats])
        m_lc = upload_lcmap(be, curve, cs)
        assert be.domain_size(m_lc) == be.domain_size(m_ref), name
        z = pack_fr(curve, inst + wit)
        ref, got = be.spmv(m_ref, z, n_rows), be.spmv(m_lc, z, n_rows)
        for k in range(3):
            assert np.array_equal(ref[k], got[k]), (name, k)
            assert unpack_fr(curve, got[k]) == orc.mat_vec_mul(curve.r, mats[k], inst + wit), (name, k)
        if cs.is_satisfied():
            td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
            pkh, _ = be.groth16_setup(m_lc, pack_fr(curve, [td.tau, td.alpha, td.beta, td.gamma, td.delta]), len(inst))
            pk = og.setup(curve, mats, len(inst), len(wit), td)
            rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
            A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
            a, b, c = be.groth16_prove(pkh, m_lc, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [rr]), pack_fr(curve, [ss]))
            assert (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b)[0], unpack_points(curve, 1, c)[0]) == (A, B, C), name
            be.pk_free(pkh)
        be.r1cs_free(m_ref); be.r1cs_free(m_lc)


def test_lcmap_errors(be):
    from snark_b200.lib import B2SError

    curve = CURVES[be.curve]
    cs = orc.circuit2(curve, 1, 1, 2)          # not finalized
    with pytest.raises(B2SError) as e:
        upload_lcmap(be, curve, cs)
    assert e.value.code == 16 and "finalize" in str(e.value)


@pytest.mark.parametrize("cid,circuit", [(0, "circuit2"), (1, "dummy")])
def test_cpp_host_proves_through_the_lcmap_path(cid, circuit):
    """The C++ mirror hands its flat LcMap to b2s_r1cs_upload_lcmap (B2S_HOST_LCMAP=1): same proof as the oracle's."""
    import subprocess

    from tests.test_host_relations import _parse_words, build

    curve = CURVES[cid]
    td_vals = [1234567, 31337, 271828, 314159, 161803]
    rr, ss = 99991, 77773
    out = subprocess.run([build(), "gpu", str(cid), circuit] + [str(v) for v in td_vals + [rr, ss]], capture_output=True, text=True,
                         timeout=300, env=dict(os.environ, B2S_HOST_LCMAP="1"))
    assert out.returncode == 0, out.stdout + out.stderr
    got = _parse_words(out.stdout)
    cs = orc.circuit2(curve, 1, 1, 2) if circuit == "circuit2" else orc.dummy_circuit(curve, 3, 5, 16, 16)
    cs.finalize()
    mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
    pk = og.setup(curve, mats, len(inst), len(wit), og.Trapdoor(*td_vals))
    A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
    assert (unpack_points(curve, 1, got["A"])[0], unpack_points(curve, 2, got["B"])[0], unpack_points(curve, 1, got["C"])[0]) == (A, B, C)
