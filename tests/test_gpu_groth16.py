"""K1 / K3 / composition parity: SpMV, witness_map and the full Groth16 prove through the C ABI
against the oracle, anchored on the reference's golden R1CS (circuit2.rs:21-43) and DummyCircuit
(sr1cs/mod.rs:296-317); proofs are additionally checked in the exponent under the known trapdoor."""
import random

import numpy as np
import pytest

from oracle import groth16 as og
from oracle import r1cs as orc
from oracle.params import BLS12_381, BN254
from tests.util import csr_from_rows, make_pk_desc, pack_fr, unpack_fr, unpack_points

pytestmark = pytest.mark.gpu
CURVES = [BLS12_381, BN254]


@pytest.fixture(scope="module", params=[0, 1], ids=["bls12_381", "bn254"])
def be(request):
    from snark_b200 import Backend

    b = Backend(curve=request.param)
    yield b
    b.close()


def upload(be, curve, mats, n_inst, n_wit):
    csr = [csr_from_rows(curve, m) for m in mats]
    return be.r1cs_upload(len(mats[0]), n_inst, n_wit, csr), csr


def circuits(curve):
    rng = random.Random(17)
    cs2 = orc.circuit2(curve, 1, 1, 2)
    cs2.finalize()
    assert cs2.to_matrices() == orc.CIRCUIT2_GOLDEN and cs2.is_satisfied()
    yield "circuit2", cs2.to_matrices(), cs2.instance_assignment, cs2.witness_assignment
    d = orc.dummy_circuit(curve, 3, 5, 16, 16)
    yield "dummy16", d.to_matrices(), d.instance_assignment, d.witness_assignment
    a, b = rng.randrange(curve.r), rng.randrange(curve.r)
    mats, inst, wit = orc.dummy_circuit_direct(curve, a, b, 40, 37)
    yield "dummy_direct", mats, inst, wit
    bc = orc.bench_circuit(curve, 25, seed=5)
    bc.finalize()
    assert bc.is_satisfied()
    yield "bench25", bc.to_matrices(), bc.instance_assignment, bc.witness_assignment


def test_spmv_and_witness_map(be):
    curve = CURVES[be.curve]
    for name, mats, inst, wit in circuits(curve):
        m, _keep = upload(be, curve, mats, len(inst), len(wit))
        z = inst + wit
        outs = be.spmv(m, pack_fr(curve, z), len(mats[0]))
        for k in range(3):
            exp = orc.mat_vec_mul(curve.r, mats[k], z)
            assert unpack_fr(curve, outs[k]) == exp, (name, k)
            assert exp == [orc.evaluate_constraint(curve.r, row, z) for row in mats[k]]
        h = unpack_fr(curve, be.witness_map(m, pack_fr(curve, z)))
        assert h == og.witness_map(curve, mats, z, len(inst)), name
        assert be.domain_size(m) == og.domain_size(len(mats[0]), len(inst))
        be.r1cs_free(m)


def test_spmv_duplicate_and_unsorted_columns(be):
    """Rows with repeated / unsorted columns must simply sum."""
    curve = CURVES[be.curve]
    r = curve.r
    A = [[(5, 2), (7, 1), (r - 1, 2), (1, 0)], [], [(1, 3), (1, 3), (1, 3)]]
    B = [[(1, 0)], [(2, 1)], []]
    C = [[], [(r - 2, 3)], [(1, 1)]]
    inst, wit = [1, 11], [13, 17]
    m, _keep = upload(be, curve, [A, B, C], 2, 2)
    outs = be.spmv(m, pack_fr(curve, inst + wit), 3)
    for k, M in enumerate((A, B, C)):
        assert unpack_fr(curve, outs[k]) == orc.mat_vec_mul(r, M, inst + wit)
    be.r1cs_free(m)


def test_r1cs_upload_rejects_bad_columns(be):
    from snark_b200 import B2SError

    curve = CURVES[be.curve]
    with pytest.raises(B2SError) as e:
        upload(be, curve, [[[(1, 9)]], [[]], [[]]], 1, 1)
    assert e.value.code == 2  # AssignmentMissing: column beyond the assignment


def test_groth16_prove_matches_oracle(be):
    curve = CURVES[be.curve]
    rng = random.Random(0xB2000003)
    for name, mats, inst, wit in circuits(curve):
        td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
        pk = og.setup(curve, mats, len(inst), len(wit), td)
        rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
        A, B, C, h = og.prove(pk, mats, inst, wit, rr, ss)
        assert og.check_in_exponent(pk, (A, B, C), inst, wit, h, rr, ss)
        assert og.verify_equation_in_exponent(pk, inst, *og.expected_proof_exponents(pk, inst, wit, h, rr, ss))
        m, _keep = upload(be, curve, mats, len(inst), len(wit))
        keep = []
        pkh = be.pk_upload(make_pk_desc(curve, pk, keep))
        a, b, c = be.groth16_prove(pkh, m, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [rr]), pack_fr(curve, [ss]))
        got = (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b)[0], unpack_points(curve, 1, c)[0])
        assert got == (A, B, C), name
        # r = s = 0 (deterministic proof) also agrees
        A0, B0, C0, _ = og.prove(pk, mats, inst, wit, 0, 0)
        a, b, c = be.groth16_prove(pkh, m, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [0]), pack_fr(curve, [0]))
        assert (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b)[0], unpack_points(curve, 1, c)[0]) == (A0, B0, C0)
        be.pk_free(pkh)
        be.r1cs_free(m)


def test_groth16_prove_with_h_query_table(be, monkeypatch):
    """Fixed-base window table of the h query (csrc/msm.cu msm_precompute: 2^(c w) P_i for every window, all windows sharing
    one bucket set) forced on for small keys: same proofs as the oracle, through the affine rounds as well."""
    monkeypatch.setenv("B2S_PK_PRECOMP_MIN", "1")
    monkeypatch.setenv("B2S_MSM_PRE_C", "7")          # 37 windows of 64 buckets instead of 13 of 2^19: small-key sized
    curve = CURVES[be.curve]
    rng = random.Random(23)
    for rounds in (None, "2"):
        if rounds:
            monkeypatch.setenv("B2S_MSM_AFFINE_ROUNDS", rounds)
        for name, mats, inst, wit in circuits(curve):
            td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
            pk = og.setup(curve, mats, len(inst), len(wit), td)
            rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
            A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
            m, _keep = upload(be, curve, mats, len(inst), len(wit))
            keep = []
            pkh = be.pk_upload(make_pk_desc(curve, pk, keep))
            a, b, c = be.groth16_prove(pkh, m, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [rr]), pack_fr(curve, [ss]))
            assert (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b)[0], unpack_points(curve, 1, c)[0]) == (A, B, C), (name, rounds)
            be.pk_free(pkh)
            be.r1cs_free(m)


def test_groth16_shards_join(be):
    """Two base-range shards + b2s_groth16_finish give the same proof as the single-GPU call."""
    from snark_b200.lib import PkDesc

    curve = CURVES[be.curve]
    rng = random.Random(5)
    mats, inst, wit = orc.dummy_circuit_direct(curve, rng.randrange(curve.r), rng.randrange(curve.r), 20, 20)
    td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
    pk = og.setup(curve, mats, len(inst), len(wit), td)
    rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
    A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
    m, _keep = upload(be, curve, mats, len(inst), len(wit))
    keep = []
    full = make_pk_desc(curve, pk, keep)
    g1b, g2b = be.g1_bytes, be.g2_bytes

    def shard(idx, nshards):
        d = PkDesc()
        for f in ("n_instance", "n_witness", "domain_size", "alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
            setattr(d, f, getattr(full, f))
        for q, off, ln, sz in (("a_query", "a_off", "a_len", g1b), ("b_g1_query", "b1_off", "b1_len", g1b),
                               ("b_g2_query", "b2_off", "b2_len", g2b), ("h_query", "h_off", "h_len", g1b),
                               ("l_query", "l_off", "l_len", g1b)):
            total = getattr(full, ln)
            lo, hi = total * idx // nshards, total * (idx + 1) // nshards
            setattr(d, q, getattr(full, q) + lo * sz)
            setattr(d, off, lo)
            setattr(d, ln, hi - lo)
        return d

    parts1, parts2 = [], []
    handles = []
    for i in range(2):
        h = be.pk_upload(shard(i, 2))
        handles.append(h)
        g1, g2 = be.groth16_prove_shard(h, m, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [rr]), pack_fr(curve, [ss]))
        parts1.append(g1)
        parts2.append(g2)
    a, b, c = be.groth16_finish(handles[0], np.concatenate(parts1), np.concatenate(parts2), 2, pack_fr(curve, [rr]), pack_fr(curve, [ss]))
    assert (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b)[0], unpack_points(curve, 1, c)[0]) == (A, B, C)
    for h in handles:
        be.pk_free(h)
    be.r1cs_free(m)


def _shard_desc(be, full, idx, nshards):
    from snark_b200.lib import PkDesc

    g1b, g2b = be.g1_bytes, be.g2_bytes
    d = PkDesc()
    for f in ("n_instance", "n_witness", "domain_size", "alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
        setattr(d, f, getattr(full, f))
    for q, off, ln, sz in (("a_query", "a_off", "a_len", g1b), ("b_g1_query", "b1_off", "b1_len", g1b), ("b_g2_query", "b2_off", "b2_len", g2b),
                           ("h_query", "h_off", "h_len", g1b), ("l_query", "l_off", "l_len", g1b)):
        total = getattr(full, ln)
        lo, hi = total * idx // nshards, total * (idx + 1) // nshards
        if q == "h_query":     # coefficient slabs: what the distributed witness map of the group hands each rank
            per = full.domain_size // nshards
            lo, hi = per * idx, min(per * (idx + 1), full.domain_size - 1)
        setattr(d, q, getattr(full, q) + lo * sz)
        setattr(d, off, lo)
        setattr(d, ln, hi - lo)
    return d


def test_group_of_one_equals_single_prove(be):
    """b2s_groth16_prove_group with world = 1 (no NCCL involved): the packed all-gather layout and the strided join give
    the single-GPU proof."""
    curve = CURVES[be.curve]
    rng = random.Random(15)
    mats, inst, wit = orc.dummy_circuit_direct(curve, rng.randrange(curve.r), rng.randrange(curve.r), 24, 24)
    td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
    pk = og.setup(curve, mats, len(inst), len(wit), td)
    rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
    A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
    m, _keep = upload(be, curve, mats, len(inst), len(wit))
    keep = []
    pkh = be.pk_upload(make_pk_desc(curve, pk, keep))
    grp = be.group_create(None, 0, 1)
    a, b, c = be.groth16_prove_group(grp, pkh, m, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [rr]), pack_fr(curve, [ss]))
    assert (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b)[0], unpack_points(curve, 1, c)[0]) == (A, B, C)
    be.group_destroy(grp)
    be.pk_free(pkh)
    be.r1cs_free(m)


def test_group_of_two_devices_one_process():
    """Two ctxs on two GPUs of one process (one host thread each), joined by the library's own NCCL communicator: the
    proof equals the oracle's.  Also covers a second ctx on another device needing the > 48 KiB shared-memory attributes
    of the NTT / MSM kernels (they are per device).  Skipped on a one-GPU box."""
    import threading

    import torch

    from snark_b200 import Backend

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    curve = CURVES[0]
    rng = random.Random(25)
    mats, inst, wit = orc.dummy_circuit_direct(curve, rng.randrange(curve.r), rng.randrange(curve.r), 40, 40)
    td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
    pk = og.setup(curve, mats, len(inst), len(wit), td)
    rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
    A, B, C, _ = og.prove(pk, mats, inst, wit, rr, ss)
    uid = Backend.group_unique_id()
    out, errs = {}, []

    def rank_main(rk):
        try:
            b = Backend(curve=0, device=rk)
            m, _k = upload(b, curve, mats, len(inst), len(wit))
            keep = []
            pkh = b.pk_upload(_shard_desc(b, make_pk_desc(curve, pk, keep), rk, 2))
            grp = b.group_create(uid, rk, 2)
            res = b.groth16_prove_group(grp, pkh, m, pack_fr(curve, inst), pack_fr(curve, wit), pack_fr(curve, [rr]), pack_fr(curve, [ss]))
            if rk == 0:
                out["proof"] = res
            b.group_destroy(grp); b.pk_free(pkh); b.r1cs_free(m); b.close()
        except Exception as e:   # noqa: BLE001 -- reported by the main thread
            errs.append((rk, repr(e)))

    ts = [threading.Thread(target=rank_main, args=(rk,)) for rk in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    a, b_, c = out["proof"]
    assert (unpack_points(curve, 1, a)[0], unpack_points(curve, 2, b_)[0], unpack_points(curve, 1, c)[0]) == (A, B, C)
