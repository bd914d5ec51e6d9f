"""Universal-setup (Marlin-style) path on the GPU: the element-wise polynomial kernels against big-int
arithmetic, and the whole prover through the C ABI against the big-int prover (bit-equal proofs) and the independent verifier."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _curves():
    from oracle.params import BLS12_381, BN254
    return {"bn254": BN254, "bls12_381": BLS12_381}


@pytest.fixture(scope="module", params=["bn254", "bls12_381"])
def backends(request):
    from oracle import marlin as om
    from snark_b200.marlin_gpu import GpuBackend

    curve = _curves()[request.param]
    gb = GpuBackend(curve=curve.curve_id, device=0)
    assert gb.r == curve.r and gb.p == curve.p and gb.coset_gen == curve.fr_generator
    assert gb.omega(10) == curve.omega(10)
    yield curve, gb, om.IntBackend(curve)
    gb.close()


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000, 4099])
def test_poly_kernels_match_bigint(backends, n):
    curve, gb, ib = backends
    rng = random.Random(1000 + n)
    r = curve.r
    a = [rng.randrange(r) for _ in range(n)]
    b = [rng.randrange(r) for _ in range(n)]
    for i in range(0, n, 5):
        a[i] = 0                                  # zeros for the batched inversion
    s = rng.randrange(r)
    da, db = gb.from_ints(a), gb.from_ints(b)
    assert gb.to_ints(da) == a                    # Montgomery round trip through the device
    launches0 = gb.launches
    assert gb.to_ints(gb.mul(da, db)) == ib.mul(a, b)
    assert gb.to_ints(gb.add(da, db)) == ib.add(a, b)
    assert gb.to_ints(gb.sub(da, db)) == ib.sub(a, b)
    assert gb.to_ints(gb.scale(da, s)) == ib.scale(a, s)
    assert gb.to_ints(gb.add_scalar(da, s)) == ib.add_scalar(a, s)
    assert gb.to_ints(gb.inv0(da)) == ib.inv0(a)
    assert gb.to_ints(gb.geom(n, s, b[0])) == ib.geom(n, s, b[0])
    assert gb.eval(db, s) == ib.eval(b, s)
    assert gb.eval(db, 0) == b[0] and gb.eval(db, 1) == sum(b) % r
    assert gb.launches - launches0 >= 10
    # views, padding, shifting (device plumbing)
    if n > 3:
        assert gb.to_ints(gb.slice(da, 1, n - 1)) == a[1:n - 1]
        assert gb.eval(gb.slice(db, 1, n), s) == ib.eval(b[1:], s)
        assert gb.to_ints(gb.pad(db, n + 5)) == b + [0] * 5
        assert gb.to_ints(gb.shifted(db, 3)) == [0] * 3 + b
        assert gb.to_ints(gb.concat([da, db])) == a + b


def test_poly_op_host_buffers_and_errors(backends):
    """The same entry points with HOST buffers (mem = 0), and the argument checks."""
    from snark_b200 import lib as L

    curve, gb, ib = backends
    rng = random.Random(7)
    n, r = 100, curve.r
    R = pow(2, 256, r)
    a = [rng.randrange(1, r) for _ in range(n)]
    b = [rng.randrange(r) for _ in range(n)]
    pack = lambda xs: np.frombuffer(b"".join((x * R % r).to_bytes(32, "little") for x in xs), dtype=np.uint32).copy()
    unpack = lambda arr: [int.from_bytes(arr.tobytes()[32 * i:32 * i + 32], "little") * pow(R, -1, r) % r for i in range(len(arr) // 8)]
    ha, hb, out = pack(a), pack(b), np.zeros(n * 8, dtype=np.uint32)
    for op, want in ((0, ib.mul(a, b)), (2, ib.sub(a, b)), (5, ib.inv0(a))):
        st = gb.lib.b2s_poly_op(gb.h, op, ha.ctypes.data, hb.ctypes.data, None, out.ctypes.data, n, L.MEM_HOST)
        assert st == 0 and unpack(out) == want
    assert gb.lib.b2s_poly_op(gb.h, 9, ha.ctypes.data, hb.ctypes.data, None, out.ctypes.data, n, L.MEM_HOST) != 0
    assert gb.lib.b2s_poly_op(gb.h, 3, ha.ctypes.data, None, None, out.ctypes.data, n, L.MEM_HOST) != 0      # scale without a scalar
    assert gb.lib.b2s_poly_op(gb.h, 0, ha.ctypes.data, None, None, out.ctypes.data, n, L.MEM_HOST) != 0      # product without b
    d = gb.from_ints(a)
    assert gb.lib.b2s_poly_op(gb.h, 5, d.data_ptr(), None, None, d.data_ptr(), n, L.MEM_DEVICE) != 0          # inversion in place


def _cases(curve, big):
    from oracle import r1cs as orc

    css = [orc.circuit2(curve, 1, 1, 2), orc.dummy_circuit(curve, 3, 5, 8, 8), orc.bench_circuit(curve, 9, seed=2)]
    if big:
        css.append(orc.dummy_circuit(curve, 2, 9, 700, 600))
    for cs in css:
        cs.finalize()
        yield cs.to_matrices(), list(cs.instance_assignment), list(cs.witness_assignment)


def test_gpu_prover_equals_bigint_prover_and_verifies(backends):
    from oracle import marlin as om
    from snark_b200 import marlin as M

    curve, gb, ib = backends
    rng = random.Random(0xB2000005)
    for mats, x, w in _cases(curve, big=(curve.name.lower().startswith("bn"))):
        info = M.index_shape(mats, len(x), len(x) + len(w))
        tau = rng.randrange(2, curve.r)
        srs_g, srs_i = gb.setup(info.D + 1, tau), ib.setup(info.D + 1, tau)
        pk_g, vk_g = M.index(gb, srs_g, mats, len(x), len(x) + len(w))
        pk_i, vk_i = M.index(ib, srs_i, mats, len(x), len(x) + len(w))
        assert vk_g.index_comms == vk_i.index_comms and vk_g.info == vk_i.info
        launches0 = gb.launches
        pg = M.prove(gb, pk_g, x, w, check=True)
        assert gb.launches > launches0
        pi = M.prove(ib, pk_i, x, w, check=True)
        assert pg.comms == pi.comms
        assert pg.evals1 == pi.evals1 and pg.evals2 == pi.evals2
        assert pg.openings == pi.openings
        assert om.verify(curve, vk_g, x, pg, tau=tau)
        bad = list(x)
        bad[-1] = (bad[-1] + 1) % curve.r
        if len(x) > 1:
            assert not om.verify(curve, vk_g, bad, pg, tau=tau)


def test_srs_is_the_powers_of_tau(backends):
    """universal_setup: srs[i] = tau^i G1 (spot checks against double-and-add) and a commitment equals p(tau) G1."""
    from oracle.ec import groups

    curve, gb, ib = backends
    G1 = groups(curve)[0]
    tau, size = 0x1234567 + curve.curve_id, 300
    srs = gb.setup(size, tau)
    gb.be.sync()
    host = srs.cpu().numpy().astype(np.uint32)
    for i in (0, 1, 2, 157, 299):
        assert gb._point(host[i]) == G1.mul(G1.gen, pow(tau, i, curve.r))
    rng = random.Random(3)
    coeffs = [rng.randrange(curve.r) for _ in range(200)]
    assert gb.commit(srs, gb.from_ints(coeffs)) == ib.commit(ib.setup(size, tau), coeffs)
    assert gb.commit(srs, gb.from_ints(coeffs), shift=100) == ib.commit(ib.setup(size, tau), coeffs, shift=100)
    assert gb.commit(srs, gb.from_ints([0] * 10)) is None
