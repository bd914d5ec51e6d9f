"""World-size-2 gloo test of the multi-GPU host logic (snark_b200/shard.py): base-range shard plan,
all-gather layout and the join -- with the C++ oracle standing in for the per-rank GPU engine, so the
arithmetic identity "sum of shard MSMs == full MSM" and the message layout are exercised on CPU."""
import os
import random
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snark_b200 import shard


def test_shard_ranges_partition():
    for total in (0, 1, 5, 1000, (1 << 24) - 1):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    plan = shard.shard_plan(2, 13, 16, 1, 2)
    # h is cut by coefficient slab [r N/G, (r+1) N/G) (clipped to the N - 1 query points), the others by base range
    assert plan == {"a_query": (7, 8), "b_g1_query": (7, 8), "b_g2_query": (7, 8), "h_query": (8, 7), "l_query": (6, 7)}
    for world in (1, 2, 4, 8):
        spans = [shard.slab_range(1 << 10, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == (1 << 10) - 1 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo == r * (1 << 10) // world for r, (lo, _) in enumerate(spans))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cnative
        from oracle import msm as omsm
        from oracle.ec import groups
        from oracle.params import BLS12_381 as curve
        from tests.util import pack_fr, pack_points, unpack_points

        G1, G2 = groups(curve)
        rng = random.Random(1234)                     # same data on every rank
        n = 37
        b1 = [G1.mul(G1.gen, rng.randrange(1, curve.r)) for _ in range(n)]
        b2 = [G2.mul(G2.gen, rng.randrange(1, curve.r)) for _ in range(n)]
        s = [rng.randrange(curve.r) for _ in range(n)]
        lo, hi = shard.shard_range(n, rank, world)

        def to_xyzz(group, aff_arr):
            """affine oracle output -> XYZZ with ZZ = ZZZ = 1 (identity -> zeros), the GPU exchange format."""
            w = len(aff_arr) // 2
            one = pack_fr(curve, [0])  # placeholder, replaced below
            from tests.util import pack_u32
            R = 1 << 384
            one = pack_u32([R % curve.p], 12)
            if not aff_arr.any():
                return np.zeros(2 * len(aff_arr), dtype=np.uint32)
            o = one if group == 1 else np.concatenate([one, np.zeros(12, dtype=np.uint32)])
            return np.concatenate([aff_arr, o, o])

        def shard_fn():
            parts = []
            for _ in range(4):   # the four G1 MSMs of a proof; here the same MSM four times
                a = cnative.msm(0, 1, pack_points(curve, 1, b1[lo:hi]), pack_fr(curve, s[lo:hi]), hi - lo, threads=1)
                parts.append(to_xyzz(1, a))
            a2 = cnative.msm(0, 2, pack_points(curve, 2, b2[lo:hi]), pack_fr(curve, s[lo:hi]), hi - lo, threads=1)
            return np.concatenate(parts), to_xyzz(2, a2)

        def finish_fn(p1, p2, w):
            # join on the CPU oracle: decode the Z = 1 points and add them up
            g1w, g2w = 48, 96
            acc1, acc2 = None, None
            for r in range(w):
                pt = unpack_points(curve, 1, p1[(r * 4) * g1w: (r * 4) * g1w + 24])[0]
                acc1 = G1.add(acc1, pt)
                pt2 = unpack_points(curve, 2, p2[r * g2w: r * g2w + 48])[0]
                acc2 = G2.add(acc2, pt2)
            return acc1, acc2

        out = shard.sharded_prove(dist, rank, world, shard_fn, finish_fn, 48, 96)
        if rank == 0:
            q.put(out == (omsm.msm_naive(G1, b1, s), omsm.msm_naive(G2, b2, s)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_window_unit_plan_balances():
    """Round-2 partition by (MSM, window) units: every unit owned exactly once, >= 95 % balance up to 8 ranks."""
    for world in (1, 2, 4, 8):
        owned, loads, balance = shard.plan_window_units(world, 13, 13)
        flat = [u for lst in owned for u in lst]
        assert len(flat) == len(set(flat)) == 4 * 13 + 13
        assert {u for u in flat if u[0] == 2} == {(2, 0, w) for w in range(13)}
        assert balance >= 0.95 and abs(sum(loads) - (52 + 13 * 3.16)) < 1e-9
    assert shard.plan_window_units(8, 16, 16)[2] >= 0.93


def test_two_rank_gloo_shard_and_join():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


# ---- distributed four-step NTT (snark_b200/dist_ntt.py), oracle arithmetic as the stand-in engine ---------------------
class _OracleNttEngine:
    """Python big-int stand-in for the GPU kernels: tiles are object arrays of ints (limbs axis of length 1)."""

    def __init__(self, curve):
        from oracle import ntt as ontt

        self.curve, self.ontt = curve, ontt

    def _w(self, N, inverse):
        w = self.curve.omega(N.bit_length() - 1)
        return pow(w, -1, self.curve.r) if inverse else w

    def _raw(self, vec, inverse):
        """size-n transform without the 1/n of the inverse (the plan applies 1/N once at the end)."""
        out = self.ontt.ntt(self.curve, [int(v) for v in vec], inverse=inverse)
        return [o * len(vec) % self.curve.r for o in out] if inverse else out

    def ntt_axis0(self, a, inverse):
        out = a.copy()
        for j in range(a.shape[1]):
            out[:, j, 0] = self._raw(list(a[:, j, 0]), inverse)
        return out

    def ntt_axis1(self, a, inverse):
        out = a.copy()
        for i in range(a.shape[0]):
            out[i, :, 0] = self._raw(list(a[i, :, 0]), inverse)
        return out

    def mul_pow(self, a, row0, col0, inverse, N):
        w, r = self._w(N, inverse), self.curve.r
        out = a.copy()
        for i in range(a.shape[0]):
            for j in range(a.shape[1]):
                out[i, j, 0] = int(a[i, j, 0]) * pow(w, (row0 + i) * (col0 + j), r) % r
        return out

    def mul_geometric(self, a, idx, use_g_inv, scale_n_inv, N):
        r, g = self.curve.r, self.curve.fr_generator
        out = a.copy()
        n_inv = pow(N, -1, r) if scale_n_inv else 1
        for i in range(a.shape[0]):
            for j in range(a.shape[1]):
                k = int(idx[i, j])
                f = n_inv
                if scale_n_inv and use_g_inv:
                    f = f * pow(g, -k, r) % r
                elif not scale_n_inv:
                    f = pow(g, k, r)
                out[i, j, 0] = int(a[i, j, 0]) * f % r
        return out


def _ntt_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ntt as ontt
        from oracle.params import BN254 as curve
        from snark_b200 import dist_ntt

        class _IntTiles(dist_ntt.FourStepNtt):
            pass

        # ints do not fit the uint32 wire format of all_to_all_tiles: ship them as 8 limbs and rebuild
        def to_limbs(t):
            flat = np.zeros(t.shape[:2] + (8,), dtype=np.uint32)
            for i in range(t.shape[0]):
                for j in range(t.shape[1]):
                    v = int(t[i, j, 0])
                    flat[i, j] = [(v >> (32 * k)) & 0xFFFFFFFF for k in range(8)]
            return flat

        def from_limbs(f):
            out = np.empty(f.shape[:2] + (1,), dtype=object)
            for i in range(f.shape[0]):
                for j in range(f.shape[1]):
                    out[i, j, 0] = sum(int(f[i, j, k]) << (32 * k) for k in range(8))
            return out

        orig = dist_ntt.all_to_all_tiles
        dist_ntt.all_to_all_tiles = lambda d, tiles, w: [from_limbs(x) for x in orig(d, [to_limbs(t) for t in tiles], w)]
        ok = True
        rng = random.Random(99)
        for log_n in (4, 6):
            N = 1 << log_n
            x = [rng.randrange(curve.r) for _ in range(N)]              # same data on every rank
            plan = dist_ntt.FourStepNtt(dist, rank, world, log_n, _OracleNttEngine(curve))
            idx_in = dist_ntt.owned_indices_in(plan.N1, plan.N2, rank, world)
            idx_out = dist_ntt.owned_indices_out(plan.N1, plan.N2, rank, world)
            tile = np.empty(idx_in.shape + (1,), dtype=object)
            for i in range(idx_in.shape[0]):
                for j in range(idx_in.shape[1]):
                    tile[i, j, 0] = x[int(idx_in[i, j])]

            def check(z, full):
                return all(int(z[i, j, 0]) == full[int(idx_out[i, j])] for i in range(z.shape[0]) for j in range(z.shape[1]))

            fwd = plan.transform(tile)
            ok &= check(fwd, ontt.ntt(curve, x))
            ok &= check(plan.transform(tile, inverse=True), ontt.ntt(curve, x, inverse=True))
            ok &= check(plan.transform(tile, coset=True), ontt.coset_ntt(curve, x))
            ok &= check(plan.transform(tile, inverse=True, coset=True), ontt.coset_intt(curve, x))
            # chained without re-gathering: forward then inverse gives x back, in the layout the chain ends in
            back = plan.transform(plan.out_as_in(fwd), inverse=True)
            ok &= check(back, x)
            # the two layouts partition the index space
            both = [torch.zeros(N, dtype=torch.int64) for _ in range(2)]
            both[0][torch.from_numpy(idx_in.reshape(-1))] = 1
            both[1][torch.from_numpy(idx_out.reshape(-1))] = 1
            for t in both:
                dist.all_reduce(t)
                ok &= bool((t == 1).all())
        if rank == 0:
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_four_step_ntt():
    """One all-to-all per transform; forward / inverse / coset forms equal the oracle's full-size transforms on the
    indices each rank ends up owning; transforms chain without re-gathering (N1 == N2)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ntt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True
