"""World-size-2 gloo test of the multi-GPU host logic (snark_b200/shard.py): base-range shard plan,
all-gather layout and the join -- with the C++ oracle standing in for the per-rank GPU engine, so the
arithmetic identity "sum of shard MSMs == full MSM" and the message layout are exercised on CPU."""
import os
import random
import socket

import numpy as np
import pytest
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
    assert plan == {"a_query": (7, 8), "b_g1_query": (7, 8), "b_g2_query": (7, 8), "h_query": (7, 8), "l_query": (6, 7)}


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
