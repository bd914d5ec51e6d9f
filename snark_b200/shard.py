"""Host-side logic of the multi-GPU path: the five MSMs of a proof are cut by
base-point range, one shard per rank; the only exchange is an all-gather of five partial sums
(EC addition is not an NCCL reduction op, so no all-reduce), after which rank 0 joins them and applies
the r/s epilogue.  `torch.distributed` does the plumbing (NCCL on GPUs; the same code runs over gloo
in tests/test_multirank_cpu.py with a stand-in engine).
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous, disjoint, exhaustive split of [0, total) -- rank r owns [lo, hi)."""
    return total * rank // world, total * (rank + 1) // world


def slab_range(domain_size, rank, world):
    """h-query shard = the coefficient slab [rank N/G, (rank+1) N/G) of h (clipped to the N - 1 query points): the layout
    the distributed witness map (csrc/dntt.cu) delivers, so no redistribution of h is needed before its MSM."""
    per = domain_size // world
    lo = per * rank
    return lo, min(lo + per, domain_size - 1)


QUERIES = (  # (descriptor field, offset field, length field, group, which length)
    ("a_query", "a_off", "a_len", 1, "n_vars"),
    ("b_g1_query", "b1_off", "b1_len", 1, "n_vars"),
    ("b_g2_query", "b2_off", "b2_len", 2, "n_vars"),
    ("h_query", "h_off", "h_len", 1, "h"),
    ("l_query", "l_off", "l_len", 1, "n_wit"),
)


def query_totals(n_instance, n_witness, domain_size):
    return {"n_vars": n_instance + n_witness, "h": domain_size - 1, "n_wit": n_witness}


def shard_plan(n_instance, n_witness, domain_size, rank, world):
    """{query name: (offset, length)} held by `rank`."""
    tot = query_totals(n_instance, n_witness, domain_size)
    plan = {}
    for name, _, _, _, which in QUERIES:
        lo, hi = slab_range(domain_size, rank, world) if which == "h" else shard_range(tot[which], rank, world)
        plan[name] = (lo, hi - lo)
    return plan


def plan_window_units(world, nwin_g1, nwin_g2, g2_cost=3.16, n_g1_msms=4):
    """Alternative partition for round 2 (not wired to kernels yet): instead of cutting every MSM by base range --
    which leaves each rank a smaller Pippenger problem (more windows, batched-affine rounds off) -- hand out whole
    (MSM, window) units over the FULL base range; every rank then works at the single-GPU operating point and only the
    replicated key (9 GiB at 2^24) is the price.  Longest-processing-time greedy; a G2 window costs `g2_cost` G1
    windows (357 ms vs 113 ms per 2^24-point MSM).
    Returns (units per rank: list of (group, msm index, window), load per rank, balance = mean load / max load)."""
    import heapq

    units = [(1.0, (1, m, w)) for m in range(n_g1_msms) for w in range(nwin_g1)] + [(float(g2_cost), (2, 0, w)) for w in range(nwin_g2)]
    heap = [(0.0, r) for r in range(world)]
    heapq.heapify(heap)
    owned = [[] for _ in range(world)]
    for cost, unit in sorted(units, key=lambda cu: (-cu[0], cu[1])):
        load, r = heapq.heappop(heap)
        owned[r].append(unit)
        heapq.heappush(heap, (load + cost, r))
    loads = [0.0] * world
    for load, r in heap:
        loads[r] = load
    return owned, loads, (sum(loads) / world) / max(loads)


def pack_partials(g1_partials, g2_partial):
    """One rank's contribution to the all-gather: 4 G1 XYZZ + 1 G2 XYZZ as one flat uint32 vector."""
    return np.concatenate([np.asarray(g1_partials, dtype=np.uint32).reshape(-1), np.asarray(g2_partial, dtype=np.uint32).reshape(-1)])


def unpack_gathered(flat, world, g1_words, g2_words):
    """Gathered vector (rank-major) -> (g1 partials [world*4*g1_words], g2 partials [world*g2_words])."""
    per = 4 * g1_words + g2_words
    m = np.asarray(flat, dtype=np.uint32).reshape(world, per)
    return np.ascontiguousarray(m[:, : 4 * g1_words]).reshape(-1), np.ascontiguousarray(m[:, 4 * g1_words:]).reshape(-1)


def all_gather_partials(dist, mine, world, device=None):
    """All-gather one flat uint32 vector per rank (int32 view for NCCL), returns numpy on every rank."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(mine).view(np.int32))
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.int32, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(np.uint32)


def sharded_prove(dist, rank, world, shard_fn, finish_fn, g1_words, g2_words, device=None):
    """One proof over `world` ranks.  shard_fn() -> (g1 partials, g2 partial) for this rank's key shard;
    finish_fn(g1 parts, g2 parts, world) -> proof, called on rank 0 only.  Returns the proof on rank 0."""
    g1, g2 = shard_fn()
    flat = all_gather_partials(dist, pack_partials(g1, g2), world, device)
    if rank != 0:
        return None
    p1, p2 = unpack_gathered(flat, world, g1_words, g2_words)
    return finish_fn(p1, p2, world)
