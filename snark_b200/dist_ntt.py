"""Host-side plan of the distributed ("four-step") NTT for the multi-GPU witness map.

NOT WIRED TO KERNELS YET: today every rank repeats the 2^24-point witness map (36 of 151 ms per proof at 8 GPUs).  This
module fixes the index conventions the GPU version will use and lets them be tested over gloo with a stand-in engine
(tests/test_multirank_cpu.py), the way snark_b200/shard.py is.

One transform of size N = N1 * N2 over W ranks, natural order in and out, ONE all-to-all:

    input index  n = N2 * n1 + n2      output index  k = k1 + N1 * k2          w = primitive N-th root
    X[k1 + N1 k2] = sum_{n2} w^(N1 n2 k2) * [ w^(n2 k1) * sum_{n1} w^(N2 n1 k1) x[N2 n1 + n2] ]
                            step 3 (size N2)      step 2          step 1 (size N1)

  layout IN   rank r holds the columns n2 in [r*B2, (r+1)*B2), B2 = N2 / W, as a[n1][j]   (n mod N2 in its block)
  step 1, 2   local: N1-point transforms down the columns, then the twiddle w^(n2 k1)
  exchange    rank r sends rows k1 in [s*B1, (s+1)*B1), B1 = N1 / W, to rank s  (B1 x B2 elements per pair)
  step 3      local: N2-point transforms along the rows
  layout OUT  rank s holds k1 in its block, all k2, as out[i][k2]                    (k mod N1 in its block)

With N1 == N2 the output layout of one transform IS the input layout of the next (transpose the local tile), so the
witness map's chain iNTT -> coset NTT -> pointwise -> coset iNTT keeps one distribution pattern throughout; the pointwise
steps and the coset scalings g^n are index-local.  The consumers adapt to the pattern instead of re-gathering: the SpMV
computes the rows n with n mod N2 in the rank's block, and the h-MSM shard takes the h_query bases at the indices the
rank ends up holding.

The field arithmetic is behind `engine` (the GPU kernels in the product, an oracle stand-in in tests):
    engine.ntt_axis0(a, inverse)          a: (n, m, ...) -> size-n transforms along axis 0, natural order, no 1/n scaling
    engine.ntt_axis1(a, inverse)          the same along axis 1
    engine.mul_pow(a, row0, col0, inverse, N)   a[i][j] *= w_N^(+-(row0 + i) * (col0 + j))
    engine.mul_geometric(a, idx, base_inv, scale_n_inv, N)   coset / 1/N scaling: a[i][j] *= f(idx[i][j])   (see below)
"""
import numpy as np


def block(total, rank, world):
    if total % world:
        raise ValueError(f"{total} does not split over {world} ranks")
    b = total // world
    return rank * b, b


def owned_indices_in(N1, N2, rank, world):
    """Global indices n = N2 n1 + n2 of the IN layout, as an (N1, B2) integer array."""
    lo, b2 = block(N2, rank, world)
    return (np.arange(N1, dtype=np.int64)[:, None] * N2) + (lo + np.arange(b2, dtype=np.int64))[None, :]


def owned_indices_out(N1, N2, rank, world):
    """Global indices k = k1 + N1 k2 of the OUT layout, as a (B1, N2) integer array."""
    lo, b1 = block(N1, rank, world)
    return (lo + np.arange(b1, dtype=np.int64))[:, None] + N1 * np.arange(N2, dtype=np.int64)[None, :]


def all_to_all_tiles(dist, tiles, world):
    """tiles[s]: the array this rank sends to rank s (all the same shape).  Returns the list received, by source
    rank.  Uses all_to_all_single where the backend has it (NCCL) and an all-gather otherwise (gloo)."""
    import torch

    if world == 1:
        return [tiles[0]]
    send = torch.from_numpy(np.ascontiguousarray(np.stack(tiles)).view(np.int32))
    try:
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1))
        out = recv
    except (RuntimeError, NotImplementedError):
        gathered = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(gathered, send)
        me = dist.get_rank()
        out = torch.stack([g[me] for g in gathered])
    arr = out.numpy().view(np.uint32)
    return [arr[s] for s in range(world)]


class FourStepNtt:
    def __init__(self, dist, rank, world, log_n, engine):
        self.dist, self.rank, self.world, self.engine = dist, rank, world, engine
        self.log_n = log_n
        self.N = 1 << log_n
        self.N1 = 1 << (log_n // 2)
        self.N2 = self.N // self.N1
        block(self.N1, rank, world), block(self.N2, rank, world)      # divisibility

    def transform(self, a, inverse=False, coset=False):
        """a: this rank's IN tile, shape (N1, B2, limbs).  Returns its OUT tile, shape (B1, N2, limbs).
        forward: X[k] = sum x[n] (g w^k)^n (g = 1 unless coset); inverse: the exact inverse (1/N and g^-k included)."""
        N1, N2, N, W, r, e = self.N1, self.N2, self.N, self.world, self.rank, self.engine
        lo2, b2 = block(N2, r, W)
        lo1, b1 = block(N1, r, W)
        if coset and not inverse:
            a = e.mul_geometric(a, owned_indices_in(N1, N2, r, W), False, False, N)      # x[n] *= g^n
        a = e.ntt_axis0(a, inverse)                                                     # step 1
        a = e.mul_pow(a, 0, lo2, inverse, N)                                            # step 2: w^(+-k1 n2)
        tiles = [a[s * b1:(s + 1) * b1] for s in range(W)]                              # rows k1 of rank s
        got = all_to_all_tiles(self.dist, tiles, W)                                     # from rank q: its columns n2
        z = np.concatenate(got, axis=1)                                                 # (B1, N2, limbs)
        z = e.ntt_axis1(z, inverse)                                                     # step 3
        if inverse:
            z = e.mul_geometric(z, owned_indices_out(N1, N2, r, W), coset, True, N)     # * 1/N (* g^-k)
        return z

    def out_as_in(self, z):
        """OUT tile (B1, N2) of one transform -> IN tile (N1, B2) of the next; needs N1 == N2 (even log_n)."""
        if self.N1 != self.N2:
            raise ValueError("chaining transforms without an exchange needs N1 == N2")
        return np.ascontiguousarray(np.swapaxes(z, 0, 1))
