"""Variable-base MSM for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

`msm_naive` is the definition (sum of double-and-add scalar multiples).  `msm_pippenger`
restates ark-ec `VariableBaseMSM::msm_bigint_wnaf` (upstream crate not
in /root/reference): window c = 3 if n < 32 else floor(ceil(log2 n) * 69 / 100) + 2, signed
radix-2^c digits, 2^(c-1) buckets per window, running-sum bucket reduction, windows combined
high -> low with c doublings.  The result is a group element, so every correct MSM agrees
after normalisation to affine.
"""
from .ec import Group


def ark_window_bits(n: int) -> int:
    """ark-ec's window-size rule (ln_without_floats(n) = ceil(log2 n) * 69 / 100)."""
    if n < 32:
        return 3
    log2 = (n - 1).bit_length()  # ceil(log2 n) for n >= 2
    return log2 * 69 // 100 + 2


def reference_add_count(n: int, scalar_bits: int) -> int:
    """G1-adds the reference algorithm performs: N*ceil(l/c) + ceil(l/c)*2^c."""
    c = ark_window_bits(n)
    w = -(-scalar_bits // c)
    return n * w + w * (1 << c)


def signed_digits(k: int, c: int, num_bits: int):
    """Signed radix-2^c digits of k, least-significant first, each in [-2^(c-1), 2^(c-1))
    except the last which absorbs the carry."""
    nw = -(-num_bits // c)
    if nw * c == num_bits:
        nw += 0  # ark-ec: digits = ceil(bits / c); final carry folded into the last digit
    digits = []
    carry = 0
    radix = 1 << c
    half = radix >> 1
    for i in range(nw):
        d = ((k >> (i * c)) & (radix - 1)) + carry
        carry = 0
        if i != nw - 1 and d >= half:
            d -= radix
            carry = 1
        digits.append(d)
    return digits


def msm_naive(G: Group, bases, scalars):
    acc = G.to_jac(None)
    for P, k in zip(bases, scalars):
        if P is None or k % G.order == 0:
            continue
        acc = G.jadd(acc, G.jmul(G.to_jac(P), k))
    return G.to_affine(acc)


def msm_pippenger(G: Group, bases, scalars, c=None):
    n = min(len(bases), len(scalars))
    if n == 0:
        return None
    num_bits = G.order.bit_length()
    c = ark_window_bits(n) if c is None else c
    digs = [signed_digits(k % G.order, c, num_bits) for k in scalars[:n]]
    nw = len(digs[0])
    zero = G.to_jac(None)
    window_sums = []
    for w in range(nw):
        # last window digits may reach 2^c - 1 + carry, so size buckets for that
        buckets = {}
        for i in range(n):
            d = digs[i][w]
            P = bases[i]
            if d == 0 or P is None:
                continue
            if d > 0:
                buckets[d] = G.jadd_affine(buckets.get(d, zero), P)
            else:
                buckets[-d] = G.jadd_affine(buckets.get(-d, zero), G.neg(P))
        running, total = zero, zero
        if buckets:
            for b in range(max(buckets), 0, -1):
                if b in buckets:
                    running = G.jadd(running, buckets[b])
                total = G.jadd(total, running)
        window_sums.append(total)
    acc = window_sums[-1]
    for w in range(nw - 2, -1, -1):
        for _ in range(c):
            acc = G.jdbl(acc)
        acc = G.jadd(acc, window_sums[w])
    return G.to_affine(acc)
