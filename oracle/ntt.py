"""Radix-2 NTT over Fr for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates ark-poly `Radix2EvaluationDomain::{fft,ifft}_in_place` and `get_coset`
(upstream crate not in /root/reference): natural order in and
out, omega_N = ROOT^(2^(S - log N)), iNTT scales by N^-1, coset NTT evaluates on g*omega^i.
Because results are exact field elements, any correct schedule is bit-identical.
"""
from .params import Curve


def dft_naive(curve: Curve, x, inverse=False):
    """O(n^2) definition: X[i] = sum_j x[j] * w^(i j).  Ground truth for the fast NTT."""
    r = curve.r
    n = len(x)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    w = curve.omega(log_n)
    if inverse:
        w = pow(w, -1, r)
    out = []
    for i in range(n):
        wi = pow(w, i, r)
        acc, t = 0, 1
        for j in range(n):
            acc += x[j] * t
            t = t * wi % r
        out.append(acc % r)
    if inverse:
        ninv = pow(n, -1, r)
        out = [v * ninv % r for v in out]
    return out


def _bitrev(i, bits):
    return int(bin(i)[2:].zfill(bits)[::-1], 2) if bits else 0


def ntt(curve: Curve, x, inverse=False):
    """Iterative Cooley-Tukey, natural -> natural."""
    r = curve.r
    n = len(x)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    a = [0] * n
    for i in range(n):
        a[_bitrev(i, log_n)] = x[i] % r
    w_n = curve.omega(log_n)
    if inverse:
        w_n = pow(w_n, -1, r)
    m = 1
    while m < n:
        w_m = pow(w_n, n // (2 * m), r)
        tw = [1] * m
        for j in range(1, m):
            tw[j] = tw[j - 1] * w_m % r
        for k in range(0, n, 2 * m):
            for j in range(m):
                t = a[k + j + m] * tw[j] % r
                u = a[k + j]
                a[k + j] = (u + t) % r
                a[k + j + m] = (u - t) % r
        m *= 2
    if inverse:
        ninv = pow(n, -1, r)
        a = [v * ninv % r for v in a]
    return a


def coset_ntt(curve: Curve, x, g=None):
    """X[i] = sum_j x[j] (g w^i)^j : scale x[j] by g^j then NTT."""
    r = curve.r
    g = curve.fr_generator if g is None else g
    t, y = 1, []
    for v in x:
        y.append(v * t % r)
        t = t * g % r
    return ntt(curve, y)


def coset_intt(curve: Curve, x, g=None):
    """Inverse of coset_ntt: iNTT then scale by g^-j."""
    r = curve.r
    g = curve.fr_generator if g is None else g
    ginv = pow(g, -1, r)
    y = ntt(curve, x, inverse=True)
    t, out = 1, []
    for v in y:
        out.append(v * t % r)
        t = t * ginv % r
    return out
