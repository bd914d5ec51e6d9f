"""BLS12-381 ate pairing and the Groth16 verification equation, for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

SURVEY.md 8(c) item 5: an implementation-independent check of proofs that does not use the setup trapdoor --
`SNARK::verify` (/root/reference/snark/src/lib.rs:57-66) for Groth16:
        e(A, B) = e(alpha_1, beta_2) * e(sum_j x_j gamma_abc_j, gamma_2) * e(C, delta_2).
Plain, slow, textbook construction (seconds per check): Fq12 = Fq[w] / (w^12 - 2 w^6 + 2) (so that w^6 = 1 + u),
G2 points moved to E(Fq12) by (x, y) -> (x / w^2, y / w^3), affine Miller loop over |x| = 0xd201000000010000, final
exponentiation by a direct power (p^12 - 1) / r.  Any bilinear, non-degenerate map gives a sound equality test;
bilinearity and non-degeneracy are asserted in tests/test_oracle_pairing.py.  BLS12-381 only.
"""
from .params import BLS12_381

P = BLS12_381.p
R_ORDER = BLS12_381.r
ATE_LOOP = 0xD201000000010000
# modulus polynomial of Fq12 over Fq: w^12 = 2 w^6 - 2
_MOD6, _MOD0 = 2, -2


class Fq12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % P for x in c] + [0] * (12 - len(c))

    @staticmethod
    def one():
        return Fq12([1])

    @staticmethod
    def zero():
        return Fq12([0])

    def __eq__(self, o):
        return self.c == o.c

    def is_zero(self):
        return not any(self.c)

    def __add__(self, o):
        return Fq12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return Fq12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return Fq12([-a for a in self.c])

    def scale(self, k):
        return Fq12([a * k for a in self.c])

    def __mul__(self, o):
        t = [0] * 23
        a, b = self.c, o.c
        for i in range(12):
            ai = a[i]
            if ai:
                for j in range(12):
                    t[i + j] += ai * b[j]
        for k in range(22, 11, -1):          # w^k = w^(k-12) * (2 w^6 - 2)
            v = t[k]
            if v:
                t[k - 6] += _MOD6 * v
                t[k - 12] += _MOD0 * v
        return Fq12(t[:12])

    def sqr(self):
        return self * self

    def pow(self, e):
        out, base = Fq12.one(), self
        while e:
            if e & 1:
                out = out * base
            base = base * base
            e >>= 1
        return out

    def inv(self):
        """Extended Euclid on polynomials over Fq (deg < 12) against the modulus."""
        def deg(p_):
            d = len(p_) - 1
            while d and p_[d] == 0:
                d -= 1
            return d

        lm, hm = [1] + [0] * 12, [0] * 13
        low = self.c + [0]
        high = [(-_MOD0) % P, 0, 0, 0, 0, 0, (-_MOD6) % P, 0, 0, 0, 0, 0, 1]   # w^12 - 2 w^6 + 2
        while deg(low):
            # r = high / low  (polynomial quotient)
            dl, dh = deg(low), deg(high)
            quo = [0] * 13
            temp = list(high)
            inv_lead = pow(low[dl], -1, P)
            for i in range(dh - dl, -1, -1):
                q = temp[dl + i] * inv_lead % P
                quo[i] = q
                for k in range(dl + 1):
                    temp[i + k] = (temp[i + k] - q * low[k]) % P
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * quo[j]) % P
                    new[i + j] = (new[i + j] - low[i] * quo[j]) % P
            lm, low, hm, high = nm, new, lm, low
        inv0 = pow(low[0], -1, P)
        return Fq12([x * inv0 for x in lm[:12]])

    def __truediv__(self, o):
        return self * o.inv()


W = Fq12([0, 1])
W2_INV = (W * W).inv()
W3_INV = (W * W * W).inv()


def embed_fq(x):
    return Fq12([x])


def embed_fq2(a):
    """c0 + c1 u with u = w^6 - 1."""
    return Fq12([a[0] - a[1], 0, 0, 0, 0, 0, a[1]])


def twist(Q):
    """G2 affine point over Fq2 -> point of y^2 = x^3 + 4 over Fq12."""
    return (embed_fq2(Q[0]) * W2_INV, embed_fq2(Q[1]) * W3_INV)


def _double(Pt):
    x, y = Pt
    m = x.sqr().scale(3) / y.scale(2)
    nx = m.sqr() - x.scale(2)
    return (nx, m * (x - nx) - y)


def _add(P1, P2):
    if P1 is None:
        return P2
    if P2 is None:
        return P1
    x1, y1 = P1
    x2, y2 = P2
    if x1 == x2:
        return _double(P1) if y1 == y2 else None
    m = (y2 - y1) / (x2 - x1)
    nx = m.sqr() - x1 - x2
    return (nx, m * (x1 - nx) - y1)


def _line(P1, P2, T):
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if not x1 == x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = x1.sqr().scale(3) / y1.scale(2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(Q, Pt):
    """Q: G2 affine (Fq2 tuples) or None; Pt: G1 affine (ints) or None.  Returns the un-exponentiated value."""
    if Q is None or Pt is None:
        return Fq12.one()
    Qt = twist(Q)
    Pe = (embed_fq(Pt[0]), embed_fq(Pt[1]))
    Rt, f = Qt, Fq12.one()
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        f = f.sqr() * _line(Rt, Rt, Pe)
        Rt = _double(Rt)
        if (ATE_LOOP >> i) & 1:
            f = f * _line(Rt, Qt, Pe)
            Rt = _add(Rt, Qt)
    return f


def final_exponentiation(f):
    return f.pow((P ** 12 - 1) // R_ORDER)


def pairing(Pt, Q):
    """e(P, Q) for P in G1, Q in G2 (affine oracle points)."""
    return final_exponentiation(miller_loop(Q, Pt))


def groth16_verify(vk, public_inputs, proof):
    """vk: dict(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1[list]); public_inputs: [x_1..x_{l-1}] (without the
    leading 1); proof: (A, B, C) affine.  e(A,B) * e(-IC, gamma) * e(-C, delta) * e(-alpha, beta) == 1 with ONE final
    exponentiation (what ark-groth16's verify_with_processed_vk does up to precomputation)."""
    from .ec import groups

    G1, _ = groups(BLS12_381)
    A, B, C = proof
    ic = vk["gamma_abc_g1"][0]
    for x, base in zip(public_inputs, vk["gamma_abc_g1"][1:]):
        ic = G1.add(ic, G1.mul(base, x))
    f = miller_loop(B, A)
    f = f * miller_loop(vk["gamma_g2"], G1.neg(ic))
    f = f * miller_loop(vk["delta_g2"], G1.neg(C))
    f = f * miller_loop(vk["beta_g2"], G1.neg(vk["alpha_g1"]))
    return final_exponentiation(f) == Fq12.one()
