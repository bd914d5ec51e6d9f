"""Ate pairings on BLS12-381 and BN254 and the Groth16 verification equation, for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

SURVEY.md 8(c) item 5: an implementation-independent check of proofs that does not use the setup trapdoor --
`SNARK::verify` (/root/reference/snark/src/lib.rs:57-66) for Groth16:
        e(A, B) = e(alpha_1, beta_2) * e(sum_j x_j gamma_abc_j, gamma_2) * e(C, delta_2).
Plain, slow, textbook construction (seconds per check), the same for both curves:
  * Fq12 = Fq[w] / (w^12 - 2 c w^6 + c^2 + 1), i.e. w^6 = c + u with u^2 = -1 (c = 1 for BLS12-381, 9 for BN254);
  * G2 points are moved to E(Fq12) by the twist isomorphism -- (x / w^2, y / w^3) for BLS12-381's M-type twist
    y^2 = x^3 + 4 (1 + u), (x w^2, y w^3) for BN254's D-type twist y^2 = x^3 + 3 / (9 + u);
  * the ate pairing f_{T,Q}(P)^((p^12 - 1) / r) with T = t - 1 (|x| = 0xd201000000010000 for BLS12-381; 6 x^2 with
    x = 4965661367192848881 for BN254 -- the plain ate loop, so no Frobenius correction steps are needed), affine
    Miller loop, final exponentiation by a direct power.
Any bilinear, non-degenerate map gives a sound equality test; bilinearity and non-degeneracy are asserted in
tests/test_oracle_pairing.py, so the construction does not have to be trusted.
"""
from .params import BLS12_381, BN254, Curve


def _make_fq12(P, c):
    """Fq[w] / (w^12 - 2 c w^6 + (c^2 + 1)) as a class over coefficient lists."""
    MOD6, MOD0 = 2 * c, -(c * c + 1)          # w^12 = MOD6 w^6 + MOD0

    class Fq12:
        __slots__ = ("c",)

        def __init__(self, coeffs):
            self.c = [x % P for x in coeffs] + [0] * (12 - len(coeffs))

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
            for k in range(22, 11, -1):          # w^k = w^(k-12) * (MOD6 w^6 + MOD0)
                v = t[k]
                if v:
                    t[k - 6] += MOD6 * v
                    t[k - 12] += MOD0 * v
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
            high = [(-MOD0) % P, 0, 0, 0, 0, 0, (-MOD6) % P, 0, 0, 0, 0, 0, 1]
            while deg(low):
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

    return Fq12


class Engine:
    """One pairing-friendly curve: its Fq12, twist map, Miller loop and final exponentiation."""

    def __init__(self, curve: Curve, xi0: int, loop: int, twist: str):
        self.curve = curve
        self.P, self.R = curve.p, curve.r
        self.xi0 = xi0                       # w^6 = xi0 + u
        self.loop = loop                     # T = |t - 1|
        self.Fq12 = _make_fq12(curve.p, xi0)
        w = self.Fq12([0, 1])
        w2, w3 = w * w, w * w * w
        self.tx, self.ty = (w2.inv(), w3.inv()) if twist == "M" else (w2, w3)

    def embed_fq(self, x):
        return self.Fq12([x])

    def embed_fq2(self, a):
        """c0 + c1 u with u = w^6 - xi0."""
        return self.Fq12([a[0] - self.xi0 * a[1], 0, 0, 0, 0, 0, a[1]])

    def twist(self, Q):
        """G2 affine point over Fq2 -> point of y^2 = x^3 + b over Fq12."""
        return (self.embed_fq2(Q[0]) * self.tx, self.embed_fq2(Q[1]) * self.ty)

    @staticmethod
    def _double(Pt):
        x, y = Pt
        m = x.sqr().scale(3) / y.scale(2)
        nx = m.sqr() - x.scale(2)
        return (nx, m * (x - nx) - y)

    @classmethod
    def _add(cls, P1, P2):
        if P1 is None:
            return P2
        if P2 is None:
            return P1
        x1, y1 = P1
        x2, y2 = P2
        if x1 == x2:
            return cls._double(P1) if y1 == y2 else None
        m = (y2 - y1) / (x2 - x1)
        nx = m.sqr() - x1 - x2
        return (nx, m * (x1 - nx) - y1)

    @staticmethod
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

    def miller_loop(self, Q, Pt):
        """Q: G2 affine (Fq2 tuples) or None; Pt: G1 affine (ints) or None.  Returns the un-exponentiated value."""
        if Q is None or Pt is None:
            return self.Fq12.one()
        Qt = self.twist(Q)
        Pe = (self.embed_fq(Pt[0]), self.embed_fq(Pt[1]))
        Rt, f = Qt, self.Fq12.one()
        for i in range(self.loop.bit_length() - 2, -1, -1):
            f = f.sqr() * self._line(Rt, Rt, Pe)
            Rt = self._double(Rt)
            if (self.loop >> i) & 1:
                f = f * self._line(Rt, Qt, Pe)
                Rt = self._add(Rt, Qt)
        return f

    def final_exponentiation(self, f):
        return f.pow((self.P ** 12 - 1) // self.R)

    def pairing(self, Pt, Q):
        """e(P, Q) for P in G1, Q in G2 (affine oracle points)."""
        return self.final_exponentiation(self.miller_loop(Q, Pt))

    def groth16_verify(self, vk, public_inputs, proof):
        """vk: dict(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1[list]); public_inputs: [x_1..x_{l-1}] (without
        the leading 1); proof: (A, B, C) affine.  e(A,B) * e(-IC, gamma) * e(-C, delta) * e(-alpha, beta) == 1 with ONE
        final exponentiation (what ark-groth16's verify_with_processed_vk does up to precomputation)."""
        from .ec import groups

        G1, _ = groups(self.curve)
        A, B, C = proof
        ic = vk["gamma_abc_g1"][0]
        for x, base in zip(public_inputs, vk["gamma_abc_g1"][1:]):
            ic = G1.add(ic, G1.mul(base, x))
        f = self.miller_loop(B, A)
        f = f * self.miller_loop(vk["gamma_g2"], G1.neg(ic))
        f = f * self.miller_loop(vk["delta_g2"], G1.neg(C))
        f = f * self.miller_loop(vk["beta_g2"], G1.neg(vk["alpha_g1"]))
        return self.final_exponentiation(f) == self.Fq12.one()


_BN_X = 4965661367192848881
assert 36 * _BN_X ** 4 + 36 * _BN_X ** 3 + 24 * _BN_X ** 2 + 6 * _BN_X + 1 == BN254.p      # BN family polynomials
assert 36 * _BN_X ** 4 + 36 * _BN_X ** 3 + 18 * _BN_X ** 2 + 6 * _BN_X + 1 == BN254.r
_BLS_X = 0xD201000000010000                                                               # |x|; x is negative
assert (_BLS_X ** 4 - _BLS_X ** 2 + 1) == BLS12_381.r and ((_BLS_X + 1) ** 2 * BLS12_381.r) // 3 + (-_BLS_X) == BLS12_381.p

_ENGINES = {}


def engine(curve: Curve) -> Engine:
    if curve.name not in _ENGINES:
        if curve is BLS12_381:
            _ENGINES[curve.name] = Engine(curve, xi0=1, loop=_BLS_X, twist="M")             # t - 1 = x
        elif curve is BN254:
            _ENGINES[curve.name] = Engine(curve, xi0=9, loop=6 * _BN_X ** 2, twist="D")     # t - 1 = 6 x^2
        else:
            raise ValueError(curve.name)
    return _ENGINES[curve.name]


# BLS12-381 shorthands (the first curve this module supported; tests use them)
_BLS = engine(BLS12_381)
Fq12 = _BLS.Fq12
miller_loop = _BLS.miller_loop
final_exponentiation = _BLS.final_exponentiation
pairing = _BLS.pairing


def groth16_verify(vk, public_inputs, proof, curve: Curve = BLS12_381):
    return engine(curve).groth16_verify(vk, public_inputs, proof)
