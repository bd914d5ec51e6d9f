"""Field-extension and elliptic-curve arithmetic for the CPU oracle (pure Python big-int).

TEST INFRASTRUCTURE ONLY (see oracle/params.py header).  Restates the group law that
ark-ec's short-Weierstrass `Affine`/`Projective` implement (SURVEY.md Appendix A.1/A.4;
upstream crate not in /root/reference).  Points: affine = (x, y) or None for infinity;
Jacobian = (X, Y, Z) with Z == zero for infinity.  A field is described by `Fld`.
"""
from .params import Curve


class Fld:
    """Minimal field interface: elements are ints (Fq) or 2-tuples (Fq2)."""

    def __init__(self, p, deg):
        self.p = p
        self.deg = deg
        if deg == 1:
            self.zero, self.one = 0, 1
        else:
            self.zero, self.one = (0, 0), (1, 0)

    # --- generic ops -----------------------------------------------------
    def add(self, a, b):
        p = self.p
        if self.deg == 1:
            return (a + b) % p
        return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)

    def sub(self, a, b):
        p = self.p
        if self.deg == 1:
            return (a - b) % p
        return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)

    def neg(self, a):
        p = self.p
        if self.deg == 1:
            return (-a) % p
        return ((-a[0]) % p, (-a[1]) % p)

    def mul(self, a, b):
        p = self.p
        if self.deg == 1:
            return a * b % p
        # (a0 + a1 u)(b0 + b1 u), u^2 = -1
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a):
        return self.mul(a, a)

    def muli(self, a, k: int):
        p = self.p
        if self.deg == 1:
            return a * k % p
        return (a[0] * k % p, a[1] * k % p)

    def inv(self, a):
        p = self.p
        if self.deg == 1:
            return pow(a, -1, p)
        n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
        return (a[0] * n % p, (-a[1]) * n % p)

    def is_zero(self, a):
        return a == self.zero

    def eq(self, a, b):
        return a == b


class Group:
    """Short-Weierstrass group y^2 = x^3 + b over `fld` (a = 0 for both curves)."""

    def __init__(self, fld: Fld, b, gen, order):
        self.f = fld
        self.b = b
        self.gen = gen
        self.order = order

    def on_curve(self, P):
        if P is None:
            return True
        f = self.f
        x, y = P
        return f.sqr(y) == f.add(f.mul(f.sqr(x), x), self.b)

    # --- affine ----------------------------------------------------------
    def neg(self, P):
        if P is None:
            return None
        return (P[0], self.f.neg(P[1]))

    def add(self, P, Q):
        f = self.f
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if P[1] == Q[1]:
                return self.dbl(P)
            return None
        lam = f.mul(f.sub(Q[1], P[1]), f.inv(f.sub(Q[0], P[0])))
        x3 = f.sub(f.sub(f.sqr(lam), P[0]), Q[0])
        y3 = f.sub(f.mul(lam, f.sub(P[0], x3)), P[1])
        return (x3, y3)

    def dbl(self, P):
        f = self.f
        if P is None or f.is_zero(P[1]):
            return None
        lam = f.mul(f.muli(f.sqr(P[0]), 3), f.inv(f.muli(P[1], 2)))
        x3 = f.sub(f.sqr(lam), f.muli(P[0], 2))
        y3 = f.sub(f.mul(lam, f.sub(P[0], x3)), P[1])
        return (x3, y3)

    # --- Jacobian (fast path for scalar-mul / MSM in the oracle) ----------
    def to_jac(self, P):
        f = self.f
        if P is None:
            return (f.one, f.one, f.zero)
        return (P[0], P[1], f.one)

    def to_affine(self, J):
        f = self.f
        X, Y, Z = J
        if f.is_zero(Z):
            return None
        zi = f.inv(Z)
        zi2 = f.sqr(zi)
        return (f.mul(X, zi2), f.mul(Y, f.mul(zi2, zi)))

    def jdbl(self, J):
        f = self.f
        X, Y, Z = J
        if f.is_zero(Z) or f.is_zero(Y):
            return (f.one, f.one, f.zero)
        A = f.sqr(X)
        B = f.sqr(Y)
        C = f.sqr(B)
        D = f.muli(f.sub(f.sub(f.sqr(f.add(X, B)), A), C), 2)
        E = f.muli(A, 3)
        F = f.sqr(E)
        X3 = f.sub(F, f.muli(D, 2))
        Y3 = f.sub(f.mul(E, f.sub(D, X3)), f.muli(C, 8))
        Z3 = f.muli(f.mul(Y, Z), 2)
        return (X3, Y3, Z3)

    def jadd(self, J1, J2):
        f = self.f
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        if f.is_zero(Z1):
            return J2
        if f.is_zero(Z2):
            return J1
        Z1Z1 = f.sqr(Z1)
        Z2Z2 = f.sqr(Z2)
        U1 = f.mul(X1, Z2Z2)
        U2 = f.mul(X2, Z1Z1)
        S1 = f.mul(f.mul(Y1, Z2), Z2Z2)
        S2 = f.mul(f.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jdbl(J1)
            return (f.one, f.one, f.zero)
        H = f.sub(U2, U1)
        R = f.sub(S2, S1)
        HH = f.sqr(H)
        HHH = f.mul(H, HH)
        V = f.mul(U1, HH)
        X3 = f.sub(f.sub(f.sqr(R), HHH), f.muli(V, 2))
        Y3 = f.sub(f.mul(R, f.sub(V, X3)), f.mul(S1, HHH))
        Z3 = f.mul(f.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def jadd_affine(self, J, P):
        if P is None:
            return J
        return self.jadd(J, (P[0], P[1], self.f.one))

    def jneg(self, J):
        return (J[0], self.f.neg(J[1]), J[2])

    def jmul(self, J, k: int):
        f = self.f
        k %= self.order
        acc = (f.one, f.one, f.zero)
        if k == 0:
            return acc
        for bit in bin(k)[2:]:
            acc = self.jdbl(acc)
            if bit == "1":
                acc = self.jadd(acc, J)
        return acc

    def mul(self, P, k: int):
        """Affine scalar multiplication k*P (double-and-add; the naive MSM building block)."""
        return self.to_affine(self.jmul(self.to_jac(P), k))


_cache = {}


def groups(curve: Curve):
    """Return (G1, G2) Group objects for `curve`."""
    if curve.name not in _cache:
        fq = Fld(curve.p, 1)
        fq2 = Fld(curve.p, 2)
        _cache[curve.name] = (
            Group(fq, curve.b, curve.g1, curve.r),
            Group(fq2, curve.b2, curve.g2, curve.r),
        )
    return _cache[curve.name]
