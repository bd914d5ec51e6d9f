// Short-Weierstrass group law (a = 0) in extended-Jacobian "XYZZ" coordinates.
//
// GPU counterpart of ark-ec's `short_weierstrass::{Affine, Projective}` additions that
// `VariableBaseMSM::msm_bigint` performs (upstream crate, not in /root/reference; SURVEY.md
// Appendix A.4).  A point (X, Y, ZZ, ZZZ) stands for x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2; the
// identity is ZZ = 0.  XYZZ is used instead of ark-ec's Jacobian because the bucket accumulation is
// dominated by accumulator += affine, which costs 8M + 2S here against 7M + 4S in Jacobian -- the
// group element, and therefore the affine result after normalisation, is the same.
//
// Affine points are (x, y) with the point at infinity encoded as (0, 0), which is never on
// y^2 = x^3 + b for b != 0 (include/b200snark.h documents the same convention for callers).
//
// F is Fp<P> (G1) or Fp2<P> (G2).  Formulas: EFD "xyzz" add-2008-s, madd-2008-s, dbl-2008-s-1,
// mdbl-2008-s-1, specialised to a = 0.  All special cases (identity operands, P == Q, P == -Q) are
// handled so the result is the exact group element for any input.
#pragma once
#include "ff.cuh"

namespace b2s {

template <class F>
struct Affine {
    F x, y;
    B2S_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    B2S_HD static Affine inf() { return {F::zero(), F::zero()}; }
    B2S_HD Affine neg() const { return {x, y.neg()}; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;

    B2S_HD static XYZZ identity() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    B2S_HD bool is_identity() const { return zz.is_zero(); }
    B2S_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return identity();
        return {p.x, p.y, F::one(), F::one()};
    }
    B2S_HD XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

    // 2 * (affine p)
    B2S_HD static XYZZ dbl_affine(const Affine<F>& p) {
        if (p.is_inf() || p.y.is_zero()) return identity();
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        F y3 = m * (s - x3) - w * p.y;
        return {x3, y3, v, w};
    }

    B2S_HD XYZZ dbl() const {
        if (is_identity() || y.is_zero()) return identity();
        F u = y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        F y3 = m * (s - x3) - w * y;
        return {x3, y3, v * zz, w * zzz};
    }

    // this += affine q   (the bucket-accumulation step)
    B2S_HD void add_affine(const Affine<F>& q) {
        if (q.is_inf()) return;
        if (is_identity()) {
            x = q.x; y = q.y; zz = F::one(); zzz = F::one();
            return;
        }
        F p = q.x * zz - x;      // U2 - X1
        F r = q.y * zzz - y;     // S2 - Y1
        if (p.is_zero()) {
            if (r.is_zero()) { *this = dbl_affine(q); }
            else { *this = identity(); }
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F qv = x * pp;
        F x3 = r.sqr() - ppp - qv.dbl();
        F y3 = r * (qv - x3) - y * ppp;
        x = x3; y = y3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += o
    B2S_HD void add(const XYZZ& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F u1 = x * o.zz;
        F u2 = o.x * zz;
        F s1 = y * o.zzz;
        F s2 = o.y * zzz;
        F p = u2 - u1;
        F r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) { *this = dbl(); }
            else { *this = identity(); }
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F qv = u1 * pp;
        F x3 = r.sqr() - ppp - qv.dbl();
        F y3 = r * (qv - x3) - s1 * ppp;
        x = x3; y = y3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }

    // Normalise (one field inversion).
    B2S_HD Affine<F> to_affine() const {
        if (is_identity()) return Affine<F>::inf();
        // 1/ZZZ gives both: 1/ZZ = ZZ^2 * (1/ZZZ)^2 * ... use  a = 1/zzz ; 1/zz = (a * zz)^2
        F a = zzz.inverse();
        F b = (a * zz).sqr();  // 1/zz   since zz^3 = zzz^2  =>  (zz/zzz)^2 = 1/zz
        return {x * b, y * a};
    }
};

// k * P by left-to-right double-and-add over little-endian 32-bit words (k NOT in Montgomery form).
template <class F>
B2S_HD XYZZ<F> scalar_mul_words(const XYZZ<F>& p, const uint32_t* k, int nwords) {
    XYZZ<F> acc = XYZZ<F>::identity();
    bool started = false;
    for (int w = nwords - 1; w >= 0; w--) {
        for (int b = 31; b >= 0; b--) {
            if (started) acc = acc.dbl();
            if ((k[w] >> b) & 1) {
                acc.add(p);
                started = true;
            }
        }
    }
    return acc;
}

}  // namespace b2s
