"""Checker for the universal-setup (Marlin-style) path.  TEST INFRASTRUCTURE ONLY (see oracle/params.py).

Two things live here, both independent of the GPU:

* `IntBackend` -- the vector backend of `snark_b200/marlin.py` over Python integers (NTT from oracle/ntt.py, commitments as
  p(tau) * G with the trapdoor known).  Running the protocol over it gives the big-int proof the GPU proof must equal bit for bit.
* `verify` -- the verifier of the scheme, written from the protocol description and sharing NO arithmetic with the prover: it
  recomputes the transcript, checks the two AHP identities at beta1 / beta2 from the claimed evaluations, and checks the two
  batched KZG openings -- either with the trapdoor (P == tau * W in G1, the pairing-free form of e(P, H) = e(W, tau H); cheap) or
  with real pairings (oracle/pairing.py; slow, used on one small case).

PARITY UNPINNED: the reference tree holds only the trait (`UniversalSetupSNARK`, /root/reference/snark/src/lib.rs:107-133);
ark-marlin / ark-poly-commit are not vendored and cannot be built here.  The scheme follows the Marlin paper's AHP (section 5,
without zero-knowledge masks) and KZG10 with shifted-power degree bounds; its transcript is this repository's own.
"""
from typing import List

from snark_b200 import marlin as M

from . import ntt as ontt
from .ec import groups
from .params import Curve


class IntBackend:
    """Vectors are Python lists of canonical integers."""

    def __init__(self, curve: Curve):
        self.curve = curve
        self.r = curve.r
        self.fq_bytes = (curve.p.bit_length() + 7) // 8
        self.coset_gen = curve.fr_generator
        self.G1 = groups(curve)[0]

    def omega(self, log_n):
        return self.curve.omega(log_n)

    # -- plumbing
    def from_ints(self, xs):
        return [x % self.r for x in xs]

    def to_ints(self, v):
        return list(v)

    def pad(self, v, n):
        assert len(v) <= n
        return list(v) + [0] * (n - len(v))

    def slice(self, v, lo, hi):
        return list(v[lo:hi])

    def concat(self, vs):
        return [x for v in vs for x in v]

    def shifted(self, v, sh):
        return [0] * sh + list(v)

    # -- arithmetic
    def ntt(self, v, inverse=False, coset=False):
        if coset:
            return ontt.coset_intt(self.curve, v) if inverse else ontt.coset_ntt(self.curve, v)
        return ontt.ntt(self.curve, v, inverse=inverse)

    def mul(self, a, b):
        return [x * y % self.r for x, y in zip(a, b)]

    def add(self, a, b):
        return [(x + y) % self.r for x, y in zip(a, b)]

    def sub(self, a, b):
        return [(x - y) % self.r for x, y in zip(a, b)]

    def scale(self, a, s):
        return [x * s % self.r for x in a]

    def add_scalar(self, a, s):
        return [(x + s) % self.r for x in a]

    def geom(self, n, c, s):
        out, t = [], c % self.r
        for _ in range(n):
            out.append(t)
            t = t * s % self.r
        return out

    def inv0(self, a):
        return [pow(x, -1, self.r) if x else 0 for x in a]

    def eval(self, coeffs, z):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * z + c) % self.r
        return acc

    # -- matrices
    def upload_matrices(self, mats, n_rows, n_cols):
        return (mats, n_rows, n_cols)

    def spmv(self, handle, z, n):
        mats, n_rows, _ = handle
        out = []
        for Mx in mats:
            v = [sum(c * z[col] for c, col in row) % self.r for row in Mx]
            out.append(v + [0] * (n - len(v)))
        return out

    # -- KZG10 with the trapdoor known: commit(p) = p(tau) G, the same group element an MSM over the powers gives
    def setup(self, size, tau):
        return {"size": size, "tau": tau % self.r}

    def srs_size(self, srs):
        return srs["size"]

    def commit(self, srs, coeffs, shift=0):
        assert shift + len(coeffs) <= srs["size"]
        e = self.eval(coeffs, srs["tau"]) * pow(srs["tau"], shift, self.r) % self.r
        return self.G1.mul(self.G1.gen, e) if e else None


# ---- verifier -------------------------------------------------------------------------------------------------------------
def _interp_eval(curve: Curve, xs: List[int], l: int, at: int) -> int:
    """x^(at): the degree < l polynomial with x^(w_l^i) = xs[i] (zero-padded), by the direct Lagrange formula."""
    r = curve.r
    w = curve.omega(M.log2(l))
    vl = (pow(at, l, r) - 1) % r
    acc = 0
    for i in range(l):
        xi = xs[i] if i < len(xs) else 0
        if xi == 0:
            continue
        wi = pow(w, i, r)
        # L_i(at) = v_l(at) w^i / (l (at - w^i))
        acc += xi * vl % r * wi % r * pow(l * (at - wi) % r, -1, r)
    return acc % r


def verify(curve: Curve, vk: M.VerifierKey, x: List[int], proof: M.Proof, tau=None, tau_g2=None, engine=None) -> bool:
    """`SNARK::verify` (snark/src/lib.rs:56-75) for the universal-setup scheme.  Exactly one of `tau` (trapdoor form of the
    opening check) or (`tau_g2`, `engine`) (real pairings) must be given."""
    r = curve.r
    G1 = groups(curve)[0]
    info = vk.info
    n, m, l, D = info.n, info.m, info.l, info.D
    fq_bytes = (curve.p.bit_length() + 7) // 8
    if len(proof.comms) != 10 or len(proof.evals1) != 6 or len(proof.evals2) != 14 or len(proof.openings) != 2:
        return False
    for P in list(proof.comms) + list(proof.openings):
        if not G1.on_curve(P):
            return False
    c_w, c_zA, c_zB, c_t, c_g1, c_g1s, c_h1, c_g2, c_g2s, c_h2 = proof.comms
    tr = M.start_transcript(r, fq_bytes, info, vk.index_comms, x)
    tr.absorb_points([c_w, c_zA, c_zB], fq_bytes)
    alpha, eta_a, eta_b, eta_c = (tr.challenge() for _ in range(4))
    tr.absorb_points([c_t, c_g1, c_g1s, c_h1], fq_bytes)
    beta1 = tr.challenge()
    tr.absorb_points([c_g2, c_g2s, c_h2], fq_bytes)
    beta2 = tr.challenge()
    tr.absorb_ints(list(proof.evals1) + list(proof.evals2))
    xi1, xi2 = tr.challenge(), tr.challenge()

    w_b, zA_b, zB_b, t_b, g1_b, h1_b = proof.evals1
    g2_b, h2_b = proof.evals2[:2]
    idx = proof.evals2[2:]
    vh = lambda X: (pow(X, n, r) - 1) % r
    vk_ = lambda X: (pow(X, m, r) - 1) % r
    # identity over H:  r(alpha, X) (eta_A zA + eta_B zB + eta_C zA zB) - t z^ = h1 v_H + X g1   at X = beta1
    if (alpha - beta1) % r == 0 or vh(alpha) == 0 or vh(beta1) == 0:
        return False
    r_ab = (vh(alpha) - vh(beta1)) * pow(alpha - beta1, -1, r) % r
    z_b = (w_b * (pow(beta1, l, r) - 1) + _interp_eval(curve, x, l, beta1)) % r
    lhs = (r_ab * (eta_a * zA_b + eta_b * zB_b + eta_c * zA_b % r * zB_b) - t_b * z_b) % r
    if lhs != (h1_b * vh(beta1) + beta1 * g1_b) % r:
        return False
    # identity over K:  a - b (X g2 + t(beta1) / |K|) = h2 v_K   at X = beta2
    scale_ab = vh(alpha) * vh(beta1) % r
    dens, vals = [], []
    for k in range(3):
        row, col, val, rc = idx[4 * k: 4 * k + 4]
        dens.append((alpha * beta1 - alpha * col - beta1 * row + rc) % r)
        vals.append(val)
    b_v = dens[0] * dens[1] * dens[2] % r
    a_v = scale_ab * (eta_a * vals[0] * dens[1] * dens[2] + eta_b * vals[1] * dens[0] * dens[2] + eta_c * vals[2] * dens[0] * dens[1]) % r
    f2_v = (beta2 * g2_b + t_b * pow(m, -1, r)) % r
    if (a_v - b_v * f2_v) % r != h2_b * vk_(beta2) % r:
        return False
    # openings: C = sum xi^i C_i, v = sum xi^i v_i;  e(C - v G + beta W, H) = e(W, tau H)
    sh1, sh2 = D - (n - 2), D - (m - 2)
    batch1 = [c_w, c_zA, c_zB, c_t, c_g1, c_h1, c_g1s]
    vals1 = list(proof.evals1) + [pow(beta1, sh1, r) * g1_b % r]
    batch2 = [c_g2, c_h2] + list(vk.index_comms) + [c_g2s]
    vals2 = list(proof.evals2) + [pow(beta2, sh2, r) * g2_b % r]
    lhs_pts = []
    for comms, vals_, beta, xi, W in ((batch1, vals1, beta1, xi1, proof.openings[0]), (batch2, vals2, beta2, xi2, proof.openings[1])):
        C, v, c = None, 0, 1
        for P, pv in zip(comms, vals_):
            C = G1.add(C, G1.mul(P, c)) if P is not None else C
            v = (v + c * pv) % r
            c = c * xi % r
        P = G1.add(C, G1.neg(G1.mul(G1.gen, v))) if v else C
        P = G1.add(P, G1.mul(W, beta)) if W is not None else P
        lhs_pts.append((P, W))
    if tau is not None:
        return all(P == (G1.mul(W, tau % r) if W is not None else None) for P, W in lhs_pts)
    assert tau_g2 is not None and engine is not None
    G2 = groups(curve)[1]
    # one combined check with a transcript-derived weight: e(P1 + rho P2, H) = e(W1 + rho W2, tau H)
    tr.absorb_points([W for _, W in lhs_pts], fq_bytes)
    rho = tr.challenge()
    Pc = G1.add(lhs_pts[0][0], G1.mul(lhs_pts[1][0], rho) if lhs_pts[1][0] is not None else None)
    Wc = G1.add(lhs_pts[0][1], G1.mul(lhs_pts[1][1], rho) if lhs_pts[1][1] is not None else None)
    f = engine.miller_loop(G2.gen, Pc) * engine.miller_loop(tau_g2, G1.neg(Wc))
    return engine.final_exponentiation(f) == engine.Fq12.one()
