"""Groth16 setup / witness_map / prove for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates ark-groth16 (`generator.rs`, `r1cs_to_qap.rs` LibsnarkReduction, `prover.rs`) as
recalled in SURVEY.md Appendix A.1/A.2/A.5 (crate not in /root/reference; parity unpinned, see
oracle/params.py).  The setup keeps its trapdoor so that a proof can be checked *in the exponent*
(Appendix A.6) with three scalar multiplications and no pairing.

Trait surface mirrored: `SNARK::circuit_specific_setup` / `SNARK::prove`
(/root/reference/snark/src/lib.rs:43-54).
"""
from dataclasses import dataclass
from typing import List, Optional

from .ec import groups
from .msm import msm_pippenger
from .ntt import coset_intt, coset_ntt, ntt
from .params import Curve
from .r1cs import mat_vec_mul


def domain_size(num_constraints, num_instance):
    n = num_constraints + num_instance
    return 1 << max((n - 1).bit_length(), 1) if n > 1 else 1


@dataclass
class Trapdoor:
    tau: int
    alpha: int
    beta: int
    gamma: int
    delta: int


@dataclass
class ProvingKey:
    curve: Curve
    alpha_g1: tuple
    beta_g1: tuple
    beta_g2: tuple
    delta_g1: tuple
    delta_g2: tuple
    gamma_g2: tuple
    gamma_abc_g1: List[tuple]
    a_query: List[Optional[tuple]]
    b_g1_query: List[Optional[tuple]]
    b_g2_query: List[Optional[tuple]]
    h_query: List[Optional[tuple]]
    l_query: List[Optional[tuple]]
    # oracle-only: discrete logs, for the pairing-free check
    trapdoor: Trapdoor = None
    a_tau: List[int] = None
    b_tau: List[int] = None
    c_tau: List[int] = None
    domain: int = 0
    num_instance: int = 0


def lagrange_at_tau(curve: Curve, N: int, tau: int):
    """u[i] = L_i(tau) over the size-N radix-2 domain: Z(tau) * w^i / (N (tau - w^i))."""
    r = curve.r
    log_n = N.bit_length() - 1
    w = curve.omega(log_n)
    z = (pow(tau, N, r) - 1) % r
    assert z != 0, "tau lies in the domain"
    ninv = pow(N, -1, r)
    u, wi = [], 1
    for _ in range(N):
        u.append(z * wi % r * ninv % r * pow((tau - wi) % r, -1, r) % r)
        wi = wi * w % r
    return u


def qap_at_tau(curve: Curve, mats, num_instance, num_vars, N, tau):
    """A_j(tau), B_j(tau), C_j(tau) per variable (Appendix A.5, LibsnarkReduction instance map)."""
    r = curve.r
    A, B, C = mats
    n = len(A)
    u = lagrange_at_tau(curve, N, tau)
    a = [0] * num_vars
    b = [0] * num_vars
    c = [0] * num_vars
    for i in range(n):
        for coeff, col in A[i]:
            a[col] = (a[col] + u[i] * coeff) % r
        for coeff, col in B[i]:
            b[col] = (b[col] + u[i] * coeff) % r
        for coeff, col in C[i]:
            c[col] = (c[col] + u[i] * coeff) % r
    for i in range(num_instance):
        a[i] = (a[i] + u[n + i]) % r
    return a, b, c


def setup(curve: Curve, mats, num_instance, num_witness, td: Trapdoor) -> ProvingKey:
    """circuit_specific_setup with a caller-supplied trapdoor (snark/src/lib.rs:43-46)."""
    r = curve.r
    G1, G2 = groups(curve)
    n = len(mats[0])
    num_vars = num_instance + num_witness
    N = domain_size(n, num_instance)
    a, b, c = qap_at_tau(curve, mats, num_instance, num_vars, N, td.tau)
    zt = (pow(td.tau, N, r) - 1) % r
    dinv = pow(td.delta, -1, r)
    ginv = pow(td.gamma, -1, r)
    g1 = lambda k: G1.mul(G1.gen, k)
    g2 = lambda k: G2.mul(G2.gen, k)
    abc = [(td.beta * a[j] + td.alpha * b[j] + c[j]) % r for j in range(num_vars)]
    h_query, t = [], zt * dinv % r
    for _ in range(N - 1):
        h_query.append(g1(t))
        t = t * td.tau % r
    return ProvingKey(
        curve=curve,
        alpha_g1=g1(td.alpha),
        beta_g1=g1(td.beta),
        beta_g2=g2(td.beta),
        delta_g1=g1(td.delta),
        delta_g2=g2(td.delta),
        gamma_g2=g2(td.gamma),
        gamma_abc_g1=[g1(abc[j] * ginv % r) for j in range(num_instance)],
        a_query=[g1(v) for v in a],
        b_g1_query=[g1(v) for v in b],
        b_g2_query=[g2(v) for v in b],
        h_query=h_query,
        l_query=[g1(abc[j] * dinv % r) for j in range(num_instance, num_vars)],
        trapdoor=td,
        a_tau=a,
        b_tau=b,
        c_tau=c,
        domain=N,
        num_instance=num_instance,
    )


def witness_map(curve: Curve, mats, z, num_instance):
    """LibsnarkReduction::witness_map_from_matrices (Appendix A.2) -> h of length N."""
    r = curve.r
    A, B, C = mats
    n = len(A)
    N = domain_size(n, num_instance)
    a = mat_vec_mul(r, A, z) + [0] * (N - n)
    b = mat_vec_mul(r, B, z) + [0] * (N - n)
    c = mat_vec_mul(r, C, z) + [0] * (N - n)
    for i in range(num_instance):
        a[n + i] = z[i]
    a, b, c = (ntt(curve, v, inverse=True) for v in (a, b, c))
    a, b, c = (coset_ntt(curve, v) for v in (a, b, c))
    g = curve.fr_generator
    zinv = pow((pow(g, N, r) - 1) % r, -1, r)
    ab = [(a[i] * b[i] - c[i]) * zinv % r for i in range(N)]
    return coset_intt(curve, ab)


def prove(pk: ProvingKey, mats, z_inst, z_wit, r_rand, s_rand, msm=msm_pippenger):
    """create_proof_with_reduction (Appendix A.1). Returns affine (A in G1, B in G2, C in G1)."""
    curve = pk.curve
    G1, G2 = groups(curve)
    z = list(z_inst) + list(z_wit)
    h = witness_map(curve, mats, z, len(z_inst))
    J1, J2 = G1.to_jac, G2.to_jac
    h_acc = J1(msm(G1, pk.h_query, h[: len(pk.h_query)]))
    l_acc = J1(msm(G1, pk.l_query, z_wit))

    def calc(G, query, vk_param, delta, rnd):
        acc = G.to_jac(msm(G, query[1:], z[1:]))
        res = G.jmul(G.to_jac(delta), rnd)
        res = G.jadd_affine(res, query[0])
        res = G.jadd(res, acc)
        return G.jadd_affine(res, vk_param)

    g_a = calc(G1, pk.a_query, pk.alpha_g1, pk.delta_g1, r_rand)
    g1_b = calc(G1, pk.b_g1_query, pk.beta_g1, pk.delta_g1, s_rand)
    g2_b = calc(G2, pk.b_g2_query, pk.beta_g2, pk.delta_g2, s_rand)
    rs = r_rand * s_rand % curve.r
    g_c = G1.jmul(g_a, s_rand)
    g_c = G1.jadd(g_c, G1.jmul(g1_b, r_rand))
    g_c = G1.jadd(g_c, G1.jneg(G1.jmul(G1.to_jac(pk.delta_g1), rs)))
    g_c = G1.jadd(g_c, l_acc)
    g_c = G1.jadd(g_c, h_acc)
    return G1.to_affine(g_a), G2.to_affine(g2_b), G1.to_affine(g_c), h


def expected_proof_exponents(pk: ProvingKey, z_inst, z_wit, h, r_rand, s_rand):
    """Appendix A.6: discrete logs (a*, b*, c*) of a correct proof under the known trapdoor."""
    curve, td = pk.curve, pk.trapdoor
    r = curve.r
    z = list(z_inst) + list(z_wit)
    ell = len(z_inst)
    a_star = (td.alpha + sum(zj * aj for zj, aj in zip(z, pk.a_tau)) + r_rand * td.delta) % r
    b_star = (td.beta + sum(zj * bj for zj, bj in zip(z, pk.b_tau)) + s_rand * td.delta) % r
    dinv = pow(td.delta, -1, r)
    wit = sum(
        z[j] * (td.beta * pk.a_tau[j] + td.alpha * pk.b_tau[j] + pk.c_tau[j]) for j in range(ell, len(z))
    ) % r
    zt = (pow(td.tau, pk.domain, r) - 1) % r
    h_tau = sum(hc * pow(td.tau, i, r) for i, hc in enumerate(h[: pk.domain - 1])) % r
    c_star = ((wit + h_tau * zt) * dinv + s_rand * a_star + r_rand * b_star - r_rand * s_rand % r * td.delta) % r
    return a_star, b_star, c_star


def check_in_exponent(pk: ProvingKey, proof, z_inst, z_wit, h, r_rand, s_rand):
    """True iff proof == (a* G1, b* G2, c* G1)."""
    G1, G2 = groups(pk.curve)
    a_star, b_star, c_star = expected_proof_exponents(pk, z_inst, z_wit, h, r_rand, s_rand)
    A, B, C = proof[:3]
    return A == G1.mul(G1.gen, a_star) and B == G2.mul(G2.gen, b_star) and C == G1.mul(G1.gen, c_star)


def verify_equation_in_exponent(pk: ProvingKey, z_inst, a_star, b_star, c_star):
    """Groth16 verification equation e(A,B) = e(alpha,beta) e(IC,gamma) e(C,delta), in the exponent."""
    td, r = pk.trapdoor, pk.curve.r
    ginv = pow(td.gamma, -1, r)
    ic = sum(
        z_inst[j] * (td.beta * pk.a_tau[j] + td.alpha * pk.b_tau[j] + pk.c_tau[j]) % r * ginv
        for j in range(len(z_inst))
    ) % r
    return (a_star * b_star - td.alpha * td.beta - ic * td.gamma - c_star * td.delta) % r == 0
