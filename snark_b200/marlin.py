"""Universal-setup SNARK for R1CS in the style of Marlin (BASELINE config 5): algebraic holographic
proof over the domains H (constraints / variables) and K (matrix non-zeros) + KZG10 polynomial commitments.

What it stands in for.  The reference tree only declares the interface -- trait `UniversalSetupSNARK`
(/root/reference/snark/src/lib.rs:107-133: `universal_setup(compute_bound, rng) -> PublicParameters`,
`index(pp, circuit, rng) -> (ProvingKey, VerifyingKey)`) on top of `SNARK::prove / verify` (lib.rs:50-75).  The
implementations (ark-marlin, ark-poly-commit) are not in the tree and cannot be built here, so NOTHING below is byte-compatible
with them: this is the same *shape* of computation (the AHP of the Marlin paper, section 5: one univariate sum-check over H
for the lincheck with z_C = z_A z_B folded in, one over K for the rational function of the matrix arithmetisation; KZG10 with
shifted-power degree bounds), WITHOUT zero-knowledge masking and with its own Fiat-Shamir transcript (SHA-256).  Parity is
therefore "unpinned" by construction; correctness is pinned by an independent verifier (oracle/marlin.py: AHP identities +
opening checks, with real pairings on small cases) and by bit-equality of the GPU prover with the big-int prover.

How the work splits (this is ark-marlin's own split: the AHP is generic over a polynomial-commitment backend and an
evaluation domain).  Host, here: the protocol logic -- which polynomial is built from which, transcript, challenges.
Device, through the C ABI (include/b200snark.h): every O(|H|) / O(|K|) operation -- `b2s_ntt` (all domain changes),
`b2s_msm_g1` (every commitment and opening proof), `b2s_fixed_base_g1` (the SRS), `b2s_spmv` on the R1CS handle (z_A, z_B and,
on the transposed matrices, t = sum_M eta_M M^T r_alpha), and the element-wise `b2s_poly_*` kernels (products, batched inverses,
geometric sequences, evaluations).  Vectors stay on the device between calls; only commitments (64-96 B), evaluations and
challenges cross PCIe.

The functions take a `be` object (vector backend).  `GpuBackend` (marlin_gpu.py) is the product; tests also run the same
protocol over the big-int backend of oracle/marlin.py to pin the GPU kernels bit for bit.
"""
import hashlib
from dataclasses import dataclass, field
from typing import Any, List, Tuple


def next_pow2(x: int) -> int:
    n = 1
    while n < x:
        n *= 2
    return n


def log2(n: int) -> int:
    assert n & (n - 1) == 0 and n > 0
    return n.bit_length() - 1


# ---- Fiat-Shamir transcript (SHA-256 chain; NOT ark-marlin's Blake2s/ChaCha FiatShamirRng) ---------------------------------
class Transcript:
    def __init__(self, r: int, label: bytes):
        self.r = r
        self.state = hashlib.sha256(b"b200-snark/marlin-style/v1:" + label).digest()

    def absorb(self, data: bytes):
        self.state = hashlib.sha256(self.state + len(data).to_bytes(8, "little") + data).digest()

    def absorb_ints(self, xs):
        nb = (self.r.bit_length() + 7) // 8
        self.absorb(b"".join(int(x).to_bytes(nb, "little") for x in xs))

    def absorb_points(self, pts, fq_bytes):
        """Affine G1 points as (x, y) canonical little-endian; infinity (None) as zeros."""
        out = []
        for P in pts:
            if P is None:
                out.append(bytes(2 * fq_bytes))
            else:
                out.append(int(P[0]).to_bytes(fq_bytes, "little") + int(P[1]).to_bytes(fq_bytes, "little"))
        self.absorb(b"".join(out))

    def challenge(self) -> int:
        """Non-zero field element from 64 bytes of output (bias < 2^-250)."""
        a = hashlib.sha256(self.state + b"\x01").digest()
        b = hashlib.sha256(self.state + b"\x02").digest()
        self.state = hashlib.sha256(self.state + b"\x03").digest()
        v = int.from_bytes(a + b, "little") % self.r
        return v if v else 1


# ---- index ------------------------------------------------------------------------------------------------------------
@dataclass
class IndexInfo:
    """Public shape of an indexed circuit."""
    n_rows: int          # constraints
    n_inst: int          # instance variables incl. the constant ONE (relations/src/gr1cs/constraint_system.rs:121)
    n_vars: int          # instance + witness
    n: int               # |H|
    m: int               # |K|
    l: int               # |H_X| = next_pow2(n_inst): the instance sits on the subgroup of H of this size
    D: int               # largest committed degree (SRS needs D + 1 powers)


def variable_positions(n_inst: int, n_vars: int, n: int, l: int) -> List[int]:
    """Index in H of every R1CS column (variable_index order: instance first, relations/src/utils/variable.rs:105-113).
    Instance variable j sits at j * (n / l) -- the points of the order-l subgroup H_X -- so that the verifier can evaluate the
    instance part x^(X) of z^ = w^ v_X + x^ on its own; witness variables fill the remaining positions in order."""
    s = n // l
    assert s >= 2 and (n_vars - n_inst) <= n - l
    pos = [j * s for j in range(n_inst)]
    for w in range(n_vars - n_inst):
        pos.append(w + w // (s - 1) + 1)
    return pos


def index_shape(mats, n_inst: int, n_vars: int) -> IndexInfo:
    n_rows = len(mats[0])
    l = next_pow2(max(n_inst, 1))
    n = next_pow2(max(n_rows, l + (n_vars - n_inst), 2 * l))
    nnz = max(sum(len(row) for row in M) for M in mats)
    m = next_pow2(max(nnz, 2))
    return IndexInfo(n_rows=n_rows, n_inst=n_inst, n_vars=n_vars, n=n, m=m, l=l, D=max(2 * n, 3 * m))


@dataclass
class MatrixArith:
    """row / col / val* / row*col of one matrix: evaluations over K, coefficients, evaluations on the 4|K| coset."""
    k_evals: List[Any]
    coeffs: List[Any]
    c_evals: List[Any]


@dataclass
class ProverKey:
    info: IndexInfo
    pos: List[int]
    mat_handle: Any          # matrices with columns mapped to H positions  (z_A, z_B)
    mat_t_handle: Any        # their transposes                              (t = sum eta_M M^T r_alpha)
    arith: List[MatrixArith]
    index_comms: List[Tuple]  # 12 G1 points: row, col, val*, row*col of A, B, C
    srs: Any


@dataclass
class VerifierKey:
    info: IndexInfo
    index_comms: List[Tuple]


def transpose_rows(rows, n_out_rows):
    out = [[] for _ in range(n_out_rows)]
    for i, row in enumerate(rows):
        for coeff, col in row:
            out[col].append((coeff, i))
    return out


def universal_setup(be, compute_bound: int, tau: int):
    """`UniversalSetupSNARK::universal_setup(compute_bound, rng)` (snark/src/lib.rs:117-123): public parameters good for every
    circuit whose `index_shape(...).D` is at most `compute_bound` -- the KZG10 powers tau^i G1, i <= compute_bound, resident
    in the backend.  The trait draws tau from `rng`; here the caller passes it (tests and the bench need the trapdoor for the
    pairing-free opening check; a production caller draws it and forgets it)."""
    return be.setup(compute_bound + 1, tau)


def index(be, srs, mats, n_inst: int, n_vars: int):
    """`UniversalSetupSNARK::index` (snark/src/lib.rs:125-132): circuit-specific keys from the universal SRS.
    mats: A, B, C as row lists [(coeff, column)], the form `to_matrices()` exports
    (relations/src/gr1cs/constraint_system.rs:768-804)."""
    r = be.r
    info = index_shape(mats, n_inst, n_vars)
    assert be.srs_size(srs) >= info.D + 1, "SRS too small for this circuit"
    n, m = info.n, info.m
    pos = variable_positions(n_inst, n_vars, n, info.l)
    w_h = be.omega(log2(n))
    # columns mapped onto H positions
    mats_h = [[[(c % r, pos[col]) for c, col in row] for row in M] for M in mats]
    mat_handle = be.upload_matrices(mats_h, n_rows=info.n_rows, n_cols=n)
    mats_t = [transpose_rows(M, n) for M in mats_h]
    mat_t_handle = be.upload_matrices(mats_t, n_rows=n, n_cols=n)
    n_inv = pow(n, -1, r)
    arith, comms = [], []
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * w_h % r
    for M in mats_h:
        rows, cols, vals = [], [], []
        for i, row in enumerate(M):
            for c, col in row:
                if c % r == 0:
                    continue
                rows.append(pw[i])
                cols.append(pw[col])
                vals.append(c * pw[col] % r * n_inv % r)      # val* = v / u_H(col, col),  u_H(y, y) = n y^-1 on H
        pad = m - len(rows)
        rows += [1] * pad
        cols += [1] * pad
        vals += [0] * pad
        rc = [a * b % r for a, b in zip(rows, cols)]
        k_evals = [be.from_ints(v) for v in (rows, cols, vals, rc)]
        coeffs = [be.ntt(v, inverse=True) for v in k_evals]
        c_evals = [be.ntt(be.pad(c, 4 * m), coset=True) for c in coeffs]
        arith.append(MatrixArith(k_evals, coeffs, c_evals))
        comms += [be.commit(srs, c) for c in coeffs]
    pk = ProverKey(info, pos, mat_handle, mat_t_handle, arith, comms, srs)
    return pk, VerifierKey(info, comms)


# ---- prove ----------------------------------------------------------------------------------------------------------------
@dataclass
class Proof:
    comms: List[Tuple]        # w, zA, zB | t, g1, g1_shifted, h1 | g2, g2_shifted, h2
    evals1: List[int]         # at beta1: w, zA, zB, t, g1, h1
    evals2: List[int]         # at beta2: g2, h2, then row, col, val*, row*col of A, B, C
    openings: List[Tuple]     # W1 (at beta1), W2 (at beta2)
    debug: dict = field(default_factory=dict)


def start_transcript(r, fq_bytes, info: IndexInfo, index_comms, x):
    tr = Transcript(r, b"prove")
    tr.absorb_ints([info.n_rows, info.n_inst, info.n_vars, info.n, info.m, info.l, info.D])
    tr.absorb_points(index_comms, fq_bytes)
    tr.absorb_ints(x)
    return tr


def divide_by_vanishing(be, coeffs, size: int):
    """coeffs of length 4 * size -> (quotient chunks as one vector of length 3 * size, remainder of length size) for the
    divisor X^size - 1:  p = sum_k p_k X^(k size)  =>  quotient chunk j = sum_{k > j} p_k, remainder = sum_k p_k."""
    p = [be.slice(coeffs, k * size, (k + 1) * size) for k in range(4)]
    s3 = p[3]
    s2 = be.add(p[2], s3)
    s1 = be.add(p[1], s2)
    rem = be.add(p[0], s1)
    return be.concat([s1, s2, s3]), rem


def assign(be, pk: ProverKey, x: List[int], w: List[int]):
    """The full assignment z = instance || witness (constraint_system.rs:193-206) laid out over H: variable j at position
    pk.pos[j], unused positions zero.  Returns the backend vector `prove_assigned` consumes (the per-proof upload)."""
    info = pk.info
    assert len(x) == info.n_inst and len(x) + len(w) == info.n_vars
    z_h = [0] * info.n
    for j, v in enumerate(list(x) + list(w)):
        z_h[pk.pos[j]] = v % be.r
    return be.from_ints(z_h)


def prove(be, pk: ProverKey, x: List[int], w: List[int], check: bool = False) -> Proof:
    """`SNARK::prove` (snark/src/lib.rs:50-54) for the universal-setup scheme.  x: instance assignment (x[0] = 1),
    w: witness assignment (constraint_system.rs:193-206)."""
    return prove_assigned(be, pk, x, assign(be, pk, x, w), check)


def prove_assigned(be, pk: ProverKey, x: List[int], z_h, check: bool = False) -> Proof:
    r, info, srs = be.r, pk.info, pk.srs
    n, m, l = info.n, info.m, info.l
    g = be.coset_gen
    tr = start_transcript(r, be.fq_bytes, info, pk.index_comms, x)
    dbg = {}

    # ---- round 1: w^, z_A^, z_B^ -----------------------------------------------------------------------------------------
    zA_e, zB_e, _ = be.spmv(pk.mat_handle, z_h, n)                    # evaluations over H (rows >= n_rows are zero)
    x_coeffs = be.ntt(be.from_ints([v % r for v in x] + [0] * (l - len(x))), inverse=True)      # x^ over H_X, degree < l
    x_h = be.ntt(be.pad(x_coeffs, n))                                 # x^ on H
    w_h = be.omega(log2(n))
    vx_h = be.add_scalar(be.geom(n, 1, pow(w_h, l, r)), r - 1)        # v_X(w^i) = (w^l)^i - 1   (zero on H_X)
    w_e = be.mul(be.sub(z_h, x_h), be.inv0(vx_h))                     # (z - x^) / v_X off H_X, 0 on H_X
    w_c, zA_c, zB_c = (be.ntt(v, inverse=True) for v in (w_e, zA_e, zB_e))
    c_r1 = [be.commit(srs, v) for v in (w_c, zA_c, zB_c)]
    tr.absorb_points(c_r1, be.fq_bytes)
    alpha, eta_a, eta_b, eta_c = (tr.challenge() for _ in range(4))
    etas = (eta_a, eta_b, eta_c)

    # ---- round 2: t, sum-check over H ------------------------------------------------------------------------------------
    vh_alpha = (pow(alpha, n, r) - 1) % r
    assert vh_alpha != 0, "alpha landed in H"
    # r(alpha, w^i) = v_H(alpha) / (alpha - w^i)
    r_e = be.scale(be.inv0(be.add_scalar(be.scale(be.geom(n, 1, w_h), r - 1), alpha)), vh_alpha)
    tA, tB, tC = be.spmv(pk.mat_t_handle, r_e, n)                     # M^T r_alpha, per matrix
    t_e = be.add(be.add(be.scale(tA, eta_a), be.scale(tB, eta_b)), be.scale(tC, eta_c))
    t_c = be.ntt(t_e, inverse=True)
    r_c = be.ntt(r_e, inverse=True)
    n4 = 4 * n
    ext = lambda c: be.ntt(be.pad(c, n4), coset=True)                 # evaluations on the coset g <w_4n>
    r4, zA4, zB4, t4, w4, x4 = (ext(c) for c in (r_c, zA_c, zB_c, t_c, w_c, x_coeffs))
    w_4n = be.omega(log2(n4))
    vx4 = be.add_scalar(be.geom(n4, pow(g, l, r), pow(w_4n, l, r)), r - 1)
    z4 = be.add(be.mul(w4, vx4), x4)                                  # z^ = w^ v_X + x^
    lin = be.add(be.add(be.scale(zA4, eta_a), be.scale(zB4, eta_b)), be.scale(be.mul(zA4, zB4), eta_c))
    q1 = be.ntt(be.sub(be.mul(r4, lin), be.mul(t4, z4)), inverse=True, coset=True)
    h1_c, rem = divide_by_vanishing(be, q1, n)                        # q1 = h1 v_H + X g1   (the sum over H is zero)
    g1_c = be.slice(rem, 1, n)
    if check:
        assert be.to_ints(be.slice(rem, 0, 1))[0] == 0, "lincheck sum over H is not zero (unsatisfied assignment?)"
        assert not any(be.to_ints(be.slice(h1_c, 2 * n, 3 * n))), "deg h1 exceeds 2|H| - 3"
    h1_c = be.slice(h1_c, 0, 2 * n)                                   # deg q1 <= 3|H| - 3, so deg h1 <= 2|H| - 3
    sh1 = info.D - (n - 2)
    c_r2 = [be.commit(srs, t_c), be.commit(srs, g1_c), be.commit(srs, g1_c, shift=sh1), be.commit(srs, h1_c)]
    tr.absorb_points(c_r2, be.fq_bytes)
    beta1 = tr.challenge()
    vh_beta1 = (pow(beta1, n, r) - 1) % r
    assert vh_beta1 != 0, "beta1 landed in H"

    # ---- round 3: sum-check over K ---------------------------------------------------------------------------------------
    t_beta1 = be.eval(t_c, beta1)
    scale_ab = vh_alpha * vh_beta1 % r
    ab = alpha * beta1 % r

    def denominators(which):
        out = []
        for ar in pk.arith:
            row, col, _, rc = ar.k_evals if which == "k" else ar.c_evals
            d = be.add_scalar(be.sub(rc, be.add(be.scale(col, alpha), be.scale(row, beta1))), ab)
            out.append(d)                                            # (alpha - row)(beta1 - col)
        return out

    den_k = denominators("k")
    f2_e = None
    for ar, d, eta in zip(pk.arith, den_k, etas):
        term = be.scale(be.mul(ar.k_evals[2], be.inv0(d)), eta * scale_ab % r)
        f2_e = term if f2_e is None else be.add(f2_e, term)
    f2_c = be.ntt(f2_e, inverse=True)
    g2_c = be.slice(f2_c, 1, m)                                       # f2 = X g2 + t(beta1) / |K|
    if check:
        assert be.to_ints(be.slice(f2_c, 0, 1))[0] * m % r == t_beta1, "sum over K does not give t(beta1)"
    m4 = 4 * m
    dA, dB, dC = denominators("c")
    vA, vB, vC = (ar.c_evals[2] for ar in pk.arith)
    b4 = be.mul(be.mul(dA, dB), dC)
    a4 = be.add(be.add(be.scale(be.mul(vA, be.mul(dB, dC)), eta_a * scale_ab % r),
                       be.scale(be.mul(vB, be.mul(dA, dC)), eta_b * scale_ab % r)),
                be.scale(be.mul(vC, be.mul(dA, dB)), eta_c * scale_ab % r))
    f2_4 = be.ntt(be.pad(f2_c, m4), coset=True)
    num = be.ntt(be.sub(a4, be.mul(b4, f2_4)), inverse=True, coset=True)
    h2_c, rem2 = divide_by_vanishing(be, num, m)                      # a - b f2 = h2 v_K exactly
    if check:
        assert not any(be.to_ints(rem2)), "a - b f2 is not divisible by v_K"
    sh2 = info.D - (m - 2)
    c_r3 = [be.commit(srs, g2_c), be.commit(srs, g2_c, shift=sh2), be.commit(srs, h2_c)]
    tr.absorb_points(c_r3, be.fq_bytes)
    beta2 = tr.challenge()

    # ---- round 4: evaluations and batched openings ----------------------------------------------------------------------------
    polys1 = [w_c, zA_c, zB_c, t_c, g1_c, h1_c]
    evals1 = [be.eval(p, beta1) for p in polys1]
    polys2 = [g2_c, h2_c] + [c for ar in pk.arith for c in ar.coeffs]
    evals2 = [be.eval(p, beta2) for p in polys2]
    tr.absorb_ints(evals1 + evals2)
    xi1, xi2 = tr.challenge(), tr.challenge()
    # the shifted polynomials ride along: X^shift g(X) evaluates to beta^shift g(beta)
    open1 = polys1 + [be.shifted(g1_c, sh1)]
    vals1 = evals1 + [pow(beta1, sh1, r) * evals1[4] % r]
    open2 = polys2 + [be.shifted(g2_c, sh2)]
    vals2 = evals2 + [pow(beta2, sh2, r) * evals2[0] % r]
    W1 = kzg_open_batch(be, srs, open1, vals1, beta1, xi1, info.D + 1)
    W2 = kzg_open_batch(be, srs, open2, vals2, beta2, xi2, info.D + 1)
    dbg.update(alpha=alpha, etas=etas, beta1=beta1, beta2=beta2, xi=(xi1, xi2), t_beta1=t_beta1)
    return Proof(c_r1 + c_r2 + c_r3, evals1, evals2, [W1, W2], dbg)


def kzg_open_batch(be, srs, polys, vals, z: int, xi: int, size: int):
    """One KZG10 witness for p = sum_i xi^i p_i at z:  commit((p(X) - p(z)) / (X - z)).  The quotient is formed on
    evaluations (a coset of size >= deg p + 1): (p(d) - v) / (d - z), element-wise with one batched inversion."""
    r = be.r
    N = next_pow2(size)
    acc, v, c = None, 0, 1
    for p, pv in zip(polys, vals):
        term = be.scale(be.pad(p, N), c)
        acc = term if acc is None else be.add(acc, term)
        v = (v + c * pv) % r
        c = c * xi % r
    e = be.ntt(acc, coset=True)
    w_N = be.omega(log2(N))
    pts = be.geom(N, be.coset_gen, w_N)                               # the coset points g w^i
    q_e = be.mul(be.add_scalar(e, r - v), be.inv0(be.add_scalar(pts, r - z)))
    q_c = be.ntt(q_e, inverse=True, coset=True)
    return be.commit(srs, be.slice(q_c, 0, size - 1))              # deg q <= deg p - 1 <= size - 2; the rest of q_c is zero


# ---- wire form (the trait bounds Proof / VerifyingKey by CanonicalSerialize + CanonicalDeserialize, snark/src/lib.rs:25-36) -------
# The repository's own framing, not ark-marlin's (nothing to pin that to): u32 little-endian counts, G1 points as uncompressed
# affine (x, y) canonical little-endian with all-zero bytes for the point at infinity, field elements canonical little-endian.
def _put_points(pts, fq_bytes):
    out = [len(pts).to_bytes(4, "little")]
    for P in pts:
        out.append(bytes(2 * fq_bytes) if P is None else int(P[0]).to_bytes(fq_bytes, "little") + int(P[1]).to_bytes(fq_bytes, "little"))
    return b"".join(out)


def _get_points(buf, off, fq_bytes):
    n = int.from_bytes(buf[off:off + 4], "little")
    off += 4
    pts = []
    for _ in range(n):
        x = int.from_bytes(buf[off:off + fq_bytes], "little")
        y = int.from_bytes(buf[off + fq_bytes:off + 2 * fq_bytes], "little")
        pts.append(None if x == 0 and y == 0 else (x, y))
        off += 2 * fq_bytes
    return pts, off


def proof_to_bytes(proof: Proof, r: int, fq_bytes: int) -> bytes:
    fr_bytes = (r.bit_length() + 7) // 8
    ints = list(proof.evals1) + list(proof.evals2)
    return (_put_points(proof.comms, fq_bytes) + _put_points(proof.openings, fq_bytes) + len(proof.evals1).to_bytes(4, "little") +
            len(proof.evals2).to_bytes(4, "little") + b"".join(int(v).to_bytes(fr_bytes, "little") for v in ints))


def proof_from_bytes(buf: bytes, r: int, fq_bytes: int) -> Proof:
    fr_bytes = (r.bit_length() + 7) // 8
    comms, off = _get_points(buf, 0, fq_bytes)
    openings, off = _get_points(buf, off, fq_bytes)
    n1 = int.from_bytes(buf[off:off + 4], "little")
    n2 = int.from_bytes(buf[off + 4:off + 8], "little")
    off += 8
    if len(buf) != off + (n1 + n2) * fr_bytes:
        raise ValueError("proof bytes: length does not match the declared counts")
    vals = [int.from_bytes(buf[off + i * fr_bytes:off + (i + 1) * fr_bytes], "little") for i in range(n1 + n2)]
    if any(v >= r for v in vals):
        raise ValueError("proof bytes: non-canonical field element")
    return Proof(comms, vals[:n1], vals[n1:], openings)


def vk_to_bytes(vk: VerifierKey, fq_bytes: int) -> bytes:
    i = vk.info
    head = b"".join(int(v).to_bytes(8, "little") for v in (i.n_rows, i.n_inst, i.n_vars, i.n, i.m, i.l, i.D))
    return head + _put_points(vk.index_comms, fq_bytes)


def vk_from_bytes(buf: bytes, fq_bytes: int) -> VerifierKey:
    f = [int.from_bytes(buf[8 * k:8 * k + 8], "little") for k in range(7)]
    comms, off = _get_points(buf, 56, fq_bytes)
    if off != len(buf):
        raise ValueError("verifier key bytes: trailing data")
    return VerifierKey(IndexInfo(*f), comms)
