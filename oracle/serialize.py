"""Compressed point / proof encodings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates ark-serialize's `serialize_compressed` as instantiated by ark-bls12-381 (zcash/IETF form) and ark-bn254
(ark-ec `SWFlags`), as recalled in SURVEY.md Appendix A.7 (crates not in /root/reference).  Pinned by the
well-known compressed BLS12-381 generators (tests/test_oracle_py.py); BN254 is pinned only by the rule itself.
"""
from .params import Curve


def _larger(curve: Curve, y):
    p = curve.p
    if isinstance(y, tuple):
        n = ((-y[0]) % p, (-y[1]) % p)
        return (y[1], y[0]) > (n[1], n[0])
    return y > (-y) % p


def point_compressed(curve: Curve, group: int, P) -> bytes:
    fq = 48 if curve.name == "bls12_381" else 32
    n = fq * group
    if curve.name == "bls12_381":
        if P is None:
            return bytes([0xC0]) + bytes(n - 1)
        x = P[0]
        body = x.to_bytes(fq, "big") if group == 1 else x[1].to_bytes(fq, "big") + x[0].to_bytes(fq, "big")
        first = body[0] | 0x80 | (0x20 if _larger(curve, P[1]) else 0)
        return bytes([first]) + body[1:]
    if P is None:
        return bytes(n - 1) + bytes([0x40])
    x = P[0]
    body = x.to_bytes(fq, "little") if group == 1 else x[0].to_bytes(fq, "little") + x[1].to_bytes(fq, "little")
    last = body[-1] | (0x80 if _larger(curve, P[1]) else 0)
    return body[:-1] + bytes([last])


def proof_compressed(curve: Curve, A, B, C) -> bytes:
    return point_compressed(curve, 1, A) + point_compressed(curve, 2, B) + point_compressed(curve, 1, C)


# ---------------------------------------------------------------------------------------------
# Uncompressed form, Vec<T> framing, VerifyingKey / ProvingKey, and the way back (CanonicalDeserialize)
# ---------------------------------------------------------------------------------------------
def point_uncompressed(curve: Curve, group: int, P) -> bytes:
    """`serialize_uncompressed`: BLS12-381 zcash form x || y big-endian (Fq2 as c1 || c0), flag bits in the first byte
    (compression bit clear, 0x40 = infinity); BN254 ark-ec form x || y little-endian (Fq2 as c0 || c1) with the SWFlags
    (0x80 = y is the larger root, 0x40 = infinity) in the last byte of y."""
    fq = 48 if curve.name == "bls12_381" else 32
    n = 2 * fq * group
    if curve.name == "bls12_381":
        if P is None:
            return bytes([0x40]) + bytes(n - 1)
        enc = (lambda v: v.to_bytes(fq, "big")) if group == 1 else (lambda v: v[1].to_bytes(fq, "big") + v[0].to_bytes(fq, "big"))
        return enc(P[0]) + enc(P[1])
    if P is None:
        return bytes(n - 1) + bytes([0x40])
    enc = (lambda v: v.to_bytes(fq, "little")) if group == 1 else (lambda v: v[0].to_bytes(fq, "little") + v[1].to_bytes(fq, "little"))
    body = enc(P[0]) + enc(P[1])
    return body[:-1] + bytes([body[-1] | (0x80 if _larger(curve, P[1]) else 0)])


def _sqrt_fq(p, a):
    """p = 3 mod 4 for both curves."""
    r = pow(a, (p + 1) // 4, p)
    return r if r * r % p == a % p else None


def _sqrt_fq2(p, a):
    """Square root in Fq[u]/(u^2 + 1) by the norm method."""
    a0, a1 = a[0] % p, a[1] % p
    if a1 == 0:
        r = _sqrt_fq(p, a0)
        if r is not None:
            return (r, 0)
        r = _sqrt_fq(p, (-a0) % p)       # sqrt(-a0) * u
        return None if r is None else (0, r)
    alpha = _sqrt_fq(p, (a0 * a0 + a1 * a1) % p)
    if alpha is None:
        return None
    inv2 = pow(2, -1, p)
    for s in (alpha, (-alpha) % p):
        delta = (a0 + s) * inv2 % p
        x0 = _sqrt_fq(p, delta)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, p) % p
            if ((x0 * x0 - x1 * x1) % p, 2 * x0 * x1 % p) == (a0, a1):
                return (x0, x1)
    return None


def point_decompress(curve: Curve, group: int, data: bytes):
    """Inverse of `point_compressed`; raises ValueError on malformed input (x not on the curve, bad flags, x >= p)."""
    p = curve.p
    fq = 48 if curve.name == "bls12_381" else 32
    if len(data) != fq * group:
        raise ValueError("length")
    if curve.name == "bls12_381":
        flags, body = data[0] & 0xE0, bytes([data[0] & 0x1F]) + data[1:]
        if not flags & 0x80:
            raise ValueError("compression bit not set")
        if flags & 0x40:
            if flags & 0x20 or any(body):
                raise ValueError("non-canonical infinity")
            return None
        sign = bool(flags & 0x20)
        x = int.from_bytes(body, "big") if group == 1 else (int.from_bytes(body[fq:], "big"), int.from_bytes(body[:fq], "big"))
    else:
        flags, body = data[-1] & 0xC0, data[:-1] + bytes([data[-1] & 0x3F])
        if flags == 0xC0:
            raise ValueError("both flags set")
        if flags & 0x40:
            if any(body):
                raise ValueError("non-canonical infinity")
            return None
        sign = bool(flags & 0x80)
        x = int.from_bytes(body, "little") if group == 1 else (int.from_bytes(body[:fq], "little"), int.from_bytes(body[fq:], "little"))
    if group == 1:
        if x >= p:
            raise ValueError("x >= p")
        y = _sqrt_fq(p, (x * x * x + curve.b) % p)
    else:
        if x[0] >= p or x[1] >= p:
            raise ValueError("x >= p")
        x2 = ((x[0] * x[0] - x[1] * x[1]) % p, 2 * x[0] * x[1] % p)
        x3 = ((x2[0] * x[0] - x2[1] * x[1]) % p, (x2[0] * x[1] + x2[1] * x[0]) % p)
        y = _sqrt_fq2(p, ((x3[0] + curve.b2[0]) % p, (x3[1] + curve.b2[1]) % p))
    if y is None:
        raise ValueError("x is not on the curve")
    neg = (-y) % p if group == 1 else ((-y[0]) % p, (-y[1]) % p)
    if _larger(curve, y) != sign:
        y = neg
    return (x, y)


def vec_framed(items) -> bytes:
    """ark-serialize `Vec<T>`: u64 little-endian length, then the elements."""
    items = list(items)
    return len(items).to_bytes(8, "little") + b"".join(items)


def verifying_key_bytes(curve: Curve, vk, compressed=True) -> bytes:
    """ark-groth16 `VerifyingKey { alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1: Vec<G1Affine> }` (field order =
    derive order; recalled, crate not in /root/reference).  vk: dict with those keys (affine oracle points)."""
    enc = point_compressed if compressed else point_uncompressed
    return (enc(curve, 1, vk["alpha_g1"]) + enc(curve, 2, vk["beta_g2"]) + enc(curve, 2, vk["gamma_g2"]) + enc(curve, 2, vk["delta_g2"])
            + vec_framed(enc(curve, 1, P) for P in vk["gamma_abc_g1"]))


def proving_key_bytes(curve: Curve, pk, compressed=True) -> bytes:
    """ark-groth16 `ProvingKey { vk, beta_g1, delta_g1, a_query, b_g1_query, b_g2_query, h_query, l_query }`.
    pk: oracle.groth16.ProvingKey."""
    enc = point_compressed if compressed else point_uncompressed
    vk = {"alpha_g1": pk.alpha_g1, "beta_g2": pk.beta_g2, "gamma_g2": pk.gamma_g2, "delta_g2": pk.delta_g2, "gamma_abc_g1": pk.gamma_abc_g1}
    out = verifying_key_bytes(curve, vk, compressed) + enc(curve, 1, pk.beta_g1) + enc(curve, 1, pk.delta_g1)
    for group, q in ((1, pk.a_query), (1, pk.b_g1_query), (2, pk.b_g2_query), (1, pk.h_query), (1, pk.l_query)):
        out += vec_framed(enc(curve, group, P) for P in q)
    return out


def proof_decompress(curve: Curve, data: bytes):
    fq = 48 if curve.name == "bls12_381" else 32
    if len(data) != 4 * fq:
        raise ValueError("length")
    return (point_decompress(curve, 1, data[:fq]), point_decompress(curve, 2, data[fq:3 * fq]), point_decompress(curve, 1, data[3 * fq:]))
