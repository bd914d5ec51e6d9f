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
