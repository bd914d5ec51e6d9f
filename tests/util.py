"""Shared helpers for the tests: limb packing between Python ints and the C-ABI layout."""
import ctypes

import numpy as np


def pack_u32(xs, n):
    """ints -> contiguous uint32 array of len(xs)*n little-endian limbs."""
    arr = np.zeros(len(xs) * n, dtype=np.uint32)
    for i, x in enumerate(xs):
        for j in range(n):
            arr[i * n + j] = (x >> (32 * j)) & 0xFFFFFFFF
    return arr


def unpack_u32(arr, n):
    arr = np.asarray(arr, dtype=np.uint32).reshape(-1, n)
    return [sum(int(row[j]) << (32 * j) for j in range(n)) for row in arr]


def ptr(a, ty=ctypes.c_uint32):
    return a.ctypes.data_as(ctypes.POINTER(ty))


# ---- points <-> limb arrays (Montgomery form, affine (x, y), infinity = all-zero) --------------
def fq_limbs(curve):
    return curve.fq_limbs64 * 2


def pack_points(curve, group, pts):
    """Affine oracle points (ints or Fq2 tuples, None = infinity) -> uint32 array in the C-ABI layout."""
    n = fq_limbs(curve)
    R = 1 << (32 * n)
    p = curve.p
    flat = []
    for P in pts:
        if P is None:
            flat += [0] * (2 * group)
        elif group == 1:
            flat += [P[0] * R % p, P[1] * R % p]
        else:
            flat += [P[0][0] * R % p, P[0][1] * R % p, P[1][0] * R % p, P[1][1] * R % p]
    return pack_u32(flat, n)


def unpack_points(curve, group, arr):
    n = fq_limbs(curve)
    Rinv = pow(1 << (32 * n), -1, curve.p)
    vals = [v * Rinv % curve.p for v in unpack_u32(arr, n)]
    out = []
    step = 2 * group
    for i in range(0, len(vals), step):
        c = vals[i : i + step]
        if all(v == 0 for v in c):
            out.append(None)
        elif group == 1:
            out.append((c[0], c[1]))
        else:
            out.append(((c[0], c[1]), (c[2], c[3])))
    return out


def pack_fr(curve, xs, mont=True):
    R = 1 << 256
    return pack_u32([(x * R % curve.r) if mont else x % curve.r for x in xs], 8)


def unpack_fr(curve, arr, mont=True):
    Rinv = pow(1 << 256, -1, curve.r)
    return [(v * Rinv % curve.r) if mont else v for v in unpack_u32(arr, 8)]
