"""G2 elements of the KZG10 keys on the host: the two or three G2Affine points of `UniversalParams` / `VerifierKey`
(kzg10/data_structures.rs:30-35 `h`, `beta_h`, `neg_powers_of_h`; :204-207) are (de)serialised and validated here, in
plain Python integers -- they are a handful of points per key, while the 2^20+ G1 powers next to them go through the GPU
decoder (csrc/wire.cuh).  Pairing-side use of these points stays with the caller (out of scope).

Encodings (un-vendored crates, restated from their published behaviour):
  * ark-bls12-381 (ZCash form, curves/util.rs): 96 / 192 bytes big-endian, x.c1 || x.c0 [|| y.c1 || y.c0], flags in the
    three top bits of byte 0 (compressed, infinity, y lexicographically largest); canonical-only like G1.
  * generic short-Weierstrass over Fq2 (BN254): Fp2 = c0 || c1, each little-endian, SWFlags in the top bits of the LAST
    byte of the flagged element (x when compressed, y otherwise).
  `y > -y` compares c1 first, then c0 (QuadExtField's Ord).
ABI layout of a G2 point: x.c0 || x.c1 || y.c0 || y.c1, each `limbs` u64 Montgomery (== ark's in-memory Fp2 { c0, c1 }).
"""
import numpy as np

from .binding import BLS12_381, BN254, fq_limbs
from .params import FQ_MODULUS, FR_MODULUS

BAD_FLAGS, NOT_CANONICAL, NOT_ON_CURVE, NOT_IN_SUBGROUP = 1, 2, 3, 4


class G2WireError(ValueError):
    def __init__(self, reason):
        super().__init__(f"G2 element failed to decode / validate (reason {reason})")
        self.reason = reason


# ---- Fq2 = Fq[u] / (u^2 + 1) ----------------------------------------------------------------------------------------------
def f2_add(a, b, p): return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)
def f2_sub(a, b, p): return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)
def f2_neg(a, p): return ((-a[0]) % p, (-a[1]) % p)
def f2_mul(a, b, p): return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)
def f2_sqr(a, p): return f2_mul(a, a, p)


def f2_inv(a, p):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
    return (a[0] * n % p, (-a[1]) * n % p)


def f2_pow(a, e, p):
    acc = (1, 0)
    while e:
        if e & 1:
            acc = f2_mul(acc, a, p)
        a = f2_sqr(a, p)
        e >>= 1
    return acc


def f2_sqrt(a, p):
    """a square root in Fq2 (p = 3 mod 4) or None -- Adj / Rodriguez-Henriquez, algorithm 9"""
    if a == (0, 0):
        return a
    a1 = f2_pow(a, (p - 3) // 4, p)
    alpha = f2_mul(f2_sqr(a1, p), a, p)
    x0 = f2_mul(a1, a, p)
    if alpha == (p - 1, 0):
        x = (-x0[1] % p, x0[0])                                  # u * x0
    else:
        b = f2_pow(f2_add((1, 0), alpha, p), (p - 1) // 2, p)
        x = f2_mul(b, x0, p)
    return x if f2_sqr(x, p) == a else None


def f2_is_larger(y, p):
    """y > -y in QuadExtField's order (c1 first, then c0)"""
    ny = f2_neg(y, p)
    return (y[1], y[0]) > (ny[1], ny[0])


def twist_b(curve):
    p = FQ_MODULUS[curve]
    if curve == BLS12_381:
        return (4, 4)                                            # 4 (1 + u)
    if curve == BN254:
        return f2_mul((3, 0), f2_inv((9, 1), p), p)              # 3 / (9 + u)
    raise ValueError("curve has no pairing / G2")


# ---- group law on the twist (affine, None = identity) -----------------------------------------------------------------------
def g2_add(curve, P, Q):
    p = FQ_MODULUS[curve]
    if P is None: return Q
    if Q is None: return P
    (x1, y1), (x2, y2) = P, Q
    if x1 == x2:
        if f2_add(y1, y2, p) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_sqr(x1, p), p), f2_inv(f2_add(y1, y1, p), p), p)
    else:
        lam = f2_mul(f2_sub(y2, y1, p), f2_inv(f2_sub(x2, x1, p), p), p)
    x3 = f2_sub(f2_sub(f2_sqr(lam, p), x1, p), x2, p)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3, p), p), y1, p))


def g2_mul(curve, k, P):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(curve, acc, acc)
        if bit == "1":
            acc = g2_add(curve, acc, P)
    return acc


def g2_on_curve(curve, P):
    if P is None:
        return True
    p = FQ_MODULUS[curve]
    x, y = P
    return f2_sqr(y, p) == f2_add(f2_mul(f2_sqr(x, p), x, p), twist_b(curve), p)


def g2_check(curve, P):
    """Valid::check of a G2Affine: on the curve and in the prime-order subgroup (r * P = O)"""
    if not g2_on_curve(curve, P):
        raise G2WireError(NOT_ON_CURVE)
    if P is not None and g2_mul(curve, FR_MODULUS[curve], P) is not None:
        raise G2WireError(NOT_IN_SUBGROUP)


def g2_wire_size(curve, compressed=True):
    if curve == BLS12_381:
        return 96 if compressed else 192
    if curve == BN254:
        return 64 if compressed else 128
    raise ValueError("curve has no pairing / G2")


def g2_serialize(curve, P, compressed=True):
    p = FQ_MODULUS[curve]
    (x, y) = ((0, 0), (0, 0)) if P is None else P
    large = P is not None and f2_is_larger(y, p)
    if curve == BLS12_381:
        b = bytearray(x[1].to_bytes(48, "big") + x[0].to_bytes(48, "big"))
        if not compressed:
            b += y[1].to_bytes(48, "big") + y[0].to_bytes(48, "big")
        b[0] |= (0x80 if compressed else 0) | (0x40 if P is None else 0) | (0x20 if compressed and large else 0)
        return bytes(b)
    flags = 0x40 if P is None else (0x80 if large else 0)
    b = bytearray(x[0].to_bytes(32, "little") + x[1].to_bytes(32, "little"))
    if not compressed:
        b += y[0].to_bytes(32, "little") + y[1].to_bytes(32, "little")
    b[-1] |= flags
    return bytes(b)


def g2_deserialize(curve, data, compressed=True, validate=True):
    """bytes -> affine point ((x0, x1), (y0, y1)) or None; raises G2WireError"""
    p = FQ_MODULUS[curve]
    sz = g2_wire_size(curve, compressed)
    b = bytearray(data[:sz])
    if len(b) != sz:
        raise ValueError("truncated G2 element")
    if curve == BLS12_381:
        fc, inf, yflag = bool(b[0] & 0x80), bool(b[0] & 0x40), bool(b[0] & 0x20)
        if fc != bool(compressed) or (yflag and (not fc or inf)):
            raise G2WireError(BAD_FLAGS)
        b[0] &= 0x1F
        if inf:
            if any(b):
                raise G2WireError(NOT_CANONICAL)
            return None
        x = (int.from_bytes(b[48:96], "big"), int.from_bytes(b[0:48], "big"))
        y = None if compressed else (int.from_bytes(b[144:192], "big"), int.from_bytes(b[96:144], "big"))
    else:
        neg, inf = bool(b[-1] & 0x80), bool(b[-1] & 0x40)
        if neg and inf:
            raise G2WireError(BAD_FLAGS)
        b[-1] &= 0x3F
        yflag = neg
        x = (int.from_bytes(b[0:32], "little"), int.from_bytes(b[32:64], "little"))
        y = None if compressed else (int.from_bytes(b[64:96], "little"), int.from_bytes(b[96:128], "little"))
    if max(x) >= p or (y is not None and max(y) >= p):
        raise G2WireError(NOT_CANONICAL)
    if inf:
        return None
    if compressed:
        r = f2_sqrt(f2_add(f2_mul(f2_sqr(x, p), x, p), twist_b(curve), p), p)
        if r is None:
            raise G2WireError(NOT_ON_CURVE)
        y = r if f2_is_larger(r, p) == yflag else f2_neg(r, p)
    P = (x, y)
    if validate:
        g2_check(curve, P)
    return P


# ---- packed ABI layout ----------------------------------------------------------------------------------------------------
def g2_to_limbs(curve, P):
    """affine point -> ((4 * limbs,) uint64 Montgomery x.c0 || x.c1 || y.c0 || y.c1, is_identity)"""
    n, p = fq_limbs(curve), FQ_MODULUS[curve]
    out = np.zeros(4 * n, dtype=np.uint64)
    if P is None:
        return out, True
    R = (1 << (64 * n)) % p
    for k, c in enumerate((P[0][0], P[0][1], P[1][0], P[1][1])):
        v = c * R % p
        for j in range(n):
            out[k * n + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out, False


def g2_from_limbs(curve, limbs, is_identity=False):
    if is_identity:
        return None
    n, p = fq_limbs(curve), FQ_MODULUS[curve]
    Rinv = pow((1 << (64 * n)) % p, -1, p)
    limbs = np.asarray(limbs, dtype=np.uint64).reshape(-1)
    c = [sum(int(limbs[k * n + j]) << (64 * j) for j in range(n)) * Rinv % p for k in range(4)]
    return ((c[0], c[1]), (c[2], c[3]))
