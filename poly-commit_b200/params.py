"""Public curve parameters of the three curves BASELINE.json names (what ark-bls12-381 / ark-bn254 / ark-pallas 0.5.0
instantiate): field moduli and the G1 generator, plus the packed Montgomery limb form the C ABI uses."""
import numpy as np

from .binding import BLS12_381, BN254, PALLAS

FQ_MODULUS = {
    BLS12_381: 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    BN254: 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    PALLAS: 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001,
}
FR_MODULUS = {
    BLS12_381: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    BN254: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    PALLAS: 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001,
}
G1_GENERATOR = {
    BLS12_381: (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
                0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    BN254: (1, 2),
    PALLAS: (FQ_MODULUS[PALLAS] - 1, 2),
}


def _limbs(v, n):
    return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(n)], dtype=np.uint64)


def fq_mont(curve, v):
    p = FQ_MODULUS[curve]
    n = (p.bit_length() + 63) // 64
    return _limbs(v % p * (1 << (64 * n)) % p, n)


def fr_mont(curve, v):
    r = FR_MODULUS[curve]
    return _limbs(v % r * (1 << 256) % r, 4)


def g1_generator(curve):
    """affine generator as Montgomery x||y limbs (the ABI's point layout)"""
    x, y = G1_GENERATOR[curve]
    return np.concatenate([fq_mont(curve, x), fq_mont(curve, y)])


def random_fr(curve, n, seed, bits=None):
    """(n, 4) uint64 values uniform below 2^bits (default: the largest power of two below r) -- valid field elements in
    either representation; synthetic polynomial coefficients / scalars for benchmarks"""
    r = FR_MODULUS[curve]
    bits = bits if bits is not None else r.bit_length() - 1
    g = np.random.default_rng(seed)
    out = g.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    top = bits - 192
    out[:, 3] &= np.uint64((1 << top) - 1)
    return out
