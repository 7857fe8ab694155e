"""TEST INFRASTRUCTURE -- ctypes loader for oracle/liboracle.so (the C oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FQ_LIMBS = {0: 6, 1: 4, 2: 4}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def num_threads():
    return lib().orc_num_threads()


def field_binop(name, curve, a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    n = a.size // (FQ_LIMBS[curve] if "fq" in name else 4)
    rc = getattr(lib(), name)(curve, _p(a), _p(b), _p(out), ctypes.c_size_t(n))
    assert rc == 0
    return out


def field_unop(name, curve, a):
    a = _u64(a)
    out = np.empty_like(a)
    n = a.size // (FQ_LIMBS[curve] if "fq" in name else 4)
    rc = getattr(lib(), name)(curve, _p(a), _p(out), ctypes.c_size_t(n))
    assert rc == 0
    return out


def g1_generator(curve):
    out = np.zeros(2 * FQ_LIMBS[curve], dtype=np.uint64)
    assert lib().orc_g1_generator(curve, _p(out)) == 0
    return out


def g1_on_curve(curve, xy, inf=None):
    xy = _u64(xy)
    n = xy.size // (2 * FQ_LIMBS[curve])
    inf = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
    return lib().orc_g1_on_curve(curve, _p(xy), _p(inf), ctypes.c_size_t(n))


def g1_mul(curve, p_xy, k, p_inf=None):
    p_xy, k = _u64(p_xy), _u64(k)
    out = np.zeros(2 * FQ_LIMBS[curve], dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    pinf = None if p_inf is None else np.ascontiguousarray(p_inf, dtype=np.uint8)
    assert lib().orc_g1_mul(curve, _p(p_xy), _p(pinf), _p(k), _p(out), _p(oinf)) == 0
    return out, int(oinf[0])


def g1_sum(curve, xy, inf=None):
    xy = _u64(xy)
    n = xy.size // (2 * FQ_LIMBS[curve])
    inf = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
    out = np.zeros(2 * FQ_LIMBS[curve], dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    assert lib().orc_g1_sum(curve, _p(xy), _p(inf), ctypes.c_size_t(n), _p(out), _p(oinf)) == 0
    return out, int(oinf[0])


def g1_fold(curve, key_xy, chal_canonical):
    key_xy, chal = _u64(key_xy), _u64(chal_canonical)
    m = key_xy.size // (2 * FQ_LIMBS[curve]) // 2
    out = np.zeros((m, 2 * FQ_LIMBS[curve]), dtype=np.uint64)
    assert lib().orc_g1_fold(curve, _p(key_xy), ctypes.c_size_t(m), _p(chal), _p(out)) == 0
    return out


def msm(curve, bases, scalars, inf=None, n=None, naive=False, nthreads=0):
    """bases (m, 2*nq) Montgomery, scalars (n, 4) CANONICAL.  Returns (xy, inf_flag)."""
    bases, scalars = _u64(bases), _u64(scalars)
    if n is None:
        n = min(bases.size // (2 * FQ_LIMBS[curve]), scalars.size // 4)
    inf = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
    out = np.zeros(2 * FQ_LIMBS[curve], dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    if naive:
        rc = lib().orc_msm_naive(curve, _p(bases), _p(inf), _p(scalars), ctypes.c_size_t(n), _p(out), _p(oinf))
    else:
        rc = lib().orc_msm_pippenger(curve, _p(bases), _p(inf), _p(scalars), ctypes.c_size_t(n), _p(out), _p(oinf), nthreads)
    assert rc == 0
    return out, int(oinf[0])


def fixed_base_batch_mul(curve, base_xy, scalars, nthreads=0):
    base_xy, scalars = _u64(base_xy), _u64(scalars)
    n = scalars.size // 4
    out = np.zeros((n, 2 * FQ_LIMBS[curve]), dtype=np.uint64)
    oinf = np.zeros(n, dtype=np.uint8)
    assert lib().orc_fixed_base_batch_mul(curve, _p(base_xy), _p(scalars), ctypes.c_size_t(n), _p(out), _p(oinf), nthreads) == 0
    return out, oinf


def fr_powers_canonical(curve, beta_mont, n):
    beta_mont = _u64(beta_mont)
    out = np.zeros((n, 4), dtype=np.uint64)
    assert lib().orc_fr_powers_canonical(curve, _p(beta_mont), ctypes.c_size_t(n), _p(out)) == 0
    return out


def fr_axpy(curve, y, c, x):
    y = _u64(y).copy(); c, x = _u64(c), _u64(x)
    assert lib().orc_fr_axpy(curve, _p(y), _p(c), _p(x), ctypes.c_size_t(x.size // 4)) == 0
    return y


def fr_div_linear(curve, p, z):
    p, z = _u64(p), _u64(z)
    n = p.size // 4
    q = np.zeros((max(n - 1, 0), 4), dtype=np.uint64)
    rem = np.zeros(4, dtype=np.uint64)
    assert lib().orc_fr_div_linear(curve, _p(p), ctypes.c_size_t(n), _p(z), _p(q), _p(rem)) == 0
    return q, rem


def fr_eval(curve, p, z):
    p, z = _u64(p), _u64(z)
    out = np.zeros(4, dtype=np.uint64)
    assert lib().orc_fr_eval(curve, _p(p), ctypes.c_size_t(p.size // 4), _p(z), _p(out)) == 0
    return out


def fr_inner_product(curve, a, b):
    a, b = _u64(a), _u64(b)
    out = np.zeros(4, dtype=np.uint64)
    assert lib().orc_fr_inner_product(curve, _p(a), _p(b), ctypes.c_size_t(a.size // 4), _p(out)) == 0
    return out


def fr_row_mul(curve, v, m, rows, cols):
    v, m = _u64(v), _u64(m)
    out = np.zeros((cols, 4), dtype=np.uint64)
    assert lib().orc_fr_row_mul(curve, _p(v), _p(m), ctypes.c_size_t(rows), ctypes.c_size_t(cols), _p(out)) == 0
    return out


def fr_domain_generator(curve, logn):
    out = np.zeros(4, dtype=np.uint64)
    assert lib().orc_fr_domain_generator(curve, logn, _p(out)) == 0
    return out


def fr_ntt(curve, coeffs, logn, naive=False):
    coeffs = _u64(coeffs)
    out = np.zeros((1 << logn, 4), dtype=np.uint64)
    fn = lib().orc_fr_ntt_naive if naive else lib().orc_fr_ntt
    assert fn(curve, _p(coeffs), ctypes.c_size_t(coeffs.size // 4), logn, _p(out)) == 0
    return out


def kzg_commit(curve, powers_of_g, coeffs, powers_of_gamma_g=None, blind=None, nthreads=0):
    powers_of_g, coeffs = _u64(powers_of_g), _u64(coeffs)
    nq = FQ_LIMBS[curve]
    pg = None if powers_of_gamma_g is None else _u64(powers_of_gamma_g)
    bl = None if blind is None else _u64(blind)
    out = np.zeros(2 * nq, dtype=np.uint64); oinf = np.zeros(1, dtype=np.uint8)
    rc = lib().orc_kzg_commit(curve, _p(powers_of_g), ctypes.c_size_t(powers_of_g.size // (2 * nq)), _p(coeffs),
                              ctypes.c_size_t(coeffs.size // 4), _p(pg), ctypes.c_size_t(0 if pg is None else pg.size // (2 * nq)),
                              _p(bl), ctypes.c_size_t(0 if bl is None else bl.size // 4), _p(out), _p(oinf), nthreads)
    return rc, out, int(oinf[0])


def kzg_open(curve, powers_of_g, coeffs, z, powers_of_gamma_g=None, blind=None, nthreads=0):
    powers_of_g, coeffs, z = _u64(powers_of_g), _u64(coeffs), _u64(z)
    nq = FQ_LIMBS[curve]
    pg = None if powers_of_gamma_g is None else _u64(powers_of_gamma_g)
    bl = None if blind is None else _u64(blind)
    out = np.zeros(2 * nq, dtype=np.uint64); oinf = np.zeros(1, dtype=np.uint8)
    rv = np.zeros(4, dtype=np.uint64)
    rc = lib().orc_kzg_open(curve, _p(powers_of_g), ctypes.c_size_t(powers_of_g.size // (2 * nq)), _p(coeffs),
                            ctypes.c_size_t(coeffs.size // 4), _p(z), _p(pg),
                            ctypes.c_size_t(0 if pg is None else pg.size // (2 * nq)),
                            _p(bl), ctypes.c_size_t(0 if bl is None else bl.size // 4), _p(out), _p(oinf), _p(rv), nthreads)
    return rc, out, int(oinf[0]), rv
