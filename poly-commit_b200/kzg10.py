"""Host mirror of the G1 side of KZG10's verifier (poly-commit/src/kzg10/mod.rs) over the C ABI -- SURVEY.md section 8f
rank 2 ("verifier-side combination MSMs").  Pairings stay with the caller (out of scope): these functions
return exactly the G1 points the reference feeds to `E::pairing` / `E::multi_pairing`.

  check         kzg10/mod.rs:314-333   inner = comm - g * value - gamma_g * random_v          (:322-325)
  batch_check   kzg10/mod.rs:337-391   total_c = sum r_i (c_i + z_i w_i) - g * sum r_i v_i - gamma_g * sum r_i rv_i,
                                        total_w = sum r_i w_i                                  (:345-373)
                                        returns normalize_batch([-total_w, total_c])           (:376-377)

The randomizers (u128::rand(rng), :371, the first one fixed to 1, :352) are an argument: they are data to the kernels.
All field elements are (.., 4) uint64 Montgomery Fr; points are Montgomery x||y rows.
"""
import numpy as np

from .binding import SCALARS_MONT, fq_limbs

from .params import FQ_MODULUS, FR_MODULUS


def _neg_limbs(limbs, mod):
    """-x for one field element given as little-endian uint64 limbs (Montgomery form negates like the integer)"""
    limbs = np.asarray(limbs, dtype=np.uint64).reshape(-1)
    v = sum(int(limbs[j]) << (64 * j) for j in range(limbs.size))
    v = (mod - v) % mod
    return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(limbs.size)], dtype=np.uint64)


def neg_point(curve, xy):
    n = fq_limbs(curve)
    xy = np.asarray(xy, dtype=np.uint64).reshape(-1).copy()
    xy[n:] = _neg_limbs(xy[n:], FQ_MODULUS[curve])
    return xy


def check_inner(eng, curve, g, gamma_g, comm, value, random_v=None):
    """kzg10/mod.rs:322-325: the G1 argument of the left-hand pairing, comm - g*value - gamma_g*random_v -> (xy, is_identity)"""
    r = FR_MODULUS[curve]
    one = np.zeros(4, dtype=np.uint64)
    one_int = (1 << 256) % r
    for j in range(4):
        one[j] = (one_int >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    bases = [np.asarray(comm, dtype=np.uint64).reshape(-1), np.asarray(g, dtype=np.uint64).reshape(-1)]
    scalars = [one, _neg_limbs(value, r)]
    if random_v is not None:
        bases.append(np.asarray(gamma_g, dtype=np.uint64).reshape(-1))
        scalars.append(_neg_limbs(random_v, r))
    return eng.msm_bases(curve, np.stack(bases), np.stack(scalars), flags=SCALARS_MONT)


def batch_check_combine(eng, curve, g, gamma_g, commitments, points, values, proofs_w, randomizers, random_vs=None):
    """kzg10/mod.rs:345-377.  commitments / proofs_w: (m, 2*limbs) points; points / values / randomizers: (m, 4) Fr;
    random_vs: (m, 4) Fr or None (no hiding).  Returns ((neg_total_w_xy, is_identity), (total_c_xy, is_identity))."""
    r = FR_MODULUS[curve]
    commitments = np.asarray(commitments, dtype=np.uint64)
    m = commitments.shape[0]
    proofs_w = np.asarray(proofs_w, dtype=np.uint64).reshape(m, -1)
    points, values, rnd = (np.asarray(a, dtype=np.uint64).reshape(m, 4) for a in (points, values, randomizers))
    rz = eng.fr_mul(curve, rnd, points)                                        # randomizer * z_i          (:360, :368)
    g_mult = eng.fr_inner_product(curve, rnd, values)                          # sum randomizer * v_i      (:363)
    bases = [commitments, proofs_w, np.asarray(g, dtype=np.uint64).reshape(1, -1)]
    scalars = [rnd, rz, _neg_limbs(g_mult, r).reshape(1, 4)]
    if random_vs is not None:
        gg_mult = eng.fr_inner_product(curve, rnd, np.asarray(random_vs, dtype=np.uint64).reshape(m, 4))   # (:364-366)
        bases.append(np.asarray(gamma_g, dtype=np.uint64).reshape(1, -1))
        scalars.append(_neg_limbs(gg_mult, r).reshape(1, 4))
    total_c = eng.msm_bases(curve, np.concatenate(bases), np.concatenate(scalars), flags=SCALARS_MONT)       # (:368, :373-374)
    total_w = eng.msm_bases(curve, proofs_w, rnd, flags=SCALARS_MONT)                                        # (:369)
    neg_w = (np.zeros_like(total_w[0]), True) if total_w[1] else (neg_point(curve, total_w[0]), False)
    return neg_w, total_c
