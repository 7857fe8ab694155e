"""Host mirror of HyraxPC's prover hot path (poly-commit/src/hyrax/mod.rs) over the C ABI.

  flat_to_matrix_column_major   hyrax/utils.rs:13-21
  pedersen_commit               hyrax/mod.rs:86-93
  commit (row loop)             hyrax/mod.rs:230-242   -> Engine.msm_batch over comb tables of com_key || h
  open: lt = mat.row_mul(l)     hyrax/mod.rs:347 -> utils.rs:127-146 -> Engine.fr_row_mul
  check: t_prime                hyrax/mod.rs:498-504   msm_bigint(&row_coms, &l_bigint) -> Engine.msm_bases (fresh bases)

The reference draws the row randomness r_i from `thread_rng()` when built with the `parallel` feature (:237-238), so its
commitments are not reproducible; here the randomness is an argument.
"""
import numpy as np

from .binding import SCALARS_MONT, SRS_COMB


def flat_to_matrix_column_major(flat, n, m):
    """(n*m, 4) flat vector -> n rows of length m, row[r][c] = flat[c*n + r]  (hyrax/utils.rs:13-21)."""
    flat = np.asarray(flat, dtype=np.uint64).reshape(m, n, 4)
    return np.ascontiguousarray(flat.transpose(1, 0, 2))


class CommitterKey:
    """com_key (dim generators) and the hiding generator h (hyrax/data_structures.rs), resident on the device with comb
    tables; h is stored as base `dim` so `pedersen_commit(row) + h * r` is one pass."""

    def __init__(self, eng, curve, com_key_xy, h_xy):
        self.eng, self.curve = eng, curve
        com_key_xy = np.asarray(com_key_xy, dtype=np.uint64)
        self.dim = com_key_xy.shape[0]
        self.srs = eng.srs_register(curve, np.concatenate([com_key_xy, np.asarray(h_xy, dtype=np.uint64).reshape(1, -1)]),
                                    flags=SRS_COMB)


def pedersen_commit(ck, scalars):
    """hyrax/mod.rs:86-93: MSM of `scalars` (Montgomery Fr) over the first len(scalars) generators."""
    scalars = np.asarray(scalars, dtype=np.uint64).reshape(-1, 4)
    return ck.eng.msm(ck.srs, scalars, flags=SCALARS_MONT)


def commit(ck, evaluations, randomness):
    """hyrax/mod.rs:230-242: evaluations (dim*dim Montgomery Fr, the multilinear polynomial's evaluation vector),
    randomness (dim Montgomery Fr, one r per row) -> (row_coms (dim affine points), mat (dim x dim rows))."""
    dim = ck.dim
    mat = flat_to_matrix_column_major(evaluations, dim, dim)
    rows = np.concatenate([mat, np.asarray(randomness, dtype=np.uint64).reshape(dim, 1, 4)], axis=1)
    row_coms, inf = ck.eng.msm_batch(ck.srs, rows, dim + 1, dim, flags=SCALARS_MONT)
    return row_coms, inf, mat


def open_row_mul(ck, mat, l):
    """lt = mat.row_mul(l)  (hyrax/mod.rs:347): l (dim) times the dim x dim matrix."""
    dim = ck.dim
    return ck.eng.fr_row_mul(ck.curve, l, np.ascontiguousarray(mat).reshape(-1, 4), dim, dim)


def check_t_prime(eng, curve, row_coms_xy, l, row_coms_inf=None):
    """hyrax/mod.rs:498-504: the verifier's multi-exponentiation of the row commitments by the tensor l (Montgomery Fr)."""
    return eng.msm_bases(curve, row_coms_xy, np.asarray(l, dtype=np.uint64).reshape(-1, 4), inf=row_coms_inf, flags=SCALARS_MONT)
