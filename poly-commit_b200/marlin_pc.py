"""Host mirror of MarlinKZG10's prover calls (poly-commit/src/marlin/marlin_pc/mod.rs) over the C ABI, non-hiding path.

  CommitterKey.powers / shifted_powers     marlin_pc/data_structures.rs:46-84
  shift_polynomial                          marlin_pc/mod.rs:34-53
  commit   (per-polynomial loop)            marlin_pc/mod.rs:172-242   -> KZG10::commit (+ shifted commitment for a degree bound)
  open     (challenge-weighted combination) marlin_pc/mod.rs:245-336   -> p += (challenge_j, poly_j) :286, witness :292-297,
                                                                          KZG10::open :310, shifted opening :317-326

The opening challenges are squeezed from a Poseidon sponge in the reference (:282, :299); here they are an argument
(the sponge is out of scope) -- they are data to the kernels.
Polynomials are (n, 4) uint64 arrays of Montgomery Fr coefficients, low degree first.
"""
import numpy as np

from .binding import SCALARS_MONT


class CommitterKey:
    def __init__(self, eng, curve, powers_xy, shifted_powers_xy=None, enforced_degree_bounds=None, flags=0):
        self.eng, self.curve = eng, curve
        self.powers = eng.srs_register(curve, powers_xy, flags=flags)                      # powers_of_g[0 ..= supported_degree]
        self.enforced_degree_bounds = sorted(enforced_degree_bounds) if enforced_degree_bounds else None
        self.shifted = eng.srs_register(curve, shifted_powers_xy, flags=flags) if shifted_powers_xy is not None else None

    def supported_degree(self):
        return len(self.powers) - 1

    def shifted_offset(self, degree_bound):
        """shifted_powers(bound) = shifted_powers[(max_bound - bound)..]   (data_structures.rs:56-83)."""
        if self.shifted is None:
            raise ValueError("UnsupportedDegreeBound")
        if degree_bound is None:
            return 0
        if degree_bound not in self.enforced_degree_bounds:
            raise ValueError("UnsupportedDegreeBound")
        return self.enforced_degree_bounds[-1] - degree_bound


def _degree(coeffs):
    nz = np.nonzero(np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4).any(axis=1))[0]
    return int(nz[-1]) if nz.size else 0


def shift_polynomial(ck, p, degree_bound):
    """marlin_pc/mod.rs:34-53: prepend (largest enforced bound - degree_bound) zero coefficients."""
    p = np.asarray(p, dtype=np.uint64).reshape(-1, 4)
    if not p.any():
        return np.zeros((0, 4), dtype=np.uint64)
    pad = ck.enforced_degree_bounds[-1] - degree_bound
    return np.concatenate([np.zeros((pad, 4), dtype=np.uint64), p])


def commit(ck, polynomials):
    """polynomials: list of (coeffs, degree_bound or None).  Returns [(comm, shifted_comm or None)] as (xy, inf) pairs."""
    out = []
    for coeffs, bound in polynomials:
        if bound is not None and (bound < _degree(coeffs) or bound not in (ck.enforced_degree_bounds or [])):
            raise ValueError("IncorrectDegreeBound")                                  # check_degrees_and_bounds, kzg10/mod.rs:424-450
        comm = ck.eng.kzg_commit(ck.powers, coeffs)                                   # :217
        shifted = None
        if bound is not None:                                                         # :219-225
            c = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)
            shifted = ck.eng.msm(ck.shifted, c, base_offset=ck.shifted_offset(bound), flags=SCALARS_MONT)
        out.append((comm, shifted))
    return out


def open(ck, polynomials, point, challenges):
    """polynomials as in commit; challenges: iterator of Montgomery Fr (one per polynomial, one more per degree bound).
    Returns the proof point w as (xy, inf)."""
    eng, cid = ck.eng, ck.curve
    ch = iter(challenges)
    nmax = max(np.asarray(c).reshape(-1, 4).shape[0] for c, _ in polynomials)
    p = np.zeros((nmax, 4), dtype=np.uint64)
    shifted_w = None
    for coeffs, bound in polynomials:
        coeffs = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        cj = next(ch)
        p[: coeffs.shape[0]] = eng.fr_axpy(cid, p[: coeffs.shape[0]], cj, coeffs)    # p += (challenge_j, polynomial)  :286
        if bound is not None:
            witness, _ = eng.fr_div_linear(cid, coeffs, point)                        # compute_witness_polynomial  :292-297
            sw = shift_polynomial(ck, witness, bound)                                 # :300
            cj1 = next(ch)
            if shifted_w is None:
                shifted_w = np.zeros((ck.enforced_degree_bounds[-1] + 1, 4), dtype=np.uint64)
            shifted_w[: sw.shape[0]] = eng.fr_axpy(cid, shifted_w[: sw.shape[0]], cj1, sw)   # shifted_w += (challenge_j_1, shifted_witness) :302
    w_xy, w_inf, _ = eng.kzg_open(ck.powers, p, point)                                # :310
    if shifted_w is None:
        return w_xy, w_inf
    # open_with_witness_polynomial(&ck.shifted_powers(None), ..., &shifted_w, ...)  :317-326 ; w += shifted_proof.w
    a = eng.msm_partial(ck.powers, eng.fr_div_linear(cid, p, point)[0], flags=SCALARS_MONT)
    b = eng.msm_partial(ck.shifted, shifted_w[: _degree(shifted_w) + 1], flags=SCALARS_MONT)
    return eng.g1_sum_xyzz(cid, np.concatenate([a, b]))


def accumulate_commitments_and_values(eng, curve, commitments, values, challenges, shift_powers=None):
    """Marlin::accumulate_commitments_and_values (marlin/mod.rs:109-148), the verifier-side combination:
         combined_comm  = sum_i challenge_i * comm_i + challenge_i' * (shifted_comm_i - value_i * shift_power(bound_i))
         combined_value = sum_i challenge_i * value_i
    commitments: list of (comm_xy, shifted_comm_xy or None, degree_bound or None); values: (m, 4) Montgomery Fr;
    challenges: iterator of Montgomery Fr in the order the sponge yields them (:123, :129-130); shift_powers: {bound: point}
    (VerifierKey::get_shift_power).  One MSM over the commitments (pcgpu_msm_bases); returns ((xy, is_identity), value)."""
    from .kzg10 import _neg_limbs
    from .params import FR_MODULUS
    ch = iter(challenges)
    values = np.asarray(values, dtype=np.uint64).reshape(-1, 4)
    bases, scalars, ch_plain = [], [], []
    for (comm, shifted, bound), v in zip(commitments, values):
        if (bound is None) != (shifted is None):
            raise ValueError("degree bound and shifted commitment must come together")      # assert_eq!, :119
        c = np.asarray(next(ch), dtype=np.uint64).reshape(4)
        bases.append(np.asarray(comm, dtype=np.uint64).reshape(-1)); scalars.append(c); ch_plain.append(c)
        if bound is not None:
            if shift_powers is None or bound not in shift_powers:
                raise ValueError("UnsupportedDegreeBound")                                  # :136
            c1 = np.asarray(next(ch), dtype=np.uint64).reshape(4)
            c1v = eng.fr_mul(curve, c1.reshape(1, 4), v.reshape(1, 4))[0]
            bases += [np.asarray(shifted, dtype=np.uint64).reshape(-1), np.asarray(shift_powers[bound], dtype=np.uint64).reshape(-1)]
            scalars += [c1, _neg_limbs(c1v, FR_MODULUS[curve])]
    combined_value = eng.fr_inner_product(curve, np.stack(ch_plain), values[: len(ch_plain)])
    return eng.msm_bases(curve, np.stack(bases), np.stack(scalars), flags=SCALARS_MONT), combined_value
