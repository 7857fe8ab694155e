"""Host mirror of SonicKZG10's prover calls (poly-commit/src/sonic_pc/mod.rs) over the C ABI, non-hiding path.

  CommitterKey.powers / shifted_powers   sonic_pc/data_structures.rs:70-114 (shifted_powers(bound) = shifted_powers_of_g[(max_bound - bound)..])
  trim's shifted_powers_of_g             sonic_pc/mod.rs:186-196   (powers_of_g[max_degree - highest_bound ..])
  commit                                 sonic_pc/mod.rs:274-337   ONE commitment per polynomial: over the shifted powers when the
                                                                   polynomial carries a degree bound (:319-325), else over powers()
  open                                   sonic_pc/mod.rs:340-382   combined_polynomial += (curr_challenge, p_i) (:373), then ONE
                                                                   KZG10::open over powers() (:379)

The opening challenges come from a sponge in the reference (:362, :375); here they are an argument (data to the kernels; the
sponge is out of scope).  Polynomials are (n, 4) uint64 arrays of Montgomery Fr coefficients, low degree
first.  Differences from MarlinKZG10 (marlin_pc.py): no second "shifted" commitment and no shifted witness -- a bounded
polynomial is committed ONLY against the shifted key.
"""
import numpy as np

from .binding import SCALARS_MONT
from .marlin_pc import CommitterKey, _degree  # same key layout: powers + shifted powers + enforced bounds  # noqa: F401


def commit(ck, polynomials):
    """polynomials: list of (coeffs, degree_bound or None) -> [(comm_xy, is_identity)]  (sonic_pc/mod.rs:274-337)"""
    out = []
    for coeffs, bound in polynomials:
        coeffs = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        if bound is not None and (bound < _degree(coeffs) or bound not in (ck.enforced_degree_bounds or [])):
            raise ValueError("IncorrectDegreeBound")                                  # check_degrees_and_bounds, kzg10/mod.rs:424-450
        if bound is None:
            out.append(ck.eng.kzg_commit(ck.powers, coeffs))                           # ck.powers()
        else:                                                                          # ck.shifted_powers(bound): an offset view
            if coeffs.shape[0] > len(ck.shifted) - ck.shifted_offset(bound):
                raise ValueError("TooManyCoefficients")
            out.append(ck.eng.msm(ck.shifted, coeffs, base_offset=ck.shifted_offset(bound), flags=SCALARS_MONT))
    return out


def open(ck, polynomials, point, challenges):
    """sonic_pc/mod.rs:340-382: one challenge per polynomial (bounded or not), one KZG10 opening of the combination.
    Returns the proof point w as (xy, is_identity)."""
    eng, cid = ck.eng, ck.curve
    ch = iter(challenges)
    nmax = max(np.asarray(c).reshape(-1, 4).shape[0] for c, _ in polynomials)
    p = np.zeros((nmax, 4), dtype=np.uint64)
    for coeffs, bound in polynomials:
        coeffs = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        if bound is not None and (bound < _degree(coeffs) or bound not in (ck.enforced_degree_bounds or [])):
            raise ValueError("IncorrectDegreeBound")
        p[: coeffs.shape[0]] = eng.fr_axpy(cid, p[: coeffs.shape[0]], next(ch), coeffs)     # combined_polynomial += (challenge, p)  :373
    w_xy, w_inf, _ = eng.kzg_open(ck.powers, p, point)                                      # :379
    return w_xy, w_inf
