"""Host mirror of InnerProductArgPC::open's halving loop (ipa_pc/mod.rs:612-722) over the device-resident
round kernels.  Names follow the reference (`compute_random_oracle_challenge`, `round_challenge`, `l_vec`,
`r_vec`, `final_comm_key`, `c`).

Deviation (documented): the reference hashes ark-serialize's `serialize_uncompressed` bytes and maps the Blake2s
digest to a field element with `Field::from_random_bytes` -- both live in un-vendored crates.  Here the hashed bytes
are the ABI's packed little-endian limbs (canonical integers for Fr, Montgomery x||y for points) and the digest is
reduced as a little-endian integer with its top bits cleared, retrying with an incremented counter exactly like
:74-87.  The challenge values are data to the kernels; parity is asserted on every group/field output.
"""
import hashlib

import numpy as np

from .binding import fq_limbs

_MODULI = {
    0: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    1: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    2: 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001,
}


def _fr_mont(curve, v):
    r = _MODULI[curve]
    m = v * (1 << 256) % r
    return np.array([(m >> (64 * j)) & (2**64 - 1) for j in range(4)], dtype=np.uint64)


def compute_random_oracle_challenge(curve, data):
    """ipa_pc/mod.rs:74-87 with Blake2s-256 (the digest the reference's tests instantiate, ipa_pc/mod.rs:1051+)."""
    r = _MODULI[curve]
    i = 0
    while True:
        h = hashlib.blake2s(data + i.to_bytes(8, "little")).digest()
        v = int.from_bytes(h, "little") & ((1 << (r.bit_length() - 1)) - 1)
        if 0 < v < r:
            return v
        i += 1


def open_rounds(eng, curve, comm_key_xy, coeffs, point, h_prime_xy, round_challenge):
    """The `while n > 1` loop (:665-711).  comm_key_xy: n affine points; coeffs: <= n Montgomery Fr; point: Montgomery
    Fr; h_prime_xy: affine h' = h * round_challenge (:631); round_challenge: the initial challenge as an int.
    Returns dict(l_vec, r_vec, final_comm_key, c, challenges)."""
    r = _MODULI[curve]
    st = eng.ipa_begin(curve, comm_key_xy, coeffs, point)
    l_vec, r_vec, chals = [], [], []
    while eng.ipa_len(st) > 1:
        l, rr = eng.ipa_round_lr(curve, st, h_prime_xy)
        l_vec.append(l)
        r_vec.append(rr)
        data = int(round_challenge).to_bytes(32, "little") + l.tobytes() + rr.tobytes()   # :681-687
        round_challenge = compute_random_oracle_challenge(curve, data)
        chals.append(round_challenge)
        inv = pow(round_challenge, -1, r)                                                    # :689
        eng.ipa_round_fold(st, _fr_mont(curve, round_challenge), _fr_mont(curve, inv))      # :691-708
    final_key, c = eng.ipa_finish(curve, st)
    return dict(l_vec=l_vec, r_vec=r_vec, final_comm_key=final_key, c=c, challenges=chals)


def check_final_key(eng, curve, comm_key_xy, challenges):
    """InnerProductArgPC::check's linear-time step (ipa_pc/mod.rs:760-766): the key the verifier recomputes from the round
    challenges, cm_commit(vk.comm_key, check_poly.compute_coeffs()) -- must equal proof.final_comm_key."""
    srs = eng.srs_register(curve, comm_key_xy)
    ch = np.stack([_fr_mont(curve, int(c)) for c in challenges])
    out = eng.ipa_check_final_key(srs, ch)
    srs.release()
    return out
