"""Host mirror of the kzg10 containers' CanonicalSerialize / CanonicalDeserialize.

  Powers          kzg10/data_structures.rs:142-177   powers_of_g: Vec<G1Affine>, powers_of_gamma_g: Vec<G1Affine>
  Commitment      kzg10/data_structures.rs:315-328   one G1Affine (derived impl)
  Proof           kzg10/data_structures.rs:479-495   w: G1Affine, random_v: Option<Fr> (derived impl)

Container framing is ark-serialize's (un-vendored, restated): a Vec is a little-endian u64 length followed by the
elements; an Option is one byte (0 / 1) followed by the value; Fr is 32 little-endian canonical bytes.  The element work
(point encoding, decompression, on-curve and subgroup validation) runs on the GPU through pcgpu_g1_serialize /
pcgpu_g1_deserialize; this file only walks the framing.
"""
import struct

import numpy as np

from .binding import WireError  # noqa: F401  (re-exported: what deserialization raises)


def _vec_serialize(eng, curve, xy, inf, compressed):
    xy = np.asarray(xy, dtype=np.uint64)
    n = xy.reshape(-1, xy.shape[-1]).shape[0] if xy.size else 0
    body = eng.g1_serialize(curve, xy, inf, compressed).tobytes() if n else b""
    return struct.pack("<Q", n) + body


def _vec_deserialize(eng, curve, data, off, compressed, validate):
    if len(data) < off + 8:
        raise ValueError("truncated input (vector length)")
    (n,) = struct.unpack_from("<Q", data, off)
    off += 8
    sz = eng.g1_wire_size(curve, compressed)
    if len(data) < off + n * sz:
        raise ValueError("truncated input (vector body)")
    xy, inf = eng.g1_deserialize(curve, np.frombuffer(data, dtype=np.uint8, count=n * sz, offset=off), n, compressed, validate)
    return xy, inf, off + n * sz


def powers_serialize(eng, curve, powers_of_g, powers_of_gamma_g, compressed=True, inf_g=None, inf_gamma=None):
    """Powers::serialize_with_mode (data_structures.rs:142-156): the two vectors back to back."""
    return (_vec_serialize(eng, curve, powers_of_g, inf_g, compressed)
            + _vec_serialize(eng, curve, powers_of_gamma_g, inf_gamma, compressed))


def powers_deserialize(eng, curve, data, compressed=True, validate=True):
    """Powers::deserialize_with_mode (data_structures.rs:159-177) -> ((xy, inf) of powers_of_g, (xy, inf) of powers_of_gamma_g).
    Powers::check is a no-op (:137-141); element validation happens inside the vector reads, as in the reference."""
    data = bytes(data)
    g_xy, g_inf, off = _vec_deserialize(eng, curve, data, 0, compressed, validate)
    h_xy, h_inf, off = _vec_deserialize(eng, curve, data, off, compressed, validate)
    return (g_xy, g_inf), (h_xy, h_inf)


def commitment_serialize(eng, curve, comm_xy, comm_inf=False, compressed=True):
    """kzg10::Commitment(G1Affine), data_structures.rs:315-328"""
    inf = np.array([1 if comm_inf else 0], dtype=np.uint8)
    return eng.g1_serialize(curve, np.asarray(comm_xy, dtype=np.uint64).reshape(1, -1), inf, compressed).tobytes()


def commitment_deserialize(eng, curve, data, compressed=True, validate=True):
    xy, inf = eng.g1_deserialize(curve, bytes(data), 1, compressed, validate)
    return xy[0], bool(inf[0])


def proof_serialize(eng, curve, w_xy, w_inf=False, random_v=None, compressed=True):
    """kzg10::Proof { w, random_v: Option<Fr> }, data_structures.rs:479-495.  random_v: (4,) uint64 Montgomery or None."""
    out = commitment_serialize(eng, curve, w_xy, w_inf, compressed)
    if random_v is None:
        return out + b"\x00"
    canon = eng.fr_from_mont(curve, np.asarray(random_v, dtype=np.uint64).reshape(1, 4))
    return out + b"\x01" + canon.astype("<u8").tobytes()
