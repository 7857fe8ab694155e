"""kzg10::UniversalParams / VerifierKey / Proof containers (kzg10/data_structures.rs:22-112, :196-262, :479-495) over the
device G1 codec + the host G2 routine, against byte strings assembled independently from oracle/pyref.py (G1 and G2
restatements) and the ark-serialize framing rules; published G2 generators (EIP-197, IETF pairing-friendly-curves draft) pin
the G2 encodings.  CPU: host-emulated kernels; GPU: the CUDA library, including a 2^20-power SRS with one corrupted element."""
import struct

import numpy as np
import pytest

from oracle import pyref
from tests import external_cases, util


@pytest.fixture(scope="module")
def emul(pc, hostcheck_path):
    e = pc.Engine(0, lib_path=hostcheck_path)
    yield e
    e.close()


def _g2_gen(cname):
    g = external_cases.kats()[cname]["g2_generator"]
    i = external_cases._i
    return ((i(g["x_c0"]), i(g["x_c1"])), (i(g["y_c0"]), i(g["y_c1"])))


@pytest.mark.parametrize("cname", ["bls12_381", "bn254"])
def test_g2_host_vs_pyref_and_published_generator(pc, cname):
    from poly_commit_b200 import g2_host
    C, G2 = pyref.Curve(cname), pyref.G2(cname)
    P = _g2_gen(cname)
    assert G2.on_curve(P) and G2.mul(C.r, P) is None and g2_host.g2_on_curve(C.id, P)
    if cname == "bls12_381":
        assert g2_host.g2_serialize(C.id, P, True).hex() == external_cases.kats()[cname]["g2_generator_compressed"]
    pts = [P, G2.mul(2, P), G2.mul(0xdeadbeefcafe, P), None]
    for Q in pts:
        assert g2_host.g2_mul(C.id, 7, Q) == G2.mul(7, Q)
        for compressed in (True, False):
            b = g2_host.g2_serialize(C.id, Q, compressed)
            assert b == G2.serialize(Q, compressed) and len(b) == g2_host.g2_wire_size(C.id, compressed) == G2.size(compressed)
            assert g2_host.g2_deserialize(C.id, b, compressed) == Q == G2.deserialize(b, compressed)
            xy, inf = g2_host.g2_to_limbs(C.id, Q)
            assert g2_host.g2_from_limbs(C.id, xy, inf) == Q
    # rejections: same reason from both implementations
    bad = []
    b = bytearray(G2.serialize(P, True))
    if cname == "bls12_381":
        b[0] ^= 0x40                                                       # infinity flag with a non-zero payload
    else:
        b[-1] |= 0xC0                                                      # both SWFlags (the generic form ignores x under the infinity flag)
    bad.append((bytes(b), True))
    x = (5, 1)
    while G2.sqrt2(G2.rhs(x)) is not None:
        x = (x[0] + 1, 1)
    bad.append((G2.serialize((x, (0, 0)), True), True))                                                                 # x without a point
    Q = None
    x = (1, 2)
    while Q is None:                                                                                                    # on the twist, outside the r-torsion
        y = G2.sqrt2(G2.rhs(x))
        if y is not None and G2.mul(C.r, (x, y)) is not None:
            Q = (x, y)
        x = (x[0] + 1, 2)
    bad.append((G2.serialize(Q, True), True)); bad.append((G2.serialize(Q, False), False))
    bad.append((G2.serialize((P[0], ((P[1][0] + 1) % C.p, P[1][1])), False), False))                                    # off the curve
    for data, compressed in bad:
        with pytest.raises(pyref.WireError) as e0:
            G2.deserialize(data, compressed)
        with pytest.raises(g2_host.G2WireError) as e1:
            g2_host.g2_deserialize(C.id, data, compressed)
        assert e0.value.reason == e1.value.reason, data.hex()
    assert g2_host.g2_deserialize(C.id, G2.serialize(Q, False), False, validate=False) == Q


def _build_params(cname, n, n_gamma, seed):
    """synthetic UniversalParams: powers beta^i G, gamma powers, h, beta_h = 7 h, two neg_powers_of_h entries"""
    C, G2 = pyref.Curve(cname), pyref.G2(cname)
    powers = util.synthetic_srs(cname, n, seed=seed)
    gamma = util.random_points(cname, n_gamma, seed=seed + 1)
    keys = np.array(sorted({0, 1, 2, 5, 9, 1 << 20, (1 << 40) + 3})[:n_gamma], dtype=np.uint64)
    h = _g2_gen(cname)
    beta_h = G2.mul(7, h)
    neg = {0: h, 37: G2.mul(11, h)}
    return C, G2, powers, keys, gamma, h, beta_h, neg


def _expected_bytes(C, G2, powers, keys, gamma, h, beta_h, neg, compressed):
    out = struct.pack("<Q", powers.shape[0]) + pyref.g1_serialize(C, C.points_from_limbs(powers), compressed)
    out += struct.pack("<Q", len(keys))
    for k, P in zip(keys, C.points_from_limbs(gamma)):
        out += struct.pack("<Q", int(k)) + pyref.g1_serialize(C, [P], compressed)
    out += G2.serialize(h, compressed) + G2.serialize(beta_h, compressed) + struct.pack("<Q", len(neg))
    for k in sorted(neg):
        out += struct.pack("<Q", k) + G2.serialize(neg[k], compressed)
    return out


def _check_universal_params(eng, pc, cname):
    from poly_commit_b200 import wire
    C, G2, powers, keys, gamma, h, beta_h, neg = _build_params(cname, 40, 5, seed=31)
    for compressed in (True, False):
        blob = wire.universal_params_serialize(eng, C.id, powers, keys, gamma, h, beta_h, neg, compressed)
        assert blob == _expected_bytes(C, G2, powers, keys, gamma, h, beta_h, neg, compressed)
        up = wire.universal_params_deserialize(eng, C.id, blob, compressed)
        assert (up["powers_of_g"][0] == powers).all() and not up["powers_of_g"][1].any()
        assert (up["powers_of_gamma_g"][0] == keys).all() and (up["powers_of_gamma_g"][1] == gamma).all()
        assert up["h"] == h and up["beta_h"] == beta_h and up["neg_powers_of_h"] == neg and up["consumed"] == len(blob)
        sz = pyref.wire_size(C, compressed)
        # an element that only validation rejects (off the curve when uncompressed; outside the subgroup on BLS12-381)
        bad = bytearray(blob)
        if not compressed:
            P = C.points_from_limbs(powers[9:10])[0]
            bad[8 + 9 * sz:8 + 10 * sz] = pyref.g1_serialize(C, [(P[0], (P[1] + 1) % C.p)], False)
            want = pyref.WIRE_NOT_ON_CURVE
        elif cname == "bls12_381":
            bad[8 + 9 * sz:8 + 10 * sz] = pyref.g1_serialize(C, [pyref.curve_point_from_x_search(C, 1000)], True)
            want = pyref.WIRE_NOT_IN_SUBGROUP
        else:
            want = None
        if want is not None:
            with pytest.raises(wire.KeyError_) as ei:
                wire.universal_params_deserialize(eng, C.id, bytes(bad), compressed)
            assert (ei.value.section, ei.value.index, ei.value.reason) == ("powers_of_g", 9, want)
            wire.universal_params_deserialize(eng, C.id, bytes(bad), compressed, validate=False)       # Validate::No accepts it
            # ... but an ENCODING error further down the stream takes precedence (fields are decoded with Validate::No first)
            goff = 8 + 40 * sz + 8 + 2 * (8 + sz) + 8                                                  # gamma entry 2's point
            if cname == "bls12_381":
                bad[goff] ^= 0x80                                                                      # compression bit flipped
            else:
                bad[goff + sz - 1] |= 0xC0                                                             # both SW flags
            with pytest.raises(wire.KeyError_) as ei:
                wire.universal_params_deserialize(eng, C.id, bytes(bad), compressed)
            assert (ei.value.section, ei.value.index, ei.value.reason) == ("powers_of_gamma_g", 2, pyref.WIRE_BAD_FLAGS)
        # beta_h outside the subgroup: reported by the whole-struct check
        x = (3, 1)
        Q = None
        while Q is None:
            y = G2.sqrt2(G2.rhs(x))
            if y is not None and G2.mul(C.r, (x, y)) is not None:
                Q = (x, y)
            x = (x[0] + 1, 1)
        blob2 = wire.universal_params_serialize(eng, C.id, powers, keys, gamma, h, Q, neg, compressed)
        with pytest.raises(wire.KeyError_) as ei:
            wire.universal_params_deserialize(eng, C.id, blob2, compressed)
        assert (ei.value.section, ei.value.reason) == ("beta_h", pyref.WIRE_NOT_IN_SUBGROUP)
        assert wire.universal_params_deserialize(eng, C.id, blob2, compressed, validate=False)["beta_h"] == Q
        with pytest.raises(ValueError):
            wire.universal_params_deserialize(eng, C.id, blob[:-3], compressed)
        # VerifierKey
        vk = wire.verifier_key_serialize(eng, C.id, powers[0], gamma[0], h, beta_h, compressed)
        assert vk == (pyref.g1_serialize(C, C.points_from_limbs(np.stack([powers[0], gamma[0]])), compressed)
                      + G2.serialize(h, compressed) + G2.serialize(beta_h, compressed))
        back = wire.verifier_key_deserialize(eng, C.id, vk, compressed)
        assert (back["g"][0] == powers[0]).all() and (back["gamma_g"][0] == gamma[0]).all() and back["h"] == h and back["beta_h"] == beta_h
    # Proof with random_v round trip
    rv = util.rand_fr(cname, 1, seed=77, mont=True)[0]
    pb = wire.proof_serialize(eng, C.id, powers[3], False, rv)
    w, winf, rv2 = wire.proof_deserialize(eng, C.id, pb)
    assert (w == powers[3]).all() and not winf and (rv2 == rv).all()
    assert wire.proof_deserialize(eng, C.id, wire.proof_serialize(eng, C.id, powers[3], False, None))[2] is None


@pytest.mark.parametrize("cname", ["bls12_381", "bn254"])
def test_universal_params_and_verifier_key(emul, pc, cname):
    _check_universal_params(emul, pc, cname)


@pytest.mark.gpu
@pytest.mark.parametrize("cname", ["bls12_381", "bn254"])
def test_gpu_universal_params_and_verifier_key(gpu_engine, pc, cname):
    _check_universal_params(gpu_engine, pc, cname)


@pytest.mark.gpu
def test_gpu_universal_params_2p20_locates_corrupted_power(gpu_engine, pc):
    """a UniversalParams of 2^20 + 1 compressed BLS12-381 powers: round trip through the device codec, then one power is
    replaced by an on-curve point outside the prime-order subgroup and the reader names its index"""
    from poly_commit_b200 import params, wire
    eng, cid = gpu_engine, pc.BLS12_381
    C, G2 = pyref.Curve("bls12_381"), pyref.G2("bls12_381")
    n = (1 << 20) + 1
    ks = params.random_fr(cid, n, 4242)
    powers = eng.fixed_base_mul(cid, params.g1_generator(cid), ks)
    h = _g2_gen("bls12_381")
    keys = np.array([0, 1, 2], dtype=np.uint64)
    blob = wire.universal_params_serialize(eng, cid, powers, keys, powers[5:8], h, G2.mul(9, h), {0: h}, True)
    assert len(blob) == 8 + n * 48 + 8 + 3 * 56 + 2 * 96 + 8 + 8 + 96
    # spot-check the device encoder against pyref on a few elements
    for i in (0, 1, 777777, n - 1):
        assert blob[8 + 48 * i:8 + 48 * (i + 1)] == pyref.g1_serialize(C, C.points_from_limbs(powers[i:i + 1]), True)
    up = wire.universal_params_deserialize(eng, cid, blob, True)
    assert (up["powers_of_g"][0] == powers).all() and (up["powers_of_gamma_g"][1] == powers[5:8]).all()
    k = 654321
    bad = bytearray(blob)
    bad[8 + 48 * k:8 + 48 * (k + 1)] = pyref.g1_serialize(C, [pyref.curve_point_from_x_search(C, 5000)], True)
    with pytest.raises(wire.KeyError_) as ei:
        wire.universal_params_deserialize(eng, cid, bytes(bad), True)
    assert (ei.value.section, ei.value.index, ei.value.reason) == ("powers_of_g", k, pyref.WIRE_NOT_IN_SUBGROUP)
