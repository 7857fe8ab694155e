"""Checks against tests/golden/external_kats.json -- vectors published OUTSIDE this repository (EIP-196 / go-ethereum,
EIP-2537, the ZCash / zkcrypto BLS12-381 test vectors and constants, pasta_curves, gnark-crypto, RFC 7693, FIPS 180-4).
Shared by the CPU tests (both oracles, the host-emulated kernels) and the GPU tests (the CUDA library through the C ABI).

What each one pins (SURVEY.md section 8c's "[ark-dep, from memory]" list):
  * curve equation, generator and group law of the three curves            -> 2G, 3G, the EIP-196 ecMul vector, Pallas 2G
  * Montgomery radix R = 2^(64 limbs) and the little-endian limb layout     -> zkcrypto's R constants and generator limbs
  * the ZCash encoding ark-bls12-381 uses for G1                            -> 1G / 2G / 3G / identity byte strings
  * Fr multiplicative generator, two-adicity, 2^s-th root of unity, and the
    domain-generator rule w_k = root^(2^(s-k)) of Radix2EvaluationDomain    -> published ROOT_OF_UNITY constants + NTT of X
"""
import hashlib
import json
import os

import numpy as np

from oracle import orc, pyref

_K = None


def kats():
    global _K
    if _K is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "external_kats.json")) as f:
            _K = json.load(f)
    return _K


def _i(s):
    return int(s, 0) if isinstance(s, str) else int(s)


def _pt(C, pair):
    return (_i(pair[0]) % C.p, _i(pair[1]) % C.p)


def published_points(cname):
    """[(k, affine point)] : k * G as published, per curve"""
    C, k = pyref.Curve(cname), kats()[cname]
    if cname == "bn254":
        return C, [(1, _pt(C, k["generator"])), (2, _pt(C, k["two_g"])), (3, _pt(C, k["three_g"]))]
    if cname == "bls12_381":
        g = pyref.g1_deserialize(C, bytes.fromhex(k["generator_compressed"]), 1, True)[0]
        return C, [(1, g), (2, _pt(C, k["two_g"]))]
    r = k["two_g_rational"]
    two = (r["x"][0] * pow(r["x"][1], -1, C.p) % C.p, r["y"][0] * pow(r["y"][1], -1, C.p) % C.p)
    return C, [(1, _pt(C, k["generator"])), (2, two)]


def published_root(cname):
    k = kats()[cname]
    return _i(k["fr_generator"]), _i(k["fr_two_adicity"]), _i(k["fr_two_adic_root_of_unity"])


def limbs(vals):
    return np.array([_i(v) for v in vals], dtype=np.uint64)


# ---------------------------------------------------------------------------------------------------------------------
def check_pyref(cname):
    C, pts = published_points(cname)
    assert pts[0][1] == C.g
    for k, P in pts:
        assert C.on_curve(P) and C.mul(k, C.g) == P and C.mul(C.r, P) is None
    gen, s, root = published_root(cname)
    assert C.two_adicity == s and C.root == root
    assert pow(gen, (C.r - 1) >> s, C.r) == root and pow(root, 1 << (s - 1), C.r) == C.r - 1
    # the multiplicative generator really generates: g^((r-1)/q) != 1 for the small prime factors we can test cheaply
    assert pow(gen, (C.r - 1) // 2, C.r) == C.r - 1
    kk = kats()[cname]
    if cname == "bn254":
        e = kk["ecmul"]
        assert C.mul(_i(e["scalar"]), _pt(C, e["point"])) == _pt(C, e["result"])
    if cname == "bls12_381":
        assert kats()["bls12_381"]["fr_two_adic_root_of_unity_decimal"] == str(root)
        for key, k in (("generator_compressed", 1), ("two_g_compressed", 2), ("three_g_compressed", 3)):
            P = C.mul(k, C.g)
            assert pyref.g1_serialize(C, [P], True).hex() == kk[key]
            assert pyref.g1_deserialize(C, bytes.fromhex(kk[key]), 1, True) == [P]
        assert pyref.g1_serialize(C, [None], True).hex() == kk["identity_compressed"]
        assert pyref.g1_serialize(C, [None], False).hex() == kk["identity_uncompressed"]
        assert pyref.g1_deserialize(C, bytes.fromhex(kk["identity_compressed"]), 1, True) == [None]
        assert pyref.g1_deserialize(C, bytes.fromhex(kk["identity_uncompressed"]), 1, False) == [None]
        assert [int(v) for v in limbs(kk["fq_montgomery_one_limbs"])] == [(C.Rq >> (64 * j)) & (2**64 - 1) for j in range(6)]
        assert [int(v) for v in limbs(kk["fr_montgomery_one_limbs"])] == [(C.Rr >> (64 * j)) & (2**64 - 1) for j in range(4)]
        xy, _ = C.points_to_limbs([C.g])
        gm = kk["generator_montgomery_limbs"]
        assert (xy[0] == np.concatenate([limbs(gm["x"]), limbs(gm["y"])])).all()
    if cname == "pallas":
        assert pow(_i(kk["fq_generator"]), (C.p - 1) >> 32, C.p) == _i(kk["fq_two_adic_root_of_unity"])
    h = kats()["hashes"]
    assert hashlib.blake2s(b"abc").hexdigest() == h["blake2s_abc"] and hashlib.sha256(b"abc").hexdigest() == h["sha256_abc"]


def check_oracle(cname):
    """the C oracle (oracle/pc_oracle.c): generator limbs, multiples, the ecMul vector through its MSM, roots of unity"""
    C, pts = published_points(cname)
    G = orc.g1_generator(C.id)
    assert C.points_from_limbs(G)[0] == pts[0][1]
    for k, P in pts:
        xy, inf = orc.g1_mul(C.id, G, C.fr_to_limbs([k], False))
        assert inf == 0 and C.points_from_limbs(xy)[0] == P
    gen, s, root = published_root(cname)
    assert C.fr_from_limbs(orc.fr_domain_generator(C.id, s), True)[0] == root
    for logn in (1, 7, 20):
        assert C.fr_from_limbs(orc.fr_domain_generator(C.id, logn), True)[0] == pow(root, 1 << (s - logn), C.r)
    one_m = orc.field_unop("orc_fr_to_mont", C.id, C.fr_to_limbs([1], False))
    if cname == "bls12_381":
        kk = kats()[cname]
        assert (one_m[0] == limbs(kk["fr_montgomery_one_limbs"])).all()
        gm = kk["generator_montgomery_limbs"]
        assert (G == np.concatenate([limbs(gm["x"]), limbs(gm["y"])])).all()
    if cname == "bn254":
        e = kats()[cname]["ecmul"]
        base, _ = C.points_to_limbs([_pt(C, e["point"])])
        for naive in (True, False):
            xy, inf = orc.msm(C.id, base, C.fr_to_limbs([_i(e["scalar"])], False), naive=naive)
            assert inf == 0 and C.points_from_limbs(xy)[0] == _pt(C, e["result"])


def check_engine(eng, cname, max_logn=12):
    """the library behind the C ABI (CUDA on the GPU box; the host-emulated kernel bodies in the CPU suite)"""
    C, pts = published_points(cname)
    gxy, _ = C.points_to_limbs([pts[0][1]])
    ks = C.fr_to_limbs([k for k, _ in pts], False)
    got = eng.fixed_base_mul(C.id, gxy[0], ks)
    for row, (_, P) in zip(got, pts):
        assert C.points_from_limbs(row)[0] == P
    # the same multiples as one-term and as multi-term MSMs on unregistered bases: k*G = sum of k_i * G with sum k_i = k
    for k, P in pts:
        xy, inf = eng.msm_bases(C.id, gxy, C.fr_to_limbs([k], False))
        assert not inf and C.points_from_limbs(xy)[0] == P
    kk = kats()[cname]
    if cname == "bn254":
        e = kk["ecmul"]
        base, _ = C.points_to_limbs([_pt(C, e["point"])])
        xy, inf = eng.msm_bases(C.id, base, C.fr_to_limbs([_i(e["scalar"])], False))
        assert not inf and C.points_from_limbs(xy)[0] == _pt(C, e["result"])
        srs = eng.srs_register(C.id, np.concatenate([base, gxy]))
        # scalar * P + 3 * G against published 3G and the published product, combined by the oracle-free affine law in pyref
        xy, inf = eng.msm(srs, C.fr_to_limbs([_i(e["scalar"]), 3], False))
        assert C.points_from_limbs(xy)[0] == C.add(_pt(C, e["result"]), _pt(C, kk["three_g"]))
    if cname == "bls12_381":
        enc = [kk["generator_compressed"], kk["two_g_compressed"], kk["three_g_compressed"], kk["identity_compressed"]]
        three = eng.fixed_base_mul(C.id, gxy[0], C.fr_to_limbs([1, 2, 3], False))
        xy = np.concatenate([three, np.zeros((1, 12), dtype=np.uint64)])
        inf = np.array([0, 0, 0, 1], dtype=np.uint8)
        assert eng.g1_serialize(C.id, xy, inf, True).tobytes().hex() == "".join(enc)
        bxy, binf = eng.g1_deserialize(C.id, bytes.fromhex("".join(enc)), 4, True)
        assert (bxy == xy).all() and (binf == inf).all()
        bxy, binf = eng.g1_deserialize(C.id, bytes.fromhex(kk["identity_uncompressed"]), 1, False)
        assert not bxy.any() and binf[0] == 1
        # Montgomery radix: from_mont(published R) == 1
        assert (eng.fr_from_mont(C.id, limbs(kk["fr_montgomery_one_limbs"]).reshape(1, 4))[0] == np.array([1, 0, 0, 0], dtype=np.uint64)).all()
    # NTT domain convention: the transform of p(X) = X is out[j] = w^j with w = published_root^(2^(s - logn))
    gen, s, root = published_root(cname)
    for logn in sorted({1, 2, 9, max_logn}):
        out = eng.ntt(C.id, C.fr_to_limbs([0, 1], True), logn)
        w = pow(root, 1 << (s - logn), C.r)
        idx = sorted({0, 1, min(2, (1 << logn) - 1), (1 << logn) - 1})
        assert C.fr_from_limbs(out[idx], True) == [pow(w, j, C.r) for j in idx]
        back = eng.ntt(C.id, out, logn, inverse=True)
        assert C.fr_from_limbs(back[:2], True) == [0, 1] and not back[2:].any()
