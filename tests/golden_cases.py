"""Golden-vector checks shared by the CPU (oracle, host emulation) and GPU test modules."""
import json
import os

import numpy as np

from oracle import orc, pyref

_G = None


def golden():
    global _G
    if _G is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.json")) as f:
            _G = json.load(f)
    return _G


def _pts(C, lst):
    return C.points_to_limbs([None if p is None else (int(p[0], 16), int(p[1], 16)) for p in lst])


def _fr(C, lst, mont):
    return C.fr_to_limbs([int(x, 16) for x in lst], mont)


def check_oracle(cname):
    """The C oracle against the Python-generated fixtures."""
    C = pyref.Curve(cname)
    g = golden()["curves"][cname]
    G = orc.g1_generator(C.id)
    assert C.points_from_limbs(G)[0] == (int(g["generator"][0], 16), int(g["generator"][1], 16))
    for m in g["multiples"]:
        k = C.fr_to_limbs([int(m["k"], 16)], False)
        xy, inf = orc.g1_mul(C.id, G, k)
        if m["point"] is None:
            assert inf == 1
        else:
            assert C.points_from_limbs(xy)[0] == (int(m["point"][0], 16), int(m["point"][1], 16))
    bases, _ = _pts(C, g["msm"]["bases"])
    sc = _fr(C, g["msm"]["scalars"], False)
    exp, _ = _pts(C, [g["msm"]["result"]])
    for naive in (True, False):
        xy, inf = orc.msm(C.id, bases, sc, naive=naive)
        assert (xy == exp[0]).all()
    fr = g["fr"]
    p, z = _fr(C, fr["coeffs"], True), _fr(C, [fr["z"]], True)[0]
    q, rem = orc.fr_div_linear(C.id, p, z)
    assert (q == _fr(C, fr["quotient"], True)).all() and (rem == _fr(C, [fr["remainder"]], True)[0]).all()
    assert (orc.fr_axpy(C.id, p, _fr(C, [fr["axpy_c"]], True)[0], _fr(C, fr["axpy_x"], True)) == _fr(C, fr["axpy_result"], True)).all()
    assert (orc.fr_inner_product(C.id, p, _fr(C, fr["axpy_x"], True)) == _fr(C, [fr["inner_product"]], True)[0]).all()
    for t in g["ntt"]:
        assert (orc.fr_domain_generator(C.id, t["logn"]) == _fr(C, [t["omega"]], True)[0]).all()
        exp = _fr(C, t["evals"], True)
        assert (orc.fr_ntt(C.id, _fr(C, t["coeffs"], True), t["logn"]) == exp).all()
        assert (orc.fr_ntt(C.id, _fr(C, t["coeffs"], True), t["logn"], naive=True) == exp).all()
    k = g["kzg"]
    powers, _ = _pts(C, k["powers"])
    poly, zz = _fr(C, k["poly"], True), _fr(C, [k["z"]], True)[0]
    rc, cxy, _ = orc.kzg_commit(C.id, powers, poly)
    assert rc == 0 and (cxy == _pts(C, [k["commitment"]])[0][0]).all()
    rc, wxy, _, _ = orc.kzg_open(C.id, powers, poly, zz)
    assert rc == 0 and (wxy == _pts(C, [k["witness"]])[0][0]).all()


def check_engine(eng, cname):
    """The engine (C ABI) against the same fixtures."""
    C = pyref.Curve(cname)
    g = golden()["curves"][cname]
    bases, _ = _pts(C, g["msm"]["bases"])
    sc = _fr(C, g["msm"]["scalars"], False)
    exp, _ = _pts(C, [g["msm"]["result"]])
    srs = eng.srs_register(C.id, bases)
    xy, inf = eng.msm(srs, sc)
    assert inf == 0 and (xy == exp[0]).all()
    # generator multiples through the fixed-base kernel
    ks = C.fr_to_limbs([int(m["k"], 16) for m in g["multiples"]], False)
    got = eng.fixed_base_mul(C.id, orc.g1_generator(C.id), ks)
    for row, m in zip(got, g["multiples"]):
        if m["point"] is None:
            assert not row.any()
        else:
            assert C.points_from_limbs(row)[0] == (int(m["point"][0], 16), int(m["point"][1], 16))
    fr = g["fr"]
    p, z = _fr(C, fr["coeffs"], True), _fr(C, [fr["z"]], True)[0]
    q, rem = eng.fr_div_linear(C.id, p, z)
    assert (q == _fr(C, fr["quotient"], True)).all() and (rem == _fr(C, [fr["remainder"]], True)[0]).all()
    assert (eng.fr_axpy(C.id, p, _fr(C, [fr["axpy_c"]], True)[0], _fr(C, fr["axpy_x"], True)) == _fr(C, fr["axpy_result"], True)).all()
    assert (eng.fr_inner_product(C.id, p, _fr(C, fr["axpy_x"], True)) == _fr(C, [fr["inner_product"]], True)[0]).all()
    k = g["kzg"]
    powers, _ = _pts(C, k["powers"])
    poly, zz = _fr(C, k["poly"], True), _fr(C, [k["z"]], True)[0]
    pg = eng.srs_register(C.id, powers)
    cxy, _ = eng.kzg_commit(pg, poly)
    assert (cxy == _pts(C, [k["commitment"]])[0][0]).all()
    wxy, _, _ = eng.kzg_open(pg, poly, zz)
    assert (wxy == _pts(C, [k["witness"]])[0][0]).all()
    if hasattr(eng, "ntt"):
        for t in g["ntt"]:
            assert (eng.ntt(C.id, _fr(C, t["coeffs"], True), t["logn"]) == _fr(C, t["evals"], True)).all()


def wire_points(cname):
    C = pyref.Curve(cname)
    g = golden()["curves"][cname]
    pts = [None if m["point"] is None else (int(m["point"][0], 16), int(m["point"][1], 16)) for m in g["multiples"]] + [None]
    return C, g["wire"], pts


def check_wire_oracle(cname):
    C, w, pts = wire_points(cname)
    assert w["count"] == len(pts)
    for compressed, key in ((True, "compressed"), (False, "uncompressed")):
        assert pyref.g1_serialize(C, pts, compressed).hex() == w[key]
        assert pyref.g1_deserialize(C, bytes.fromhex(w[key]), len(pts), compressed) == pts


def check_wire_engine(eng, cname):
    C, w, pts = wire_points(cname)
    xy, inf = C.points_to_limbs(pts)
    for compressed, key in ((True, "compressed"), (False, "uncompressed")):
        assert eng.g1_serialize(C.id, xy, inf, compressed).tobytes().hex() == w[key]
        bxy, binf = eng.g1_deserialize(C.id, bytes.fromhex(w[key]), len(pts), compressed)
        assert (bxy == xy).all() and (binf == inf).all()


def check_row_mul_kat(eng):
    """utils.rs:274-286 test_row_mul, verbatim numbers, on every curve's Fr."""
    kat = golden()["reference_kats"]["row_mul"]
    for cname in ("bls12_381", "bn254", "pallas"):
        C = pyref.Curve(cname)
        m = C.fr_to_limbs([x for row in kat["rows"] for x in row], True)
        v = C.fr_to_limbs(kat["v"], True)
        exp = C.fr_to_limbs(kat["result"], True)
        assert (orc.fr_row_mul(C.id, v, m, 3, 3) == exp).all()
        if eng is not None:
            assert (eng.fr_row_mul(C.id, v, m, 3, 3) == exp).all()
