#!/usr/bin/env python3
"""Generates tests/golden/vectors.json from the Python big-integer oracle (oracle/pyref.py) and the
reference's own small-integer KATs.  Committed together with its output.

The reference cannot be run here (Rust, no toolchain) and pins no group-element results in its tests
(SURVEY.md section 8c), so these vectors are produced by an implementation that shares no code with either the C
oracle or the CUDA kernels: affine chord-and-tangent arithmetic on Python ints.  Values are plain integers in
hex (canonical, NOT Montgomery); tests convert to the ABI layout.

Reference KATs embedded verbatim:
  * utils.rs:274-286 test_row_mul: rows [[10,100,4],[23,1,0],[55,58,9]], v = [12,41,55] -> [4088,4431,543]
  * linear_codes/utils.rs:303-331 test_reed_solomon: fft(coeffs)[j] == p(w^j) on the zero-padded domain
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402


def hx(v):
    return hex(v)


def main():
    rnd = random.Random(20260923)
    out = {"generated_by": "tests/golden/make_golden.py (oracle/pyref.py, Python big integers)", "curves": {}}
    for name in ("bls12_381", "bn254", "pallas"):
        C = pyref.Curve(name)
        G = (C.g[0] % C.p, C.g[1] % C.p)
        assert C.on_curve(G) and C.mul(C.r, G) is None
        cur = {"p": hx(C.p), "r": hx(C.r), "generator": [hx(G[0]), hx(G[1])]}
        # small multiples and a few random multiples of G
        ks = [1, 2, 3, 5, C.r - 1, (C.r - 1) // 2, 0xdeadbeef] + [rnd.randrange(C.r) for _ in range(3)]
        cur["multiples"] = [{"k": hx(k), "point": None if (P := C.mul(k, G)) is None else [hx(P[0]), hx(P[1])]} for k in ks]
        # MSM fixture: 12 bases, scalars with edge values
        base_ks = [rnd.randrange(1, C.r) for _ in range(12)]
        bases = [C.mul(k, G) for k in base_ks]
        scalars = [0, 1, C.r - 1, 2, 65535, 65536, 1 << 254 if C.r > (1 << 254) else (1 << 253), 12345] + [rnd.randrange(C.r) for _ in range(4)]
        res = C.msm(bases, scalars)
        # consistency: sum k_i s_i * G
        assert res == C.mul(sum(k * s for k, s in zip(base_ks, scalars)) % C.r, G)
        cur["msm"] = {"bases": [[hx(b[0]), hx(b[1])] for b in bases], "scalars": [hx(s) for s in scalars],
                      "result": None if res is None else [hx(res[0]), hx(res[1])]}
        # Fr: division by (X - z), evaluation, axpy
        coeffs = [rnd.randrange(C.r) for _ in range(37)]
        z = rnd.randrange(C.r)
        q, rem = pyref.poly_div_linear(coeffs, z, C.r)
        assert rem == pyref.poly_eval(coeffs, z, C.r)
        cfac = rnd.randrange(C.r)
        other = [rnd.randrange(C.r) for _ in range(37)]
        cur["fr"] = {"coeffs": [hx(c) for c in coeffs], "z": hx(z), "quotient": [hx(c) for c in q], "remainder": hx(rem),
                     "axpy_c": hx(cfac), "axpy_x": [hx(c) for c in other],
                     "axpy_result": [hx((a + cfac * b) % C.r) for a, b in zip(coeffs, other)],
                     "inner_product": hx(sum(a * b for a, b in zip(coeffs, other)) % C.r)}
        # NTT: reed_solomon semantics (zero-padded, natural order, out[j] = p(w^j)); deg 2^i - 1, rho_inv = 3
        ntt = []
        for i in (1, 2, 3):
            m = 1 << i
            pol = [rnd.randrange(C.r) for _ in range(m)]
            logn = (3 * m - 1).bit_length()  # GeneralEvaluationDomain::new(m * rho_inv) rounds up to a power of two
            w = C.domain_generator(logn)
            ntt.append({"coeffs": [hx(c) for c in pol], "logn": logn, "omega": hx(w),
                        "evals": [hx(v) for v in pyref.ntt_naive(pol, logn, w, C.r)]})
        cur["ntt"] = ntt
        # KZG commit/open on a toy SRS: powers beta^i G
        beta = rnd.randrange(C.r)
        n = 9
        powers = [C.mul(pow(beta, i, C.r), G) for i in range(n)]
        poly = [rnd.randrange(C.r) for _ in range(n)]
        poly[0] = 0  # a leading (low-index) zero coefficient
        zz = rnd.randrange(C.r)
        comm = C.msm(powers, poly)
        wq, _ = pyref.poly_div_linear(poly, zz, C.r)
        wit = C.msm(powers, wq)
        # KZG identity in the exponent: p(beta) - p(z) == (beta - z) * w(beta)
        assert (pyref.poly_eval(poly, beta, C.r) - pyref.poly_eval(poly, zz, C.r)) % C.r == (beta - zz) * pyref.poly_eval(wq, beta, C.r) % C.r
        cur["kzg"] = {"powers": [[hx(b[0]), hx(b[1])] for b in powers], "poly": [hx(c) for c in poly], "z": hx(zz),
                      "commitment": [hx(comm[0]), hx(comm[1])], "witness": [hx(wit[0]), hx(wit[1])]}
        # G1 wire formats (ark-serialize restatement, oracle/pyref.py): the multiples above, compressed and uncompressed.
        # Only the BLS12-381 generator's compressed form is a published vector; the rest pins the restatement against drift.
        wpts = [None if m["point"] is None else (int(m["point"][0], 16), int(m["point"][1], 16)) for m in cur["multiples"]] + [None]
        cur["wire"] = {"compressed": pyref.g1_serialize(C, wpts, True).hex(), "uncompressed": pyref.g1_serialize(C, wpts, False).hex(),
                       "count": len(wpts)}
        out["curves"][name] = cur
    out["reference_kats"] = {"row_mul": {"source": "poly-commit/src/utils.rs:274-286", "rows": [[10, 100, 4], [23, 1, 0], [55, 58, 9]],
                                         "v": [12, 41, 55], "result": [4088, 4431, 543]}}
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote tests/golden/vectors.json")


if __name__ == "__main__":
    main()
