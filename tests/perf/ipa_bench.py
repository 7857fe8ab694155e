#!/usr/bin/env python3
"""Wall-clock of the InnerProductArgPC::open halving loop (cfg3: Pallas, 2^18) through the device-resident round API,
with a per-phase split (l/r MSMs + inner products vs folds)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
pc = pkgload.load()
from poly_commit_b200 import ipa_pc
from oracle import orc, pyref
from tests import util

def main():
    eng = pc.Engine(0)
    cname = "pallas"; C = pyref.Curve(cname); logn = 18; n = 1 << logn
    beta = util.rand_fr(cname, 1, 1001, mont=True)[0]
    key = eng.fixed_base_mul(C.id, orc.g1_generator(C.id), orc.fr_powers_canonical(C.id, beta, n))
    h_prime = util.random_points(cname, 1, seed=41)[0]
    coeffs = util.rand_fr_fast(cname, n, seed=42)
    point = util.rand_fr(cname, 1, seed=43, mont=True)[0]
    ipa_pc.open_rounds(eng, C.id, key, coeffs, point, h_prime, 1)   # warm-up
    t0 = time.perf_counter()
    st = eng.ipa_begin(C.id, key, coeffs, point)
    t_begin = time.perf_counter() - t0
    t_lr = t_fold = t_hash = 0.0
    rc = 7
    per_round = []
    while eng.ipa_len(st) > 1:
        per_round.append([eng.ipa_len(st), 0.0, 0.0])
        t1 = time.perf_counter(); l, li, r, ri = eng.ipa_round_lr(C.id, st, h_prime, with_inf=True); t_lr += time.perf_counter() - t1
        per_round[-1][1] = round((time.perf_counter() - t1) * 1e3, 3)
        t1 = time.perf_counter()
        rc = ipa_pc.compute_random_oracle_challenge(C.id, ipa_pc.round_transcript(eng, C.id, rc, l, li, r, ri)); t_hash += time.perf_counter() - t1
        inv = pow(rc, -1, C.r)
        t1 = time.perf_counter(); eng.ipa_round_fold(st, ipa_pc._fr_mont(C.id, rc), ipa_pc._fr_mont(C.id, inv)); t_fold += time.perf_counter() - t1
        per_round[-1][2] = round((time.perf_counter() - t1) * 1e3, 3)
    t1 = time.perf_counter(); eng.ipa_finish(C.id, st); t_fin = time.perf_counter() - t1
    tot = time.perf_counter() - t0
    print(json.dumps({"workload": "IPA open halving loop, Pallas, 2^18, 18 rounds (host buffers in, device-resident rounds)",
                      "total_ms": round(tot * 1e3, 2), "begin_upload_ms": round(t_begin * 1e3, 2),
                      "lr_msm_ip_ms": round(t_lr * 1e3, 2), "folds_ms": round(t_fold * 1e3, 2),
                      "transcript_ms": round(t_hash * 1e3, 2), "finish_ms": round(t_fin * 1e3, 2), "freeze": os.environ.get("PCGPU_IPA_FREEZE", "1"),
                      "per_round_n_lr_ms_fold_ms": per_round}))

def resident_key():
    """Same open with the committer key RESIDENT in a device buffer (the protocol of the KZG configurations, whose SRS is
    registered once) and the coefficients written to the device inside the timed region."""
    eng = pc.Engine(0)
    cname = "pallas"; C = pyref.Curve(cname); logn = 18; n = 1 << logn
    beta = util.rand_fr(cname, 1, 1001, mont=True)[0]
    key = eng.fixed_base_mul(C.id, orc.g1_generator(C.id), orc.fr_powers_canonical(C.id, beta, n))
    h_prime = util.random_points(cname, 1, seed=41)[0]
    coeffs = util.rand_fr_fast(cname, n, seed=42)
    point = util.rand_fr(cname, 1, seed=43, mont=True)[0]
    d_key = eng.buffer(2 * n); d_key.write(key.reshape(-1, 4))        # 64-byte affine points = two 32-byte elements each
    d_co = eng.buffer(n)
    ref = ipa_pc.open_rounds(eng, C.id, key, coeffs, point, h_prime, 7)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        d_co.write(coeffs)
        st = eng.ipa_begin(C.id, d_key.ptr(), d_co.ptr(), point, n=n, n_coeffs=n, flags=pc.DEVICE_PTRS)
        rc = 7
        ls = []
        while eng.ipa_len(st) > 1:
            l, li, r, ri = eng.ipa_round_lr(C.id, st, h_prime, with_inf=True)
            ls.append(l)
            rc = ipa_pc.compute_random_oracle_challenge(C.id, ipa_pc.round_transcript(eng, C.id, rc, l, li, r, ri))
            eng.ipa_round_fold(st, ipa_pc._fr_mont(C.id, rc), ipa_pc._fr_mont(C.id, pow(rc, -1, C.r)))
        fk, c = eng.ipa_finish(C.id, st)
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None or ms < best else best
    ok = all((a == b).all() for a, b in zip(ls, ref["l_vec"])) and (fk == ref["final_comm_key"]).all()
    print(json.dumps({"workload": "IPA open halving loop, Pallas, 2^18, committer key resident on the device, coefficients uploaded inside the timed region",
                      "total_ms_best_of_3": round(best, 2), "same_proof_as_host_buffer_path": bool(ok)}))


if __name__ == "__main__":
    main()
    resident_key()
