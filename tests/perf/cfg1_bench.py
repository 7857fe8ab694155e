#!/usr/bin/env python3
"""BASELINE.json configs[0] (the reference's own CPU-runnable case): MarlinKZG10 / BLS12-381, degree 2^10, commit + open of
one polynomial through the C ABI with host buffers (synchronous calls, one context) next to the CPU oracle port on the same
box.  Latency-bound on both sides: this is the small-MSM path (csrc/msm_small.cuh), not the bucket pipeline."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
pc = pkgload.load()
from oracle import orc, pyref
from tests import util


def main():
    eng = pc.Engine(0)
    cname = "bls12_381"; C = pyref.Curve(cname)
    n = (1 << 10) + 1
    beta = util.rand_fr(cname, 1, 1001, mont=True)[0]
    bases = eng.fixed_base_mul(C.id, orc.g1_generator(C.id), orc.fr_powers_canonical(C.id, beta, n))
    srs = eng.srs_register(C.id, bases, flags=pc.SRS_PRECOMPUTE)
    poly = util.rand_fr_fast(cname, n, seed=5)
    z = util.rand_fr(cname, 1, seed=6, mont=True)[0]
    for _ in range(5):
        comm = eng.kzg_commit(srs, poly); w = eng.kzg_open(srs, poly, z)
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps):
        comm = eng.kzg_commit(srs, poly); w = eng.kzg_open(srs, poly, z)
    gpu = (time.perf_counter() - t0) / reps
    rc, exy, einf = orc.kzg_commit(C.id, bases, poly)
    assert rc == 0 and (exy == comm[0]).all()
    rc, wxy, winf, _ = orc.kzg_open(C.id, bases, poly, z)
    assert rc == 0 and (wxy == w[0]).all()
    creps = 20
    t0 = time.perf_counter()
    for _ in range(creps):
        orc.kzg_commit(C.id, bases, poly); orc.kzg_open(C.id, bases, poly, z)
    cpu = (time.perf_counter() - t0) / creps
    print(json.dumps({"workload": "MarlinKZG10 commit+open, degree 2^10, BLS12-381 (cfg1), one polynomial, synchronous C-ABI calls, host buffers",
                      "gpu_ms_per_poly": round(gpu * 1e3, 3), "gpu_polys_per_s": round(1 / gpu, 1),
                      "cpu_port_ms_per_poly": round(cpu * 1e3, 3), "cpu_polys_per_s": round(1 / cpu, 1),
                      "cpu_threads": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count()))}))


if __name__ == "__main__":
    main()
