#!/usr/bin/env python3
"""Where does a small MSM spend its time?  Wall clock per call and the per-stage device times (pcgpu_profile_get)
for n = 2^8 .. 2^14, raw bases (no window folding) -- the shape of IPA's late rounds and of cfg1 (degree 2^10)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
pc = pkgload.load()
from oracle import orc, pyref
from tests import util

STAGES = ["digits", "scan", "scatter", "tasks", "accumulate", "reduce", "host_tail", "div", "axpy", "ntt", "comb", "affine", "pair0"]


def main():
    eng = pc.Engine(0)
    out = {}
    for cname in ("bls12_381", "pallas"):
        C = pyref.Curve(cname)
        nmax = 1 << 14
        beta = util.rand_fr(cname, 1, 1001, mont=True)[0]
        bases = eng.fixed_base_mul(C.id, orc.g1_generator(C.id), orc.fr_powers_canonical(C.id, beta, nmax))
        srs = eng.srs_register(C.id, bases)
        for logn in (8, 10, 12, 14):
            n = 1 << logn
            sc = util.rand_fr_fast(cname, n, seed=3)
            for _ in range(3):
                eng.msm(srs, sc, flags=pc.SCALARS_MONT)
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                eng.msm(srs, sc, flags=pc.SCALARS_MONT)
            wall = (time.perf_counter() - t0) / reps
            eng.profile_enable(True)
            for _ in range(reps):
                eng.msm(srs, sc, flags=pc.SCALARS_MONT)
            st = {}
            for i, name in enumerate(STAGES):
                ms, cnt = eng.profile_get(i)
                if cnt:
                    st[name] = round(ms / reps, 4)
            eng.profile_enable(False)
            out[f"{cname}_2p{logn}"] = {"wall_ms": round(wall * 1e3, 4), "stages_ms": st}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
