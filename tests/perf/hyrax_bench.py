#!/usr/bin/env python3
"""Hyrax commit rows (BASELINE.json cfg4: 22 variables -> 2^11 Pedersen commitments of 2^11 + 1 terms over one com_key, BN254):
fixed-base comb at several window widths (PCGPU_COMB_C), device-resident; rows checked against the bucket pipeline."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pkgload
pc = pkgload.load()
from poly_commit_b200 import params

def main():
    eng = pc.Engine(0)
    cid, dim = pc.BN254, 1 << 11
    n = dim + 1
    ks = torch.from_numpy(params.random_fr(cid, n, 5).view(np.int64)).cuda()
    b = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    F = pc.DEVICE_PTRS
    eng.fixed_base_mul(cid, params.g1_generator(cid), ks.data_ptr(), n=n, flags=F, out=b.data_ptr())
    mat = torch.from_numpy(params.random_fr(cid, dim * n, 6).view(np.int64)).cuda()
    plain = eng.srs_register(cid, b.data_ptr(), n=n, flags=F)
    ref = [eng.msm(plain, mat.data_ptr() + r * n * 32, n=n, flags=F | pc.SCALARS_MONT)[0] for r in (0, 1, dim - 1)]
    for c in sys.argv[1:] or ["8", "12", "14", "16"]:
        os.environ["PCGPU_COMB_C"] = c
        os.environ["PCGPU_COMB_MAX_GB"] = "90"
        t0 = time.perf_counter()
        srs = eng.srs_register(cid, b.data_ptr(), n=n, flags=F | pc.SRS_COMB)
        torch.cuda.synchronize(); t_build = time.perf_counter() - t0
        out, inf = eng.msm_batch(srs, mat.data_ptr(), n, dim, flags=F | pc.SCALARS_MONT)
        ok = all((out[r] == e).all() for r, e in zip((0, 1, dim - 1), ref))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            eng.msm_batch(srs, mat.data_ptr(), n, dim, flags=F | pc.SCALARS_MONT)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / reps * 1e3
        W = (254 + int(c) - 1) // int(c)
        print(json.dumps({"workload": "hyrax commit rows cfg4 (2^11 x (2^11+1), BN254)", "comb_c": int(c), "windows": W,
                          "table_GB": round(n * W * (1 << (int(c) - 1)) * 64 / 1e9, 2), "table_build_s": round(t_build, 3), "ok": ok,
                          "ms": round(ms, 3), "scalar_mults_per_s": round(dim * n / (ms / 1e3))}), flush=True)
        srs.release()
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
