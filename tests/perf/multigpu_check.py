#!/usr/bin/env python3
"""Run under torchrun on N GPUs (NCCL): index-range-sharded MSM with the NCCL all_gather point-sum
(SURVEY.md 8e partitioning B) against the oracle, at 2^16 (direct) and 2^20 (vs the single-GPU result of rank 0).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/perf/multigpu_check.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import pkgload

pc = pkgload.load()
from poly_commit_b200 import sharded  # noqa: E402
from oracle import orc, pyref  # noqa: E402
from tests import util  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = pc.Engine(local)
    cname = "bls12_381"
    C = pyref.Curve(cname)
    res = {}
    for logn in (16, 20):
        n = (1 << logn) + 1
        beta = util.rand_fr(cname, 1, 1001, mont=True)[0]
        pows = orc.fr_powers_canonical(C.id, beta, n)
        bases = eng.fixed_base_mul(C.id, orc.g1_generator(C.id), pows)      # every rank builds the same SRS
        sc = util.rand_fr_fast(cname, n, seed=77)                           # Montgomery scalars (same on all ranks)
        sm = sharded.ShardedMsm(eng, C.id, bases, dist, flags=pc.SRS_PRECOMPUTE, device=torch.device("cuda", local))
        got = sm.msm(sc, flags=pc.SCALARS_MONT)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            got = sm.msm(sc, flags=pc.SCALARS_MONT)
        torch.cuda.synchronize(); dist.barrier()
        dt = (time.perf_counter() - t0) / reps
        if rank == 0:
            if logn <= 16:
                exp = orc.msm(C.id, bases, orc.field_unop("orc_fr_from_mont", C.id, sc))
            else:
                full = eng.srs_register(C.id, bases, flags=pc.SRS_PRECOMPUTE)
                exp = eng.msm(full, sc, flags=pc.SCALARS_MONT)
            ok = bool((got[0] == exp[0]).all() and got[1] == exp[1])
            res[f"sharded_msm_2p{logn}"] = {"ok": ok, "ms_per_msm_host_buffers": round(dt * 1e3, 3), "world": world,
                                            "scalar_mults_per_s": round(n / dt)}
    # four-step NTT sharded over the ranks (NCCL all-to-all between the passes) against the single-GPU transform
    for logn in (16, 22):
        x = util.rand_fr_fast(cname, (1 << logn) - 3, seed=5)
        sn = sharded.ShardedNtt(eng, C.id, logn, dist, device=torch.device("cuda", local))
        got_n = sn.forward(x)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        got_n = sn.forward(x)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        exp_n = eng.ntt(C.id, x, logn)
        okn = torch.tensor([int(bool((got_n == exp_n).all()))], device="cuda")
        dist.all_reduce(okn, op=dist.ReduceOp.MIN)
        if rank == 0:
            res[f"sharded_ntt_2p{logn}"] = {"ok": bool(okn.item()), "ms_host_buffers": round(dt * 1e3, 3), "world": world}
    # every rank must hold the same result
    h = torch.tensor([int(got[0][0] & np.uint64(0x7FFFFFFF))], device="cuda")
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    if rank == 0:
        res["all_ranks_agree"] = all(int(x.item()) == int(h.item()) for x in hs)
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
