#!/usr/bin/env python3
"""Hardware check of the multi-GPU splits -- run under torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/perf/multigpu_check.py

  B. one MSM sharded by index range, device-resident scalars: fused NVLink point-sum (pcgpu_msm_peer) and the NCCL
     all-gather baseline, against the oracle (2^16) and against the single-GPU MSM (2^20, 2^22); time per MSM, max over ranks
  C. one NTT sharded by the four-step split: pass 1 with the exchange fused into its stores (pcgpu_ntt_pass1_peer + flag
     barrier) and the NCCL all-to-all baseline, against the single-GPU transform (2^16, 2^20, 2^22)
  A. a batch of polynomials sharded by polynomial (commit_batch_sharded) against per-polynomial commits
Prints one JSON object on rank 0."""
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
from poly_commit_b200 import params, sharded  # noqa: E402


def timed(fn, reps, dev):
    """wall time per call, barrier + device sync on both sides, max over ranks"""
    torch.cuda.synchronize(dev); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = pc.Engine(local)
    cid = pc.BLS12_381
    peers = sharded.PeerGroup(eng, dist, device=dev)
    res = {"world": world}
    oracle_logn = 16
    for logn in (16, 20, 22):
        n = (1 << logn) + 1
        ks = torch.from_numpy(params.random_fr(cid, n, 1001).view(np.int64)).to(dev)       # same SRS on every rank
        d_bases = torch.empty((n, 12), dtype=torch.int64, device=dev)
        eng.fixed_base_mul(cid, params.g1_generator(cid), ks.data_ptr(), n=n, flags=pc.DEVICE_PTRS, out=d_bases.data_ptr())
        sc_host = params.random_fr(cid, n, 77)                                             # Montgomery scalars (uniform)
        d_sc = torch.from_numpy(sc_host.view(np.int64)).to(dev)
        fl = pc.SCALARS_MONT | pc.DEVICE_PTRS
        out = {}
        for mode in ("peer", "nccl"):
            sm = sharded.ShardedMsm(eng, cid, d_bases.data_ptr(), dist, flags=pc.SRS_PRECOMPUTE, device=dev, n=n,
                                    peers=peers if mode == "peer" else None, mode=mode)
            lo, hi = sm.local_slice()
            ptr = d_sc.data_ptr() + lo * 32
            got = sm.msm(ptr, flags=fl, n=n)
            out[mode] = got
            res[f"msm_2p{logn}_{mode}_ms"] = round(timed(lambda: sm.msm(ptr, flags=fl, n=n), 5, dev) * 1e3, 3)
            sm.srs.release()
        full = eng.srs_register(cid, d_bases.data_ptr(), n=n, flags=pc.DEVICE_PTRS | pc.SRS_PRECOMPUTE)
        exp = eng.msm(full, d_sc.data_ptr(), n=n, flags=fl)
        res[f"msm_2p{logn}_single_gpu_ms"] = round(timed(lambda: eng.msm(full, d_sc.data_ptr(), n=n, flags=fl), 5, dev) * 1e3, 3)
        ok = all((out[m][0] == exp[0]).all() and out[m][1] == exp[1] for m in out)
        if logn == oracle_logn and rank == 0:
            from oracle import orc
            bases_h = d_bases.cpu().numpy().view(np.uint64)
            e2 = orc.msm(cid, bases_h, orc.field_unop("orc_fr_from_mont", cid, sc_host))
            ok = ok and bool((e2[0] == exp[0]).all())
        okt = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        res[f"msm_2p{logn}_ok"] = bool(okt.item())
        res[f"msm_2p{logn}_speedup_peer"] = round(res[f"msm_2p{logn}_single_gpu_ms"] / res[f"msm_2p{logn}_peer_ms"], 2)
        if logn == 20:
            # A: 2 * world polynomials of degree 2^20 sharded by polynomial, gathered; against per-polynomial commits on rank 0
            npoly = 2 * world
            polys = [torch.from_numpy(params.random_fr(cid, n, 500 + i).view(np.int64)).to(dev) for i in range(npoly)]
            mine = sharded.poly_assignment(npoly, rank, world)
            got_b, _ = sharded.commit_batch_sharded(eng, full, [(polys[i].data_ptr(), n) for i in mine], dist, device=dev,
                                                    flags=pc.DEVICE_PTRS, num_polys=npoly)
            okb = all((eng.kzg_commit(full, polys[i].data_ptr(), n=n, flags=pc.DEVICE_PTRS)[0] == got_b[i]).all() for i in range(npoly))
            okt = torch.tensor([int(okb)], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            res["commit_batch_sharded_ok"] = bool(okt.item())
            del polys
        full.release()
        del d_bases, d_sc, ks
        torch.cuda.empty_cache()

    # C: sharded NTT
    for logn in (16, 20, 22):
        n_in = (1 << logn) - 3
        x = torch.from_numpy(params.random_fr(cid, n_in, 5).view(np.int64)).to(dev)
        exp = torch.empty((1 << logn, 4), dtype=torch.int64, device=dev)
        eng.ntt(cid, x.data_ptr(), logn, n_in=n_in, flags=pc.DEVICE_PTRS, out=exp.data_ptr())
        res[f"ntt_2p{logn}_single_gpu_ms"] = round(timed(lambda: eng.ntt(cid, x.data_ptr(), logn, n_in=n_in, flags=pc.DEVICE_PTRS, out=exp.data_ptr()), 10, dev) * 1e3, 4)
        pn = sharded.PeerNtt.from_group(eng, cid, logn, peers)
        rows = pn.N1 // world
        out_local = torch.empty((pn.N2, rows, 4), dtype=torch.int64, device=dev)
        pn.forward_rank(x.data_ptr(), n_in, out_local.data_ptr())
        # this rank's slice of the natural-order result: X[k1 + N1 k2], k1 = rank * rows + k1_local
        want = exp.view(pn.N2, pn.N1, 4)[:, rank * rows:(rank + 1) * rows, :]
        ok = bool((out_local == want).all())
        res[f"ntt_2p{logn}_peer_ms"] = round(timed(lambda: pn.forward_rank(x.data_ptr(), n_in, out_local.data_ptr()), 10, dev) * 1e3, 4)
        ok = ok and bool((out_local == want).all())
        sn = sharded.ShardedNtt(eng, cid, logn, dist, device=dev)
        o2 = sn.forward_device(x, n_in)
        ok = ok and bool((o2 == want).all())
        res[f"ntt_2p{logn}_nccl_ms"] = round(timed(lambda: sn.forward_device(x, n_in), 10, dev) * 1e3, 4)
        okt = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        res[f"ntt_2p{logn}_ok"] = bool(okt.item())
    peers.close()
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
