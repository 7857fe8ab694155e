"""Host-side multi-GPU logic with world_size 2 over gloo on CPU: index-range sharding of one MSM +
all_gather point-sum (SURVEY 8e partitioning B) and polynomial assignment (partitioning A).  Each rank drives
the host-emulation build of the C ABI (tests/host_emul) -- the device kernels themselves are covered by the
GPU parity tests."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, lib_path, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pkgload
    pc = pkgload.load()
    from poly_commit_b200 import sharded
    from oracle import orc, pyref
    from tests import util
    eng = pc.Engine(0, lib_path=lib_path)
    ok = True
    for cname, n in (("bls12_381", 101), ("bn254", 64)):
        C = pyref.Curve(cname)
        bases = util.random_points(cname, n, seed=3)
        sc = util.rand_fr(cname, n, seed=4, mont=False)
        sm = sharded.ShardedMsm(eng, C.id, bases, dist)
        got = sm.msm(sc)
        exp = orc.msm(C.id, bases, sc)
        ok &= bool((got[0] == exp[0]).all() and got[1] == exp[1])
        # fewer scalars than bases: the last rank's slice may be empty
        got = sm.msm(sc[:40])
        exp = orc.msm(C.id, bases, sc[:40])
        ok &= bool((got[0] == exp[0]).all())
    # partitioning A: every polynomial handled exactly once, commitments gathered to all ranks
    polys = list(range(7))
    mine = sharded.poly_assignment(len(polys), rank, world)
    C = pyref.Curve("bls12_381")
    powers = util.synthetic_srs("bls12_381", 33, seed=2)
    pg = eng.srs_register(C.id, powers)
    out = np.zeros((len(polys), 12), dtype=np.uint64)
    for i in mine:
        out[i] = eng.kzg_commit(pg, util.rand_fr("bls12_381", 33, seed=200 + i, mont=True))[0]
    gathered = sharded.all_gather_bytes(out, dist)
    total = sum(g.reshape(len(polys), 12) for g in gathered)  # each row non-zero on exactly one rank
    for i in polys:
        rc, exy, _ = orc.kzg_commit(C.id, powers, util.rand_fr("bls12_381", 33, seed=200 + i, mont=True))
        ok &= bool((total[i] == exy).all())
    # four-step NTT sharded over the ranks with an all-to-all between the passes (host emulation: device ptr == host ptr)
    for cname, logn, n_in in (("bls12_381", 12, 4000), ("bn254", 13, 8192)):
        C = pyref.Curve(cname)
        x = util.rand_fr(cname, n_in, seed=9, mont=True)
        sn = sharded.ShardedNtt(eng, C.id, logn, dist)
        got = sn.forward(x)
        ok &= bool((got == orc.fr_ntt(C.id, x, logn)).all())
        back = sn.forward(got, inverse=True)
        ok &= bool((back[:n_in] == x).all() and not back[n_in:].any())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_range_and_assignment(pc):
    from poly_commit_b200 import sharded
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 1):
        for world in (1, 2, 4, 8):
            rs = [sharded.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    assert sorted(sum((sharded.poly_assignment(64, r, 8) for r in range(8)), [])) == list(range(64))


def test_sharded_msm_gloo_world2(pc, hostcheck_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, hostcheck_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
