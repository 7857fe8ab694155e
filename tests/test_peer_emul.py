"""The fused multi-GPU paths (csrc/peer.cuh, sharded.PeerGroup / ShardedMsm(mode="peer") / PeerNtt.from_group) under host
emulation: every "rank" is a host THREAD with its own context of tests/host_emul/libpcgpu_hostcheck.so, peer windows are
plain host allocations, flags are spun on with a wall-clock budget.  Checks the record format, the epoch protocol, slice
bookkeeping and the fallbacks (small slices, unfolded tables) against the oracle.  The NVLink run is tests/perf/multigpu_check.py."""
import threading

import numpy as np
import pytest

from oracle import orc, pyref
from tests import util


class ThreadDist:
    """the four torch.distributed calls PeerGroup / ShardedMsm need, for ranks that are threads of this process"""

    class _Shared:
        def __init__(self, world):
            self.world, self.barrier, self.slots = world, threading.Barrier(world), [None] * world

    def __init__(self, shared, rank):
        self.s, self.rank = shared, rank

    def get_rank(self):
        return self.rank

    def get_world_size(self):
        return self.s.world

    def barrier(self):
        self.s.barrier.wait()

    def all_gather(self, outs, t):
        self.s.slots[self.rank] = t.clone()
        self.s.barrier.wait()
        for r in range(self.s.world):
            outs[r].copy_(self.s.slots[r])
        self.s.barrier.wait()


def _run_ranks(world, fn):
    shared = ThreadDist._Shared(world)
    res, errs = [None] * world, []

    def work(r):
        try:
            res[r] = fn(r, ThreadDist(shared, r))
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not errs, errs
    return res


@pytest.mark.parametrize("cname,n,world,flags_name", [("bls12_381", 9000, 2, "SRS_PRECOMPUTE"), ("bn254", 12400, 3, None),
                                                      ("pallas", 50, 2, None)])
def test_msm_peer_threads(pc, hostcheck_path, cname, n, world, flags_name):
    """index-sharded MSM with the point-sum pushed through the peer windows: folded tables (bit planes travel), unfolded
    tables (record too large -> one-plane fallback) and slices below the small-MSM threshold; two MSMs back to back so the
    epoch protocol is exercised; every rank must return the oracle's point."""
    from poly_commit_b200 import sharded
    C = pyref.Curve(cname)
    bases = util.random_points(cname, n, seed=11)
    sc = [util.rand_fr(cname, n, seed=12 + k, mont=False) for k in range(2)]
    sc[1][: n // 2] = 0                                                       # second call: half the scalars are zero
    exp = [orc.msm(C.id, bases, s) for s in sc]
    flags = getattr(pc, flags_name) if flags_name else 0

    def rank_fn(r, dist):
        eng = pc.Engine(0, lib_path=hostcheck_path)
        peers = sharded.PeerGroup(eng, dist)
        sm = sharded.ShardedMsm(eng, C.id, bases, dist, flags=flags, peers=peers)
        out = [sm.msm(s) for s in sc]
        short = sm.msm(sc[0][: n // 3])                                       # later ranks hold an empty slice
        peers.close()
        eng.close()
        return out, short

    res = _run_ranks(world, rank_fn)
    exp_short = orc.msm(C.id, bases, sc[0][: n // 3])
    for out, short in res:
        for got, e in zip(out, exp):
            assert (got[0] == e[0]).all() and got[1] == e[1]
        assert (short[0] == exp_short[0]).all()


def test_msm_peer_missing_rank_times_out(pc, hostcheck_path, monkeypatch):
    """a peer that never arrives yields PCGPU_E_PEER after the bounded wait, not a hang (emulation budget shortened)"""
    from poly_commit_b200 import sharded
    monkeypatch.setenv("PCGPU_EMUL_PEER_WAIT_MS", "300")
    C = pyref.Curve("bn254")
    bases = util.random_points("bn254", 40, seed=13)
    sc = util.rand_fr("bn254", 40, seed=14, mont=False)

    def rank_fn(r, dist):
        eng = pc.Engine(0, lib_path=hostcheck_path)
        peers = sharded.PeerGroup(eng, dist)
        sm = sharded.ShardedMsm(eng, C.id, bases, dist, peers=peers)
        err = None
        if r == 0:                                                            # rank 1 never calls msm
            with pytest.raises(pc.binding.PcgpuError) as ei:
                sm.msm(sc)
            err = ei.value.code
        peers.close()
        eng.close()
        return err

    res = _run_ranks(2, rank_fn)
    assert res[0] == -9


@pytest.mark.parametrize("cname,logn,world", [("bls12_381", 12, 2), ("bn254", 13, 4)])
def test_peer_ntt_threads(pc, hostcheck_path, cname, logn, world):
    """PeerNtt.from_group / forward_rank: pass 1 stores into the owners' row buffers, flag barrier, pass 2; three transforms
    in a row (alternating row buffers) against the single-call transform"""
    from poly_commit_b200 import sharded
    C = pyref.Curve(cname)
    xs = [util.rand_fr(cname, (1 << logn) - 3 * k, seed=20 + k, mont=True) for k in range(3)]
    ref_eng = pc.Engine(0, lib_path=hostcheck_path)
    exp = [ref_eng.ntt(C.id, x, logn) for x in xs]
    ref_eng.close()

    def rank_fn(r, dist):
        eng = pc.Engine(0, lib_path=hostcheck_path)
        peers = sharded.PeerGroup(eng, dist)
        pn = sharded.PeerNtt.from_group(eng, C.id, logn, peers)
        rows = pn.N1 // world
        outs = []
        for x in xs:
            o = np.zeros((pn.N2, rows, 4), dtype=np.uint64)
            pn.forward_rank(x.ctypes.data, x.shape[0], o.ctypes.data)
            outs.append(o)
        peers.close()
        eng.close()
        return outs

    res = _run_ranks(world, rank_fn)
    for k in range(3):
        got = np.stack([res[r][k] for r in range(world)], 0).transpose(1, 0, 2, 3).reshape(-1, 4)
        assert (got == exp[k]).all()
