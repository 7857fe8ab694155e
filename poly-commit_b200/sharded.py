"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing) -- SURVEY.md section 8e.

A. shard by polynomial  (cfg5: 64 polynomials over 8 GPUs; the reference's serial per-polynomial loop,
   marlin_pc/mod.rs:192): `poly_assignment` + `commit_batch_sharded`; SRS replicated, no data-path collective, the
   commitments are gathered at the end.
B. shard ONE MSM by index range: rank g owns bases[lo_g:hi_g) permanently (its own window-folded tables) and holds the
   matching scalar slice ON THE DEVICE; each rank produces its part of the sum and the parts are added:
     `ShardedMsm(mode="peer")`  the point-sum is fused into the tail of the rank's Pippenger pipeline: bit-plane sums are
                                stored straight into the peers' NVLink-mapped windows (pcgpu_msm_peer); no collective call
     `ShardedMsm(mode="nccl")`  baseline: projective partials (192 bytes) all-gathered with NCCL, summed on every rank
C. shard ONE NTT by the four-step split: rank g transforms its columns (pass 1, step-2 twiddles fused), rows are
   exchanged, rank g transforms its rows (pass 2); the output stays sharded ([k2][k1_local] per rank) for the consumer.
     `PeerNtt`      pass 1 stores every element straight into the owner's row buffer over NVLink (pcgpu_ntt_pass1_peer),
                    epoch flags in the peer windows replace the barrier (pcgpu_peer_signal / pcgpu_peer_wait)
     `ShardedNtt`   baseline: pass 1 into a local matrix, `all_to_all_single`, pass 2
Nothing here stages through the host: inputs and outputs are device pointers / torch tensors.
"""
import numpy as np

from .binding import DEVICE_PTRS, fq_limbs


def shard_range(n, rank, world):
    """Contiguous index range [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def poly_assignment(num_polys, rank, world):
    """Polynomials handled by `rank`: {rank, rank + world, ...} (round robin, like SURVEY 8e partitioning A)."""
    return list(range(rank, num_polys, world))


def all_gather_bytes(arr, dist, device=None):
    """all_gather of equal-sized uint64 arrays through torch.distributed (NCCL on GPUs, gloo on CPU)."""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [o.cpu().numpy().view(np.uint64) for o in outs]


class PeerGroup:
    """One peer window per rank, mapped by every other rank (CUDA IPC over NVLink).  `dist` only carries the 64-byte
    handles at construction; afterwards the ranks talk through the windows.  `extra_bytes` > 0 appends a second,
    larger window (the NTT row buffers)."""

    def __init__(self, engine, dist, device=None):
        import torch
        self.eng, self.dist = engine, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device
        self._mapped = []
        self.local, self.win = self._exchange(engine.peer_window_bytes())
        self.epoch = {}

    def _exchange(self, nbytes):
        import torch
        ptr, handle = self.eng.peer_alloc(nbytes)
        t = torch.from_numpy(handle.copy())
        if self.device is not None:
            t = t.to(self.device)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        win = []
        for r, o in enumerate(outs):
            if r == self.rank:
                win.append(ptr)
            else:
                p = self.eng.peer_open(o.cpu().numpy())
                self._mapped.append(p)
                win.append(p)
        self._owned = getattr(self, "_owned", []) + [ptr]
        return ptr, win

    def alloc_shared(self, nbytes):
        """another symmetric buffer (e.g. NTT row buffers): -> (local pointer, [pointer on every rank])"""
        return self._exchange(nbytes)

    def next_epoch(self, channel):
        self.epoch[channel] = self.epoch.get(channel, 0) + 1
        return self.epoch[channel]

    def barrier(self, channel=1):
        """device-side barrier over the windows' flags"""
        e = self.next_epoch(channel)
        self.eng.peer_signal(self.win, self.rank, channel, e)
        self.eng.peer_wait(self.local, self.world, channel, e)

    def close(self):
        # every rank must have stopped touching the windows before they are unmapped / freed
        self.dist.barrier()
        for p in self._mapped:
            self.eng.peer_close(p)
        self._mapped = []
        self.dist.barrier()
        for p in getattr(self, "_owned", []):
            self.eng.peer_free(p)
        self._owned = []


class ShardedMsm:
    """One MSM split by index range across the ranks of a process group (partitioning B).

    bases: this rank's view of the FULL base array -- a numpy (n, 2*limbs) array, or a device pointer with `n` given
    (DEVICE_PTRS); only rows [lo, hi) are registered (with their own window-folded tables when `flags` says so).
    """

    def __init__(self, engine, curve, bases_xy, dist, rank=None, world=None, flags=0, device=None, n=None, peers=None, mode=None):
        self.eng, self.curve, self.dist, self.device = engine, curve, dist, device
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.peers = peers
        self.mode = mode or ("peer" if peers is not None else "nccl")
        psz = 2 * fq_limbs(curve)
        if isinstance(bases_xy, (int, np.integer)):
            self.n = int(n)
            self.lo, self.hi = shard_range(self.n, self.rank, self.world)
            self.srs = engine.srs_register(curve, int(bases_xy) + self.lo * psz * 8, n=self.hi - self.lo, flags=flags | DEVICE_PTRS)
        else:
            bases_xy = np.ascontiguousarray(bases_xy, dtype=np.uint64).reshape(-1, psz)
            self.n = bases_xy.shape[0]
            self.lo, self.hi = shard_range(self.n, self.rank, self.world)
            self.srs = engine.srs_register(curve, bases_xy[self.lo:self.hi], flags=flags)  # this rank's slice only

    def local_slice(self, n=None):
        n = self.n if n is None else min(self.n, n)
        return min(self.lo, n), min(self.hi, n)

    def msm(self, scalars, flags=0, n=None):
        """scalars: the FULL (n, 4) numpy array (each rank reads only its slice), or -- with DEVICE_PTRS -- a device
        pointer to THIS RANK'S slice (local_slice(n) elements, n = total length).  -> (affine xy, is_identity) on every rank."""
        if flags & DEVICE_PTRS:
            lo, hi = self.local_slice(n)
        else:
            scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
            lo, hi = self.local_slice(scalars.shape[0] if n is None else n)
            scalars = scalars[lo:hi]
        if self.mode == "peer":
            return self.eng.msm_peer(self.srs, scalars, self.peers.win, self.rank, self.peers.next_epoch(0), n=hi - lo, flags=flags)
        part = self.eng.msm_partial(self.srs, scalars, n=hi - lo, flags=flags)
        parts = all_gather_bytes(part, self.dist, self.device)
        return self.eng.g1_sum_xyzz(self.curve, np.concatenate(parts))


def commit_batch_sharded(engine, srs, polys, dist, device=None, flags=0, num_polys=None):
    """Partitioning A (cfg5): this rank commits to ITS polynomials (`polys`: list of arrays or (device_ptr, n) tuples, in the
    order of poly_assignment(num_polys, rank, world)) with pcgpu_kzg_commit_batch; the commitments are all-gathered so that
    every rank returns the full (num_polys, 2*limbs) array in polynomial order."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    num_polys = len(polys) * world if num_polys is None else num_polys
    mine = poly_assignment(num_polys, rank, world)
    assert len(mine) == len(polys)
    per = (num_polys + world - 1) // world
    nq = 2 * fq_limbs(srs.curve)
    out, inf = engine.kzg_commit_batch(srs, polys, flags=flags) if polys else (np.zeros((0, nq), dtype=np.uint64), np.zeros(0, dtype=np.uint8))
    buf = np.zeros((per, nq + 1), dtype=np.uint64)
    buf[:len(mine), :nq] = out
    buf[:len(mine), nq] = inf
    t = torch.from_numpy(buf.view(np.int64))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.zeros((num_polys, nq), dtype=np.uint64)
    finf = np.zeros(num_polys, dtype=np.uint8)
    for r, o in enumerate(outs):
        a = o.cpu().numpy().view(np.uint64)
        for j, i in enumerate(poly_assignment(num_polys, r, world)):
            full[i], finf[i] = a[j, :nq], a[j, nq]
    return full, finf


class ShardedNtt:
    """Four-step NTT over the ranks of a process group with an all-to-all between the passes (the collective baseline of
    PeerNtt).  Buffers are torch tensors on `device` (NCCL) or CPU tensors (gloo, host-emulation tests)."""

    def __init__(self, engine, curve, logn, dist, device=None):
        self.eng, self.curve, self.logn, self.dist, self.device = engine, curve, logn, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.m1, self.m2 = engine.ntt_split(logn)
        if self.m2 == 0:
            raise ValueError("transform too small to shard (single block pass)")
        self.N1, self.N2 = 1 << self.m1, 1 << self.m2
        if self.N1 % self.world or self.N2 % self.world:
            raise ValueError("world size must divide both factors")

    def forward_device(self, x, n_in, inverse=False):
        """x: torch int64 tensor (>= n_in rows of 4) holding the input on this rank's device -> torch tensor (N2, rows, 4):
        this rank's slice [k2][k1_local] of the natural-order output X[k1 + N1 k2], k1 = rank * rows + k1_local."""
        import torch
        W, N1, N2 = self.world, self.N1, self.N2
        cols, rows = N2 // W, N1 // W
        dev = x.device
        sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
        sync()
        a = torch.empty((N1, cols, 4), dtype=torch.int64, device=dev)                   # local A[k1][n2 - lo]
        self.eng.ntt_pass(self.curve, self.logn, 1, self.rank * cols, cols, x.data_ptr(), n_in, a.data_ptr(), inverse)
        send = a.view(W, rows, cols, 4)
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv.view(-1), send.view(-1))
        # recv[src][k1_local][n2_local] -> rows[k1_local][src * cols + n2_local]
        rowbuf = recv.permute(1, 0, 2, 3).contiguous()
        out_local = torch.empty((N2, rows, 4), dtype=torch.int64, device=dev)           # [k2][k1_local]
        sync()
        self.eng.ntt_pass(self.curve, self.logn, 2, self.rank * rows, rows, rowbuf.data_ptr(), rows * N2, out_local.data_ptr(), inverse)
        return out_local

    def gather(self, out_local):
        """assemble the natural-order output on every rank (tests / small sizes only: this is the all-gather a consumer of
        sharded output does not need)"""
        import torch
        W = self.world
        gathered = [torch.empty_like(out_local) for _ in range(W)]
        self.dist.all_gather(gathered, out_local)
        full = torch.stack(gathered, dim=0).permute(1, 0, 2, 3).contiguous()            # [k2][r][k1_local] = natural order
        return full.reshape(self.N1 * self.N2, 4).cpu().numpy().view(np.uint64)

    def forward(self, coeffs, inverse=False):
        """coeffs: (n_in, 4) uint64 on the host (the full input on every rank) -> (2^logn, 4) uint64 natural-order output on
        every rank.  Convenience wrapper for tests; benchmarks call forward_device."""
        import torch
        dev = self.device if self.device is not None else torch.device("cpu")
        x = torch.from_numpy(np.ascontiguousarray(coeffs, dtype=np.uint64).view(np.int64).reshape(-1, 4).copy()).to(dev)
        return self.gather(self.forward_device(x, x.shape[0], inverse))


class PeerNtt:
    """Four-step NTT with the exchange fused into pass 1's stores (`pcgpu_ntt_pass1_peer`): every pass-1 block writes its
    column's elements straight into the owners' row buffers over NVLink peer mappings -- no staging buffer and no separate
    all-to-all; epoch flags replace the barrier before pass 2.

    Two drivers share the kernels:
      * multi-process (`PeerNtt.from_group(engine, curve, logn, peers)`): one rank per process, row buffers are a symmetric
        allocation of the PeerGroup; `forward_rank` runs this rank's part.
      * in-process (`PeerNtt(engines, curve, logn)` + `forward`): one process drives all "devices" -- the host-emulation
        unit test (tests/test_hostcheck.py::test_ntt_pass1_with_fused_exchange), where pointers are plain host buffers."""

    def __init__(self, engines, curve, logn):
        self.engines, self.curve, self.logn = engines, curve, logn
        self.world = len(engines)
        self._dims(engines[0])

    def _dims(self, eng):
        self.m1, self.m2 = eng.ntt_split(self.logn)
        if self.m2 == 0:
            raise ValueError("transform too small to shard (single block pass)")
        self.N1, self.N2 = 1 << self.m1, 1 << self.m2
        if self.N1 % self.world or self.N2 % self.world:
            raise ValueError("world size must divide both factors")

    @classmethod
    def from_group(cls, engine, curve, logn, peers):
        self = cls.__new__(cls)
        self.engines, self.curve, self.logn, self.world, self.peers = [engine], curve, logn, peers.world, peers
        self._dims(engine)
        rows = self.N1 // self.world
        # two row buffers, used alternately: a rank may start storing transform k+1 while slower peers are still reading
        # transform k's rows in pass 2; buffer reuse (k+2) is safe because every rank signals the barrier of transform k+1
        # only after its own pass 2 of transform k
        self.rowbufs = [peers.alloc_shared(rows * self.N2 * 32) for _ in range(2)]
        self.count = 0
        return self

    def forward_rank(self, in_ptr, n_in, out_ptr, inverse=False):
        """this rank's part: in_ptr = the (zero-padded at n_in) input on this device, out_ptr = N2*rows result slice
        [k2][k1_local].  Device pointers; returns when the slice is complete."""
        p, eng = self.peers, self.engines[0]
        rows, cols = self.N1 // self.world, self.N2 // self.world
        row_local, row_ptrs = self.rowbufs[self.count & 1]
        self.count += 1
        eng.ntt_pass1_peer(self.curve, self.logn, p.rank * cols, cols, in_ptr, n_in, row_ptrs, inverse=inverse)
        p.barrier(channel=1)
        eng.ntt_pass(self.curve, self.logn, 2, p.rank * rows, rows, row_local, rows * self.N2, out_ptr, inverse=inverse)

    def forward(self, in_ptrs, n_in, row_ptrs, out_ptrs, inverse=False, sync=None):
        """in-process driver: in_ptrs[g] / row_ptrs[g] / out_ptrs[g] live on device g; `sync()` drains all devices"""
        rows, cols = self.N1 // self.world, self.N2 // self.world
        for g, e in enumerate(self.engines):
            e.ntt_pass1_peer(self.curve, self.logn, g * cols, cols, in_ptrs[g], n_in, row_ptrs, inverse=inverse)
        if sync is not None:
            sync()
        for g, e in enumerate(self.engines):
            e.ntt_pass(self.curve, self.logn, 2, g * rows, rows, row_ptrs[g], rows * self.N2, out_ptrs[g], inverse=inverse)
