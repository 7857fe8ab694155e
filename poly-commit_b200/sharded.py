"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing) -- SURVEY.md section 8e.

A. shard by polynomial  (cfg5: 64 polynomials over 8 GPUs; the reference's serial per-polynomial loop,
   marlin_pc/mod.rs:192): `poly_assignment`; SRS replicated, no data-path collective, results gathered.
B. shard ONE MSM by index range: rank g owns bases[lo_g:hi_g) permanently and receives the matching scalar
   slice; each rank produces a projective (XYZZ) partial; the "NCCL point-sum" is an all_gather of the
   world_size x 192-byte partials followed by a local sum on every rank (NCCL has no reduction operator for
   elliptic-curve points).  `ShardedMsm`.
"""
import numpy as np

from .binding import fq_limbs


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


class ShardedMsm:
    """One MSM split by index range across the ranks of a process group (partitioning B)."""

    def __init__(self, engine, curve, bases_xy, dist, rank=None, world=None, flags=0, device=None):
        self.eng, self.curve, self.dist, self.device = engine, curve, dist, device
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        bases_xy = np.ascontiguousarray(bases_xy, dtype=np.uint64).reshape(-1, 2 * fq_limbs(curve))
        self.n = bases_xy.shape[0]
        self.lo, self.hi = shard_range(self.n, self.rank, self.world)
        self.srs = engine.srs_register(curve, bases_xy[self.lo:self.hi], flags=flags)  # this rank's slice only

    def msm(self, scalars, flags=0):
        """scalars: the FULL (n, 4) array (each rank reads only its slice) -> (affine xy, is_identity) on every rank."""
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        n = min(self.n, scalars.shape[0])
        lo, hi = min(self.lo, n), min(self.hi, n)
        part = self.eng.msm_partial(self.srs, scalars[lo:hi], n=hi - lo, flags=flags)
        parts = all_gather_bytes(part, self.dist, self.device)
        return self.eng.g1_sum_xyzz(self.curve, np.concatenate(parts))


class ShardedNtt:
    """Four-step NTT over the ranks of a process group: rank g transforms the columns n2 of its slice
    (pass 1, step-2 twiddles fused), an all-to-all hands every rank whole rows k1, pass 2 transforms them, an all-gather
    assembles the natural-order output on every rank.  `xp` abstracts where buffers live: torch CUDA tensors with NCCL on
    GPUs; plain CPU tensors with gloo in the host-emulation tests ("device" pointers are then host pointers)."""

    def __init__(self, engine, curve, logn, dist, device=None):
        self.eng, self.curve, self.logn, self.dist, self.device = engine, curve, logn, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.m1, self.m2 = engine.ntt_split(logn)
        if self.m2 == 0:
            raise ValueError("transform too small to shard (single block pass)")
        self.N1, self.N2 = 1 << self.m1, 1 << self.m2
        if self.N1 % self.world or self.N2 % self.world:
            raise ValueError("world size must divide both factors")

    def forward(self, coeffs, inverse=False):
        """coeffs: (n_in, 4) uint64 (the full input on every rank) -> (2^logn, 4) uint64 natural-order output on every rank."""
        import torch
        W, N1, N2 = self.world, self.N1, self.N2
        cols, rows = N2 // W, N1 // W
        dev = self.device if self.device is not None else torch.device("cpu")
        x = torch.from_numpy(np.ascontiguousarray(coeffs, dtype=np.uint64).view(np.int64).reshape(-1, 4).copy()).to(dev)
        n_in = x.shape[0]
        # the engine launches on its own stream and synchronises it before returning; torch / NCCL work in between runs on
        # torch's current stream, so drain that before handing pointers to the engine
        sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
        sync()
        a = torch.empty((N1, cols, 4), dtype=torch.int64, device=dev)                   # local A[k1][n2 - lo]
        self.eng.ntt_pass(self.curve, self.logn, 1, self.rank * cols, cols, x.data_ptr(), n_in, a.data_ptr(), inverse)
        # all-to-all: block r of my rows (k1 in rank r's range) goes to rank r
        send = a.reshape(W, rows, cols, 4).contiguous()
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv.view(-1), send.view(-1))
        # recv[src][k1_local][n2_local] -> rows[k1_local][src * cols + n2_local]
        rowbuf = recv.permute(1, 0, 2, 3).contiguous().reshape(rows, N2, 4)
        out_local = torch.empty((N2, rows, 4), dtype=torch.int64, device=dev)           # [k2][k1_local]
        sync()
        self.eng.ntt_pass(self.curve, self.logn, 2, self.rank * rows, rows, rowbuf.data_ptr(), rows * N2, out_local.data_ptr(), inverse)
        gathered = [torch.empty_like(out_local) for _ in range(W)]
        self.dist.all_gather(gathered, out_local)
        full = torch.stack(gathered, dim=0).permute(1, 0, 2, 3).contiguous()            # [k2][r][k1_local] = natural order
        return full.reshape(N1 * N2, 4).cpu().numpy().view(np.uint64)


class PeerNtt:
    """Four-step NTT over the GPUs of ONE process with the exchange fused into pass 1's stores (`pcgpu_ntt_pass1_peer`):
    engine g runs on device g, owns `rows = N1 / world` rows, and every pass-1 block writes its column's elements straight
    into the owners' row buffers over NVLink peer mappings -- no staging buffer and no separate all-to-all; pass 2 starts
    after one device-wide synchronisation.  `alloc(rank, nbytes)` returns a device pointer on device `rank` that every other
    device can store to (the caller enables peer access, e.g. torch tensors after `cudaDeviceEnablePeerAccess`); `engines`
    are `Engine` objects, one per device.  Under host emulation all "devices" are the host and the pointers are numpy buffers
    (tests/test_hostcheck.py::test_ntt_pass1_with_fused_exchange).  STATUS: the kernel and this host logic are verified under
    emulation only; the multi-GPU run over NVLink is the first item of the next round."""

    def __init__(self, engines, curve, logn):
        self.engines, self.curve, self.logn = engines, curve, logn
        self.world = len(engines)
        self.m1, self.m2 = engines[0].ntt_split(logn)
        if self.m2 == 0:
            raise ValueError("transform too small to shard (single block pass)")
        self.N1, self.N2 = 1 << self.m1, 1 << self.m2
        if self.N1 % self.world or self.N2 % self.world:
            raise ValueError("world size must divide both factors")

    def forward(self, in_ptrs, n_in, row_ptrs, out_ptrs, inverse=False, sync=None):
        """in_ptrs[g]: the (zero-padded at n_in) input on device g; row_ptrs[g]: rows*N2-element exchange buffer on device g,
        writable from every device; out_ptrs[g]: N2*rows-element result slice [k2][k1_local] on device g.  `sync()` must drain
        all devices (between the passes every buffer has to be complete)."""
        rows, cols = self.N1 // self.world, self.N2 // self.world
        for g, e in enumerate(self.engines):
            e.ntt_pass1_peer(self.curve, self.logn, g * cols, cols, in_ptrs[g], n_in, row_ptrs, inverse=inverse)
        if sync is not None:
            sync()
        for g, e in enumerate(self.engines):
            e.ntt_pass(self.curve, self.logn, 2, g * rows, rows, row_ptrs[g], rows * self.N2, out_ptrs[g], inverse=inverse)
