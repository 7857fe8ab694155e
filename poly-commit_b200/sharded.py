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
