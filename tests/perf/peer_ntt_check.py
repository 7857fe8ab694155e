#!/usr/bin/env python3
"""NVLink run of the fused-exchange NTT (pcgpu_ntt_pass1_peer + pcgpu_ntt_pass(2), sharded.PeerNtt): ONE process drives all
visible GPUs, one Engine per device, exchange buffers peer-mapped with cudaDeviceEnablePeerAccess; result compared with the
single-GPU transform and timed with CUDA events next to the NCCL all-to-all variant's two kernels.
  gpurun --gpus 2 -- python tests/perf/peer_ntt_check.py
STATUS: written at the end of round 1 when no GPU time was left -- NOT YET RUN on hardware (the kernel and the host driver it
calls are verified under host emulation, tests/test_hostcheck.py::test_ntt_pass1_with_fused_exchange)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import pkgload

pc = pkgload.load()
from poly_commit_b200 import params, sharded  # noqa: E402


def enable_peer_access(world):
    from cuda import cudart
    for a in range(world):
        cudart.cudaSetDevice(a)
        for b in range(world):
            if a != b:
                err, can = cudart.cudaDeviceCanAccessPeer(a, b)
                assert int(can) == 1, f"device {a} cannot access device {b}"
                (err,) = cudart.cudaDeviceEnablePeerAccess(b, 0)
                assert int(err) in (0, 704), err          # 704 = already enabled


def main():
    world = torch.cuda.device_count()
    assert world >= 2, "needs at least two GPUs"
    enable_peer_access(world)
    cid = pc.BLS12_381
    res = {"world": world}
    engines = [pc.Engine(g) for g in range(world)]
    for logn in (16, 20, 22):
        n_in = (1 << logn) - 3
        x = params.random_fr(cid, n_in, 5)
        pn = sharded.PeerNtt(engines, cid, logn)
        rows = pn.N1 // world
        ins, rowbufs, outs = [], [], []
        for g in range(world):
            with torch.cuda.device(g):
                ins.append(torch.from_numpy(x.view(np.int64)).cuda(g))
                rowbufs.append(torch.zeros((rows, pn.N2, 4), dtype=torch.int64, device=f"cuda:{g}"))
                outs.append(torch.empty((pn.N2, rows, 4), dtype=torch.int64, device=f"cuda:{g}"))

        def sync_all():
            for g in range(world):
                torch.cuda.synchronize(g)

        def run():
            pn.forward([t.data_ptr() for t in ins], n_in, [t.data_ptr() for t in rowbufs], [t.data_ptr() for t in outs], sync=sync_all)
            sync_all()

        run()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            run()
        dt = (time.perf_counter() - t0) / reps
        got = torch.stack([o.cpu() for o in outs], 0).permute(1, 0, 2, 3).contiguous().reshape(-1, 4).numpy().view(np.uint64)
        exp = engines[0].ntt(cid, x, logn)
        res[f"peer_ntt_2p{logn}"] = {"ok": bool((got == exp).all()), "ms_device_resident": round(dt * 1e3, 3)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
