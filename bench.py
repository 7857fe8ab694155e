#!/usr/bin/env python3
"""bench.py -- MarlinKZG10/BLS12-381 commit+open at degree 2^20 (BASELINE.json configs[1]) on N B200s.

A step = one KZG10 commit + one KZG10 open of one degree-2^20 polynomial (2^20+1 uniform Fr coefficients,
hiding_bound=None, degree_bound=None: the protocol of bench-templates/src/lib.rs:69-138) = two G1 MSMs of
2^20(+1) terms, one division by (X - z), with F::into_bigint fused into the MSM digit pass.

  value   whole-job polys/s with the coefficient vectors already resident in HBM (PCGPU_DEVICE_PTRS)
  e2e     the same through the C ABI with pinned HOST buffers (H2D of the coefficients inside the timed region,
          D2H of the two 96-byte points)
  roofline    dominant kernel (bucket accumulate): algorithmic bytes (128 B per scalar-mult) over the
              average launch duration from CUDA events on the launching stream; peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle port (oracle/, OpenMP over Pippenger windows) timed on this box's host cores
  --impl reference   times that CPU path alone (the reference's Rust cannot be built here: no cargo/rustc)

Multi-GPU: ranks shard by polynomial (SURVEY.md 8e partitioning A; the reference's per-polynomial loop,
marlin_pc/mod.rs:192), SRS replicated, no data-path collective; weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_DEG = 20
CURVE = "bls12_381"
ALGO_BYTES_PER_SCALAR_MULT = 128  # 96 B affine base + 32 B scalar
# dram__bytes_read.sum + dram__bytes_write.sum of one MsmAffinePairBody<Bls12381, round 0> launch, from the committed
# ncu --set full capture profiles/r01_final_ncu_prof_pair0_final.txt (5.710451 GB + 2.001377 GB)
NCU_TRAFFIC_BYTES = {20: 5710451000 + 2001377000}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="pcgpu", choices=["pcgpu", "reference"])
    ap.add_argument("--log-deg", type=int, default=LOG_DEG)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=4, help="polynomials in flight per GPU (one context + stream each)")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "timestamp")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """start of the timed region: only samples taken after this moment are reported"""
        import datetime
        self.t_mark = datetime.datetime.now()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        import datetime
        rows = []
        for r in self.rows:
            if len(r) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(r[9], "%Y/%m/%d %H:%M:%S.%f")
            except Exception:
                continue
            if getattr(self, "t_mark", None) is None or ts >= self.t_mark:
                rows.append(r)
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    """threads the CPU arm may use: every core this process is allowed on (torchrun exports OMP_NUM_THREADS=1, which
    would otherwise cripple the OpenMP default)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def oracle_step(orc, cid, bases, coeffs, z, nthreads=0):
    rc, cxy, cinf = orc.kzg_commit(cid, bases, coeffs, nthreads=nthreads)
    assert rc == 0
    rc, wxy, winf, _ = orc.kzg_open(cid, bases, coeffs, z, nthreads=nthreads)
    assert rc == 0
    return cxy, wxy


def pippenger_ops(n):
    """point additions of the oracle's Pippenger (ark-ec window rule) for an n-term 255-bit MSM."""
    import math
    c = 3 if n < 32 else int(math.log2(n)) * 69 // 100 + 2
    w = (255 + c - 1) // c
    return w * (n + 2 * ((1 << c) - 1))


def cpu_reference_run(args, log_deg, steps, warmup, budget_s):
    """The CPU path (oracle port of the reference's dataflow; OpenMP across windows like ark-ec's Rayon MSM)."""
    import numpy as np
    from oracle import orc, pyref
    from tests import util
    C = pyref.Curve(CURVE)
    n_full = (1 << log_deg) + 1
    import math
    c_win = 3 if n_full < 32 else int(math.log2(n_full)) * 69 // 100 + 2
    n_windows = (255 + c_win - 1) // c_win
    # like ark-ec's Rayon MSM the port parallelises over Pippenger windows only, so at most n_windows threads do work
    # (splitting the index range as well was tried: 8 slices x 17 windows on the 128-thread box ran 2x SLOWER -- every
    # slice pays its own 2^c-bucket reduction and the bucket arrays fall out of cache)
    cores = min(host_threads(), n_windows)
    # SRS for the CPU run: random multiples of G (fixed-base batch mul on the host is the slow part, so the
    # base set is 2^14 distinct points tiled -- MSM cost does not depend on the base values)
    tile = 1 << 14
    pts = util.random_points(CURVE, tile, seed=99)
    z = util.rand_fr(CURVE, 1, seed=4, mont=True)[0]

    def run(n):
        reps = (n + tile - 1) // tile
        bases = np.tile(pts, (reps, 1))[:n]
        coeffs = util.rand_fr_fast(CURVE, n, 7)
        t0 = time.perf_counter()
        oracle_step(orc, C.id, bases, coeffs, z, nthreads=host_threads())
        return time.perf_counter() - t0

    # probe at 1/16 size to choose the sample
    t_probe = run((1 << max(log_deg - 4, 8)) + 1)
    est_full = t_probe * pippenger_ops(n_full) / pippenger_ops((1 << max(log_deg - 4, 8)) + 1)
    if est_full * (steps + warmup) <= budget_s:
        n_s, scale, sample = n_full, 1.0, f"full workload: commit+open of one degree-2^{log_deg} polynomial per step"
    else:
        shift = 2
        while shift < 8 and est_full / (1 << shift) * (steps + warmup) > budget_s:
            shift += 1
        n_s = (1 << (log_deg - shift)) + 1
        scale = pippenger_ops(n_full) / pippenger_ops(n_s)
        sample = (f"commit+open of a degree-2^{log_deg - shift} polynomial per step, time scaled x{scale:.2f} by the "
                  f"Pippenger point-addition count to degree 2^{log_deg}")
    for _ in range(warmup):
        run(n_s)
    times = [run(n_s) for _ in range(steps)]
    t = sum(times) / len(times) * scale
    sample += f"; OpenMP over the {n_windows} Pippenger windows ({host_threads()} host threads available)"
    return {"value": 1.0 / t, "unit": "polys/s", "cores": cores, "kind": "port", "sample": sample,
            "ms_per_step": t * 1e3, "msm_scalar_mults_per_s": 2 * n_full / t}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    log_deg = args.log_deg
    workload = f"MarlinKZG10 commit+open, 1 poly/step, degree 2^{log_deg}, BLS12-381, hiding_bound=None, degree_bound=None"

    if args.impl == "reference":
        if rank != 0:
            return
        steps = args.steps if args.steps is not None else 3
        warmup = args.warmup if args.warmup is not None else 1
        r = cpu_reference_run(args, log_deg, steps, warmup, budget_s=float(os.environ.get("PCGPU_REF_BUDGET_S", "200")))
        line = {"impl": "reference", "metric": "MarlinKZG10/BLS12-381 commit+open polys/s at deg 2^20", "value": r["value"],
                "unit": "polys/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64-limb Montgomery (CPU)",
                "data": "synthetic", "config": {"workload": workload, "reference": "CPU oracle port of the reference dataflow "
                "(ark-ec/ark-poly cannot be built here: no Rust toolchain)"},
                "msm_scalar_mults_per_s": r["msm_scalar_mults_per_s"],
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "polys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    steps = args.steps if args.steps is not None else 20
    warmup = args.warmup if args.warmup is not None else 3
    import numpy as np
    import torch
    import pkgload
    pc = pkgload.load()
    from poly_commit_b200 import params  # the oracle is imported by the cpu_baseline leg only (cpu_port_run)

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    eng = pc.Engine(local_rank)  # raises without the CUDA library / an sm_100 device
    cid = pc.CURVES[CURVE]
    n = (1 << log_deg) + 1

    # ---- synthetic SRS on the device: P_i = k_i G for seeded random k_i (g.batch_mul, kzg10/mod.rs:76, with random
    # scalars in place of the powers of beta: the commit / open kernels do not depend on the structure of the bases), then
    # window-folded tables
    pows = torch.from_numpy(params.random_fr(cid, n, 1001).view(np.int64)).cuda()
    d_bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    eng.fixed_base_mul(cid, params.g1_generator(cid), pows.data_ptr(), n=n, flags=pc.DEVICE_PTRS, out=d_bases.data_ptr())
    srs = eng.srs_register(cid, d_bases.data_ptr(), n=n, flags=pc.DEVICE_PTRS | pc.SRS_PRECOMPUTE)
    del pows

    # ---- polynomials: distinct per rank and per step (rotating), pinned host copies + device copies
    n_polys = 4
    host_polys = [torch.from_numpy(params.random_fr(cid, n, 100 + rank * n_polys + i).view(np.int64)).pin_memory() for i in range(n_polys)]
    dev_polys = [h.cuda() for h in host_polys]
    z = params.random_fr(cid, 1, 4)[0]
    # `inflight` independent polynomials are processed concurrently, each on its own context (own stream + workspace),
    # sharing the read-only SRS tables: the latency-bound tails of one MSM overlap the multiply-bound phase of another.
    import threading
    inflight = max(1, args.inflight)
    engines = [eng] + [pc.Engine(local_rank) for _ in range(inflight - 1)]

    def step_dev(i, e=None):
        e = e or eng
        d = dev_polys[i % n_polys]
        c = e.kzg_commit(srs, d.data_ptr(), n=n, flags=pc.DEVICE_PTRS)
        w = e.kzg_open(srs, d.data_ptr(), z, n=n, flags=pc.DEVICE_PTRS)
        return c, w

    def step_host(i, e=None):
        e = e or eng
        h = host_polys[i % n_polys].numpy().view(np.uint64)
        c = e.kzg_commit(srs, h, n=n)
        w = e.kzg_open(srs, h, z, n=n)
        return c, w

    def run_steps(fn, k):
        """k steps spread over the in-flight contexts (thread j takes steps j, j+inflight, ...)."""
        if inflight == 1:
            for i in range(k):
                fn(i)
            return
        def work(j):
            for i in range(j, k, inflight):
                fn(i, engines[j])
        ts = [threading.Thread(target=work, args=(j,)) for j in range(inflight)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the library's calls are synchronous (each ends with a sync of its own stream), so events recorded on the
        # current stream before the first call and after the last one bracket all the work of all contexts
        e0.record()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_steps(fn, k)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), wall_ms)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    sampler.start()                      # started before the warm-up so nvidia-smi is already streaming samples
    run_steps(step_dev, warmup * inflight)
    sampler.mark()
    l0 = eng.launch_count()
    ms_dev = timed(step_dev, steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    eng.profile_enable(False)
    run_steps(step_host, 2 * inflight)
    ms_host = timed(step_host, steps)
    # kernel-level timings: a separate pass with ONE polynomial in flight, CUDA events around every stage on the
    # launching stream (with several contexts in flight the per-stage times would include the other context's kernels)
    eng.profile_enable(True)
    prof_steps = min(steps, 8)
    for i in range(prof_steps):
        step_dev(i)
    acc_ms, acc_cnt = eng.profile_get(4)
    stage_ms = {name: eng.profile_get(s)[0] / max(prof_steps, 1) for s, name in
                enumerate(["digits_count", "scan", "scatter", "tasks", "bucket_accumulate", "bucket_reduce", "final_host", "fr_division"])}
    stage_ms["affine_pair_rounds"] = eng.profile_get(11)[0] / max(prof_steps, 1)
    pair0_ms, pair0_cnt = eng.profile_get(12)
    eng.profile_enable(False)

    if rank != 0:
        return
    polys = steps * world
    value = polys / (ms_dev / 1e3)
    e2e = polys / (ms_host / 1e3)
    peak, peak_src = measured_peaks()
    # compute roofline that actually binds the MSM: wide integer multiply-adds per second (measured live on this GPU)
    imad_peak = eng.measure_imad_peak()
    # multiply count of one 2^log_deg MSM: entries = n * W windows, 3 affine rounds at 6.2 modmuls, the rest XYZZ at 9.5,
    # 288 wide multiplies per 12-limb Montgomery product
    msm_entries = n * 16
    wide_per_msm = (msm_entries * (7.0 / 8.0) * 6.2 + msm_entries * (1.0 / 8.0) * 9.5) * 288
    msm_kernel_ms = (stage_ms["affine_pair_rounds"] + stage_ms["bucket_accumulate"]) / 2
    # dominant kernel = round 0 of the batched-affine pair rounds (one launch per MSM, touches every (base, scalar) pair)
    dom_ms = pair0_ms / pair0_cnt if pair0_cnt else (acc_ms / max(acc_cnt, 1))
    dom_name = "run_kernel_occ<MsmAffinePairBody<Bls12381, true>>" if pair0_cnt else "run_persistent_kernel<MsmAccumulateBody<Bls12381>>"
    achieved = (n * ALGO_BYTES_PER_SCALAR_MULT / 1e9) / (dom_ms / 1e3) if dom_ms else None
    line = {
        "metric": "MarlinKZG10/BLS12-381 commit+open polys/s at deg 2^20", "value": value, "unit": "polys/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32-limb Montgomery (Fq 381-bit x12, Fr 255-bit x8)",
        "data": "synthetic",
        "config": {"workload": workload, "parallelism": f"poly-sharded x{world}, SRS replicated, {inflight} polynomials in flight per GPU",
                   "l2": "per-step working set (window-folded SRS tables 1.6 GB gather + 34 MB coefficients, rotating "
                         "polynomials) exceeds the 126 MB L2; no explicit flush"},
        "msm_scalar_mults_per_s": 2 * n * polys / (ms_dev / 1e3),
        "stage_ms_per_step": stage_ms,
        "e2e": {"value": e2e, "unit": "polys/s", "h2d_bytes_per_step": 2 * n * 32 + 2 * 32, "d2h_bytes_per_step": 2 * 96 + 2 * 4,
                "ms_per_step": ms_host / steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": NCU_TRAFFIC_BYTES.get(log_deg), "peak_source": peak_src,
                     "launch_ms": dom_ms,
                     "compute_roofline": {"bound": "int32 multiply pipe (IMAD.WIDE.U32)", "peak_wide_mul_per_s": imad_peak,
                                          "achieved_wide_mul_per_s": wide_per_msm / (msm_kernel_ms / 1e3) if msm_kernel_ms else None,
                                          "frac": (wide_per_msm / (msm_kernel_ms / 1e3) / imad_peak) if (msm_kernel_ms and imad_peak) else None,
                                          "kernels": "MsmAffinePairBody x3 rounds + MsmAccumulateBody, per MSM"},
                     "note": "MSM is INT32-multiply bound (~3.4k IMAD.WIDE per 128 algorithmic bytes); the HBM fraction "
                             "is reported because north_star asks for it"},
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_run(args, log_deg, steps=1, warmup=0, budget_s=25.0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
