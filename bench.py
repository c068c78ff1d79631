#!/usr/bin/env python3
"""Benchmark of the hot path on BASELINE.json's headline workload (config #4: candidate acquisition sweep).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full weighted_distance_merit pass (candidate_srbf.py:31-55) over one batch of DYCORS-shaped
candidates: fused surrogate value + min distance + min/max statistics, then num_pts = 4 merit / argmin /
distance-update rounds.  Workload: d = 50, n = 20 000 centers (Cubic kernel, linear tail, coefficients from a real
fit of those centers), 2^20 candidates PER GPU (weak scaling; rows sharded over the ranks, centers replicated,
NCCL moves only statistics, (merit, index) pairs and the winning row).

metric   candidate x center pairs / s, whole job
value    device-resident candidates (inputs in HBM before the timed region), CUDA events, max over ranks
e2e      the public API call with HOST (pinned) candidate arrays: H2D of the batch and D2H of the picks inside;
         `e2e_pageable` is the same call with a plain (pageable) numpy array, what a pySOT caller hands over
roofline the fused score kernel against the FP64 FMA pipe, measured live (the path is FP64-compute-bound, see
         DESIGN.md): achieved = pairs * (3d + 6) flops / kernel time; the HBM figure is reported beside it;
         `rooflines` adds predict and predict_deriv (config #3 shape, device tensors, CUDA events)
parity   (N = 1) the GPU path against the oracle on the CPU baseline's own sample of the workload, outside the timed region
sweep    config #4's candidate counts: 2^20 .. 2^26 candidates in total (split over the ranks), and one strong-scaling
         point at 2^24 in total; candidates drawn on the device by b200rbf_generate
fit      BASELINE's second headline number: RBF fit seconds at N = 20 000 (TPS, d = 10), rank 0 only, beside the
         reference's CPU fit (oracle port, all host cores) at N = 1 000 / 5 000 / 20 000 (`cpu_baseline_fit`)

--impl reference times the CPU restatement of the reference (oracle/, same scipy/numpy routines as pySOT) on a
bounded sample of the same workload with every host core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, N_CENTERS, M_PER_GPU, NUM_PTS = 50, 20000, 1 << 20, 4
WEIGHTS = [0.3, 0.5, 0.8, 0.95]
DTOL = 1e-3
FLOPS_PER_PAIR = 3 * D + 6  # sub + fma per coordinate, sqrt, phi (2), weight fma (2), running min
WORKLOAD = "config#4 candidate acquisition: d=50, n=20000 centers, cubic+linear tail, 2^20 DYCORS candidates per GPU, num_pts=4"


def config_dict(world):
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "global_candidates": M_PER_GPU * world, "centers": N_CENTERS, "dim": D,
            "num_pts": NUM_PTS, "parallelism": "candidate rows sharded x%d, centers replicated" % world,
            "l2": "inputs larger than L2 (419 MB of candidates per GPU per step)"}


def make_centers():
    rng = np.random.default_rng(0)
    X = rng.random((N_CENTERS, D))
    fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
    Xpend = np.random.default_rng(1).random((7, D))
    return X, fX, Xpend


def dycors_candidates_host(m, xbest, seed):
    """DYCORS-shaped candidates (candidate_dycors.py:73-86 in distribution): prob_perturb 0.4, sigma 0.2, truncated
    normal in [0, 1], rows with no perturbed coordinate get one in [0, d-2]."""
    import scipy.stats as st

    rng = np.random.default_rng(seed)
    mask = rng.random((m, D)) < 0.4
    none = np.where(mask.sum(axis=1) == 0)[0]
    mask[none, rng.integers(0, D - 1, size=len(none))] = True
    a, b = (0.0 - xbest) / 0.2, (1.0 - xbest) / 0.2
    u = rng.random((m, D))
    lo, hi = st.norm.cdf(a), st.norm.cdf(b)
    z = st.norm.ppf(lo + u * (hi - lo))
    return np.where(mask, np.clip(xbest + 0.2 * z, 0.0, 1.0), xbest)


def dycors_candidates_device(m, xbest, seed, device):
    """Same distribution generated on the device with torch's Philox generator (tensor hand-off, not product code)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    xb = torch.from_numpy(xbest).to(device)
    mask = torch.rand((m, D), generator=g, device=device, dtype=torch.float64) < 0.4
    none = (mask.sum(dim=1) == 0).nonzero().ravel()
    if none.numel() > 0:
        mask[none, torch.randint(0, D - 1, (none.numel(),), generator=g, device=device)] = True
    lo = torch.special.ndtr((0.0 - xb) / 0.2)
    hi = torch.special.ndtr((1.0 - xb) / 0.2)
    u = torch.rand((m, D), generator=g, device=device, dtype=torch.float64)
    z = torch.special.ndtri(lo + u * (hi - lo))
    return torch.where(mask, torch.clamp(xb + 0.2 * z, 0.0, 1.0), xb).contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = str(gpu_index), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9 and parts[0] == self.idx:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[5 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_scoring_baseline(X, fX, Xpend, coef, xbest, budget_s=15.0, threads=None, keep=False):
    """The oracle port (same scipy cdist / numpy ufunc / dot calls as pySOT) on a bounded sample, all host threads."""
    from oracle import rbf_oracle as orc

    threads = threads or max(1, os.cpu_count() or 1)
    o = orc.OracleRBF(D, np.zeros(D), np.ones(D))
    o.add_points(X, fX)
    o.c, o.updated = coef.reshape(-1, 1), True
    chunk = 1024
    probe = dycors_candidates_host(chunk * min(threads, 8), xbest, 11)
    t0 = time.perf_counter()
    orc.score_tiled(o, X, probe, Xpend, chunk=chunk, threads=threads)
    rate = probe.shape[0] / (time.perf_counter() - t0)
    m = int(min(max(rate * budget_s, chunk * threads), 1 << 19))
    m = max(chunk, (m // chunk) * chunk)
    cand = dycors_candidates_host(m, xbest, 12)
    t0 = time.perf_counter()
    fv, dm = orc.score_tiled(o, X, cand, Xpend, chunk=chunk, threads=threads)
    picked = orc.select_from_scores(cand, fv, dm, WEIGHTS, NUM_PTS, DTOL)
    dt = time.perf_counter() - t0
    out = {"value": m * N_CENTERS / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
           "sample": "%d candidates x %d centers, d=%d, tiled reference path (%d-row chunks on a %d-thread pool), "
                     "%.1f s" % (m, N_CENTERS, D, chunk, threads, dt), "seconds": dt, "m": m}
    if keep:
        out["_sample"] = (cand, fv.ravel(), dm.ravel(), picked)
    return out


def cpu_fit_baseline(sizes=(1000, 5000, 20000)):
    """The reference's _fit (rbf.py:110-130, 164-168: cdist, lu_factor, two solve_triangular) through the oracle port on
    BASELINE config #3's inputs, all host cores (OpenBLAS threads), one cold run per size."""
    from oracle import rbf_oracle as orc

    out = []
    for n in sizes:
        need_gb = 5.0 * 8.0 * (n + 11) ** 2 / 1e9  # dists, Phi, A, LU copy, L / U copies of the reference
        try:
            import psutil

            if psutil.virtual_memory().available / 1e9 < need_gb + 4.0:
                out.append({"n": n, "skipped": "needs ~%.0f GB of host memory" % need_gb})
                continue
        except Exception:
            pass
        rng = np.random.default_rng(0)
        X = rng.random((n, 10))
        fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
        o = orc.OracleRBF(10, np.zeros(10), np.ones(10), "tps", "linear")
        o.add_points(X, fX)
        t0 = time.perf_counter()
        o.fit()
        out.append({"n": n, "d": 10, "kernel": "tps", "seconds": time.perf_counter() - t0, "cores": os.cpu_count(),
                    "kind": "port"})
        del o
    return out


def run_reference(args, rank):
    """--impl reference: the CPU restatement of the reference, rank 0 only."""
    if rank != 0:
        return
    X, fX, Xpend = make_centers()
    coef = np.random.default_rng(2).standard_normal(N_CENTERS + D + 1)
    xbest = X[np.argmin(fX)]
    per_step = max(3.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_scoring_baseline(X, fX, Xpend, coef, xbest, budget_s=per_step)
        if s >= args.warmup:
            vals.append(r)
    tot_pairs = sum(r["m"] * N_CENTERS for r in vals)
    tot_s = sum(r["seconds"] for r in vals)
    v = tot_pairs / tot_s
    cb = dict(vals[-1])
    cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "candidate x center pairs/s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(max(1, args.gpus)),
        "note": "CPU reference path (oracle port of pySOT's numpy / scipy calls, tiled over all host cores); each step is a "
                "bounded sample of the workload",
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ native arm
def fit_benchmark(pb, peaks, sizes=(1000, 5000, 20000)):
    """BASELINE config #3: TPS + linear tail, d = 10, X ~ U[0,1]^10 seed 0, fX = sin(3 sum X) + X0^2.  Returns the fit
    records and the roofline entries of predict / predict_deriv (device tensors in and out, CUDA events)."""
    import torch

    out, roofs = [], []
    for n in sizes:
        rng = np.random.default_rng(0)
        X = rng.random((n, 10))
        fX = (np.sin(3.0 * X.sum(axis=1)) + X[:, 0] ** 2)[:, None]
        s = pb.GpuRBFInterpolant(dim=10, lb=np.zeros(10), ub=np.ones(10), kernel=pb.TPSKernel(), tail=pb.LinearTail(10))
        s.add_points(X, fX)
        best_dev, best_wall = 1e30, 1e30
        for _ in range(2 if n >= 20000 else 3):
            s.updated = False
            t0 = time.perf_counter()
            s._fit()
            best_wall = min(best_wall, time.perf_counter() - t0)
            best_dev = min(best_dev, s.fit_seconds)
        N = n + 11
        # algorithmic work of the REFERENCE's algorithm (LU of the saddle-point system).  The null-space
        # Cholesky executes n^3 / 3 + the rank-3t update and W = Phi Q (2 n^2 * 5 tp) + assembly instead.
        flops = 2.0 / 3.0 * N ** 3 + 2.0 * N ** 2 + 0.5 * n * n * (3 * 10 + 4)
        executed = n ** 3 / 3.0 + 2.0 * n * n * 5 * 16 + 0.5 * n * n * (3 * 10 + 4) + 4.0 * n * n
        resid = float(np.max(np.abs(s.predict(X[:2000]) + s.eta * s.c[11:11 + min(n, 2000)] - fX[:2000])))
        rec = {"n": n, "d": 10, "kernel": "tps", "solver": s.handle.last_fit_method, "seconds_device": best_dev,
               "seconds_host_to_host": best_wall, "tflops": flops / best_dev / 1e12,
               "frac_of_dmma_peak": flops / best_dev / 1e12 / peaks["dmma_tflops"],
               "tflops_executed": executed / best_dev / 1e12,
               "frac_executed": executed / best_dev / 1e12 / peaks["dmma_tflops"], "interp_residual": resid}
        if n == sizes[-1]:
            # config #3's other half: predict on 1e5 points (seed 1), and predict_deriv, host arrays in and out ...
            q = np.random.default_rng(1).random((100000, 10))
            s.predict(q)  # untimed: sizes the staging buffers
            t0 = time.perf_counter()
            s.predict(q)
            rec["predict_pairs_per_s"] = q.shape[0] * n / (time.perf_counter() - t0)
            s.predict_deriv(q)
            t0 = time.perf_counter()
            s.predict_deriv(q)
            rec["predict_deriv_pairs_per_s"] = q.shape[0] * n / (time.perf_counter() - t0)
            # ... and device-resident: the kernels alone against the FP64 datapath (3d + 6 / 5d + 6 flops)
            dev = torch.device("cuda", s._device)
            m2 = 1 << 20
            qd = torch.rand((m2, 10), dtype=torch.float64, device=dev)
            od = torch.empty(m2, dtype=torch.float64, device=dev)
            gd = torch.empty((m2, 10), dtype=torch.float64, device=dev)
            stream = torch.cuda.current_stream(dev)
            s.handle.set_stream(stream.cuda_stream)
            try:
                for name, fn, fl in (("predict", lambda: s.handle.predict(qd, s.lb, s.ub, out=od, m=m2), 3 * 10 + 6),
                                     ("predict_deriv", lambda: s.handle.predict_deriv(qd, s.lb, s.ub, out=gd, m=m2), 5 * 10 + 6)):
                    fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(3):
                        fn()
                    e1.record(stream)
                    torch.cuda.synchronize(dev)
                    ms = e0.elapsed_time(e1) / 3
                    tf = float(m2) * n * fl / (ms * 1e-3) / 1e12
                    roofs.append({"kernel": "%s (TPS, d=10, n=%d, m=2^20 device-resident)" % (name, n), "bound": "tensor",
                                  "achieved": tf, "peak": peaks["dfma_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["dfma_tflops"],
                                  "flops_per_pair": fl, "launch_ms": ms, "pairs_per_s": float(m2) * n / (ms * 1e-3)})
            finally:
                s.handle.set_stream(None)
            del qd, od, gd
        out.append(rec)
        del s
    return out, roofs


def run_native(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import pysot_b200 as pb
    from pysot_b200.sharded import GpuShardBackend, sharded_select, sharded_weighted_distance_merit

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    X, fX, Xpend = make_centers()
    lb, ub = np.zeros(D), np.ones(D)
    xbest = X[np.argmin(fX)]
    rbf = pb.GpuRBFInterpolant(dim=D, lb=lb, ub=ub, device=local_rank)  # Cubic + LinearTail (reference default)
    rbf.add_points(X, fX)
    t0 = time.perf_counter()
    fit50 = None
    if rank == 0:  # ONE fit (it does not shard: "replicas only"); the replicas receive the coefficients
        rbf._fit()
        fit50 = {"n": N_CENTERS, "d": D, "kernel": "cubic", "solver": rbf.handle.last_fit_method,
                 "seconds_device": rbf.fit_seconds,
                 "seconds_host_to_host": time.perf_counter() - t0,
                 "note": "cold: the first fit of the process, it allocates the 3.4 GB workspace inside the timed region; "
                         "the warm figures are in `fit`"}
    if world > 1:
        cvec = torch.from_numpy(rbf.c.ravel().copy() if rank == 0 else np.empty(N_CENTERS + D + 1)).to(device)
        dist.broadcast(cvec, src=0)
        if rank != 0:
            rbf.set_coefficients(cvec.cpu().numpy())  # b200rbf_load
        del cvec
    h = rbf.handle
    Xdist = np.vstack((X, Xpend))

    cand_dev = dycors_candidates_device(M_PER_GPU, xbest, 3 + rank, device)
    cand_pin = torch.empty((M_PER_GPU, D), dtype=torch.float64, pin_memory=True)
    cand_pin.copy_(cand_dev)
    cand_host = cand_pin.numpy()
    torch.cuda.synchronize()
    backend = GpuShardBackend(rbf)

    def step_device():
        return sharded_select(backend, cand_dev, Xdist, True, WEIGHTS, NUM_PTS, DTOL, offset=rank * M_PER_GPU)

    def step_e2e():
        return sharded_weighted_distance_merit(NUM_PTS, rbf, X, fX, cand_host, WEIGHTS, Xpend, DTOL, offset=rank * M_PER_GPU)

    # ---- value: device-resident inputs, CUDA events on the stream the kernels run on (torch's current stream)
    for _ in range(args.warmup):
        picks = step_device()
    barrier()
    sampler = ClockSampler(local_rank if os.environ.get("CUDA_VISIBLE_DEVICES") is None else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
    if rank == 0:
        sampler.start()
    l0 = h.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record()
    for _ in range(args.steps):
        picks = step_device()
        kernel_ms.append(h.last_score_ms)
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sum_over_ranks(h.launches - l0)
    clocks = sampler.stop() if rank == 0 else None
    pairs_per_step = float(M_PER_GPU) * world * N_CENTERS
    value = pairs_per_step * args.steps / (dev_ms * 1e-3)

    # ---- e2e: public API, host (pinned) candidates, H2D + D2H inside, wall clock between device syncs
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pts, idx = step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = pairs_per_step * args.steps / e2e_s
    h2d = M_PER_GPU * D * 8 + 7 * D * 8 + 2 * D * 8
    d2h = NUM_PTS * (16 + D * 8) + 4 * 8

    # ---- the same call with a PAGEABLE numpy array (what a pySOT caller hands over): staged through the handle's
    # pinned double buffer chunk by chunk
    cand_pageable = np.array(cand_host, copy=True)
    sharded_weighted_distance_merit(NUM_PTS, rbf, X, fX, cand_pageable, WEIGHTS, Xpend, DTOL, offset=rank * M_PER_GPU)
    barrier()
    t0 = time.perf_counter()
    pg_steps = max(1, min(args.steps, 3))
    for _ in range(pg_steps):
        pts_pg, idx_pg = sharded_weighted_distance_merit(NUM_PTS, rbf, X, fX, cand_pageable, WEIGHTS, Xpend, DTOL, offset=rank * M_PER_GPU)
    barrier()
    pg_s = max_over_ranks(time.perf_counter() - t0)
    e2e_pageable = {"value": pairs_per_step * pg_steps / pg_s, "unit": "pairs/s", "ms_per_step": 1e3 * pg_s / pg_steps,
                    "steps": pg_steps, "same_picks_as_pinned": bool(np.array_equal(idx_pg, idx))}
    del cand_pageable

    # ---- config #4's candidate counts (total over the ranks), candidates drawn on the device by b200rbf_generate
    import types

    _Prob = types.SimpleNamespace(dim=D, lb=lb, ub=ub, int_var=np.array([], dtype=int))

    def timed_device_steps(cand_t, steps):
        off = rank * int(cand_t.shape[0])
        sharded_select(backend, cand_t, Xdist, True, WEIGHTS, NUM_PTS, DTOL, offset=off)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            sharded_select(backend, cand_t, Xdist, True, WEIGHTS, NUM_PTS, DTOL, offset=off)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / steps

    sweep, strong = [], None
    if not args.skip_extras:
        del cand_dev
        torch.cuda.empty_cache()
        for log2m in (20, 22, 24, 26):
            m_tot = 1 << log2m
            m_loc = m_tot // world
            c_t = pb.generate_on_device("dycors", _Prob, rbf, X, fX, prob_perturb=0.4, sampling_radius=0.2, num_cand=m_loc,
                                        seed=1000 + 17 * log2m + rank)
            ms = timed_device_steps(c_t, 2 if log2m <= 24 else 1)
            rec = {"global_candidates": m_tot, "per_gpu": m_loc, "ms_per_step": ms,
                   "value": float(m_tot) * N_CENTERS / (ms * 1e-3), "unit": "pairs/s"}
            sweep.append(rec)
            if log2m == 24:
                strong = {"global_candidates": m_tot, "n_gpus": world, "ms_per_step": ms, "value": rec["value"],
                          "unit": "pairs/s", "scaling": "strong",
                          "note": "fixed global candidate count split over the ranks; compare across --gpus runs"}
            del c_t
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0), peaks measured live on this device
    peaks = h.fp64_peaks(ms=300.0)
    mp, mp_kind = measured_peaks()
    k_ms = float(np.mean(kernel_ms))
    k_tflops = float(M_PER_GPU) * N_CENTERS * FLOPS_PER_PAIR / (k_ms * 1e-3) / 1e12
    alg_bytes = M_PER_GPU * (8 * D + 16)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "score_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch_" + h.last_score_kernel, tj.get("dram_bytes_per_launch"))
        except Exception:
            traffic = None
    which = h.last_score_kernel
    if which == "int8":
        kname = ("score_i8_kernel<NK=2, cubic> (norm-expansion distances; dot products on the INT8 tensor cores: tcgen05.mma.kind::i8, "
                 "six base-256 digit planes of the 46-bit fixed-point coordinates, exact int32 accumulation in TMEM; near pairs recomputed)")
        knote = ("achieved counts SURVEY 8(d)'s ALGORITHMIC 3d+6 flops per pair (direct differences). The kernel executes 52 int8 MMAs "
                 "(M128 N64 K32) per 128 x 64 pairs on the tensor cores + ~12 FP64 and ~20 integer operations per pair; on sm_100a the "
                 "tcgen05 MMAs and FP64 instructions time-share (measured additive, tools/umma_overlap.cu), so the launch time is "
                 "MMA + drain + FP64 epilogue; a fraction above 1 is arithmetic moved off the FP64 datapath, not a faster pipe")
    else:
        kname = "score_mma_kernel<K4=13, cubic> (norm-expansion distances, DMMA dot products, near pairs recomputed)"
        knote = ("achieved counts SURVEY 8(d)'s ALGORITHMIC 3d+6 flops per pair (direct differences); the kernel executes "
                 "2*52 (DMMA) + 7 on the shared FP64 FMA/DMMA datapath, so a fraction above 1 is arithmetic savings, not a faster pipe "
                 "(ncu: that datapath is 87.5 % busy)")
    roofline = {
        "kernel": kname, "kernel_id": which,
        "bound": "tensor", "bound_detail": "FP64: DMMA (fp64 tensor op) and DFMA share one datapath on sm_100a (measured: "
                                           "36.5 vs 36.6 TFLOP/s, 36.0 interleaved); the bf16 tensor peak does not apply; HBM traffic is < 0.1 % of its peak",
        "achieved": k_tflops,
        "peak": peaks["dfma_tflops"], "unit": "TFLOP/s", "frac": k_tflops / peaks["dfma_tflops"],
        "peak_source": "b200rbf_fp64_peaks: dependency-free DFMA loop on this device, 300 ms (MEASURED_PEAKS.json has no fp64 entry)",
        "flops_per_pair": FLOPS_PER_PAIR,
        "note": knote,
        "launch_ms": k_ms, "share_of_step": k_ms * args.steps / dev_ms,
        "dmma_peak_tflops": peaks["dmma_tflops"], "traffic": traffic,
        "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": mp["hbm_gbs"],
                "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / mp["hbm_gbs"], "peak_kind": mp_kind},
    }
    # the same launch on EXECUTED operations (what ncu's pipe counters see), beside the algorithmic figure above
    pairs_per_s = float(M_PER_GPU) * N_CENTERS / (k_ms * 1e-3)
    if which == "int8":
        kpad = 32 * ((D + 31) // 32)  # int8 k-steps of 32 bytes
        i8_ops = 26 * kpad * 2        # 26 digit-plane products per pair, multiply + add
        roofline["executed"] = {
            "int8_ops_per_pair": i8_ops, "int8_tops": pairs_per_s * i8_ops / 1e12, "int8_nominal_peak_tops": 4500.0,
            "int8_frac_of_nominal": pairs_per_s * i8_ops / 1e12 / 4500.0,
            "fp64_ops_per_pair": 12, "fp64_tflops": pairs_per_s * 12 * 2 / 1e12,
            "fp64_frac": pairs_per_s * 12 * 2 / 1e12 / peaks["dfma_tflops"],
            "note": "nominal dense int8 peak (no measured int8 figure in MEASURED_PEAKS.json); M128 N64 K32 MMAs complete at "
                    "27.6 ns each on one SM (tools/umma_rate2.cu) = 9.5 TMAC/s = 19 TOPS per SM, 62 % of the nominal peak: the "
                    "per-instruction floor for N <= 64 (N = 128 would need 896 of the 512 TMEM columns for the seven "
                    "accumulators), and the tensor phase is 42 % of the tile time; FP64 operations counted as FMAs (2 flops); the two "
                    "phases and the TMEM drain run one after the other (ncu: tensor-core pipe 44 %, ALU 32 %, XU 16 %, "
                    "FP64 14-24 % busy, summing to ~100 %)"}
    else:
        ex = 2 * 4 * ((D + 3) // 4) + 7
        roofline["executed"] = {"fp64_ops_per_pair": ex, "fp64_tflops": pairs_per_s * ex / 1e12,
                                "fp64_frac": pairs_per_s * ex / 1e12 / peaks["dfma_tflops"],
                                "note": "2 x 4 ceil(d / 4) DMMA flops + 7 FP64 operations per pair on the shared datapath"}
    line = {
        "metric": "candidate x center pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(world),
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps, "host_memory": "pinned"},
        "e2e_pageable": e2e_pageable, "sweep": sweep, "strong_scaling": strong,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "picked": [int(i) for i in picks[0]],
        "fit_d50": fit50,
    }
    if world == 1 and not args.skip_extras:
        coef = rbf.c.ravel().copy()
        cb = cpu_scoring_baseline(X, fX, Xpend, coef, xbest, keep=True)
        line["cpu_baseline"] = {k: v for k, v in cb.items() if k in ("value", "unit", "cores", "kind", "sample")}
        # parity on the CPU baseline's own sample (same coefficients: the GPU's real fit), default DMMA kernel
        c_s, fv_o, dm_o, pick_o = cb["_sample"]
        idx_g, _, fv_g, dm_g, _ = pb.merit_select_indices(NUM_PTS, rbf, X, c_s, WEIGHTS, Xpend, DTOL, want_scores=True)
        line["parity"] = {
            "sample": "%d candidates x %d centers (the cpu_baseline sample), coefficients from the GPU fit" % (len(fv_o), N_CENTERS),
            "fvals_relmax": float(np.max(np.abs(fv_g - fv_o)) / np.max(np.abs(fv_o))),
            "dmerit_relmax": float(np.max(np.abs(dm_g - dm_o)) / np.max(np.abs(dm_o))),
            "picks_gpu": [int(i) for i in idx_g], "picks_oracle": [int(i) for i in pick_o],
            "picks_identical": bool(np.array_equal(idx_g, pick_o)),
            "bars": "fvals 1e-10, dmerit 1e-12, picks identical (north_star); golden-vector versions of this check at "
                    "2^17 candidates and for the N = 5 000 / 20 000 fits: tests/test_gpu_large_shapes.py"}
        del cb
        fits, roofs = fit_benchmark(pb, peaks)
        line["fit"] = fits
        line["rooflines"] = [dict(roofline, name="score (headline)")] + roofs
        line["cpu_baseline_fit"] = cpu_fit_baseline()
        for f, c in zip(line["fit"], line["cpu_baseline_fit"]):
            if "seconds" in c:
                f["cpu_seconds"] = c["seconds"]
                f["speedup_vs_cpu"] = c["seconds"] / f["seconds_host_to_host"]
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--skip-extras", action="store_true", help="skip the CPU baseline and the fit sweep")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch ourselves as one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--skip-extras"] if args.skip_extras else [])
        sys.exit(subprocess.call(cmd))
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
