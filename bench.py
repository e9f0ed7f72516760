#!/usr/bin/env python
"""bench.py -- stars/sec of the full-grid brute-force likelihood sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2]

A step = one pass of the hot path (loglike + label priors + lnpost's first threshold, i.e.
``bf_sweep_batch``) over one synthetic catalogue of ``nstar`` stars against the whole grid.
N=1 workload: BASELINE.json configs[1] (1k stars, 8 bands, 1M-point grid).  For N>1 (torchrun, one
rank per GPU) every rank sweeps its own ``nstar`` stars against a replica of the grid (rank 0
builds it, one NCCL broadcast, no collective in the hot loop): weak scaling.

  value            stars/s from CUDA-event device time (first kernel -> last kernel of each call,
                   grid and scratch resident in HBM), max over ranks
  e2e              stars/s through the call a user makes, BruteForce.fit's per-object body
                   (bf_fit_batch: the same sweep, then lnpost with the default Galactic prior, evidence
                   and resampling on the device; host float64 photometry in, Ndraws posterior samples
                   per star out): wall clock of the calls, host preparation, H2D and D2H included
  e2e_records      the same for bf_sweep_batch, which ships every selected model's record to the host
                   for a host-side lnpost (user-supplied prior callables): PCIe-bound
  roofline         dominant kernel k_magfit: algorithmic bytes (Nmodel x Nfilt x 12 B per star per
                   pass) / its CUDA-event time, against the measured HBM copy peak
  cpu_baseline     the C oracle (port of the reference's loglike) on this box's host cores
--impl reference   times only that CPU path (the reference is Python+numba and cannot travel to
                   the GPU box; oracle/loglike_ref.c is its pinned restatement).
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

METRIC = "stars/sec (full-grid loglike sweep)"
GRID_KIND = "locus"   # mock.make_grid kind: "locus" (MIST-like stellar locus) or "tilt" (degenerate stress grid)
CONFIG_NAMES = {
    1: "C1: 1 star x 10k models x 5 bands",
    2: "C2: 1k stars x 1M models x 8 bands",
    3: "C3: 3M models x 12 bands, avlim (0,6), 10% band drop-outs",
    5: "C5: 3M models x 8 bands",
}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def read_traffic():
    """DRAM bytes per k_magfit launch from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "magfit_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(cfg_id, nstar, rank):
    from brutus_b200 import mock
    cfg = dict(mock.CONFIGS[cfg_id])
    if nstar:
        cfg["nstar"] = nstar
    return cfg


def make_inputs(cfg_id, cfg, rank, need_grid=True):
    from brutus_b200 import mock
    grid = labels = None
    if need_grid:
        grid, labels = mock.make_grid(cfg["nmodel"], cfg["nfilt"], seed=1000 + cfg_id, kind=GRID_KIND)
    return grid, labels


def model_priors(labels):
    """Static inputs of lnpost for the mock grid: IMF prior over 'mini' (fit()'s default `lnprior`,
    brutus/fitting.py:1296-1300) and the 'feh' / 'loga' labels of the Galactic prior."""
    from brutus_b200 import fitting
    names = labels.dtype.names
    lnprior = fitting.imf_lnprior(labels["mini"]) if "mini" in names else None
    return dict(lnprior=lnprior, feh=labels["feh"] if "feh" in names else None,
                loga=labels["loga"] if "loga" in names else None)


def make_stars(cfg_id, cfg, grid, rank):
    from brutus_b200 import mock
    return mock.make_stars(grid, cfg["nstar"], seed=2000 + cfg_id + 100 * rank, av_max=cfg["av_max"],
                           dropout=cfg["dropout"])


def host_cores():
    """Cores this process may run on.  torchrun exports OMP_NUM_THREADS=1, which would make the CPU
    arm single-threaded: the thread count is therefore passed to the oracle explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(cfg, grid, stars, nthreads, nsample):
    """Times the oracle (port of the reference loglike) on `nsample` stars with `nthreads` threads."""
    from oracle import oracle
    oracle.build()
    nt = nthreads or host_cores()
    sl = slice(0, nsample)
    t0 = time.perf_counter()
    oracle.loglike_batch(stars["flux"][sl], stars["err"][sl], stars["mask"][sl], grid,
                         parallax=stars["parallax"][sl], parallax_err=stars["parallax_err"][sl],
                         nthreads=nt, avlim=cfg["avlim"])
    dt = time.perf_counter() - t0
    return nsample / dt, nt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = workload(args.config, args.nstar, 0)
    grid, _ = make_inputs(args.config, cfg, 0)
    from oracle import oracle
    oracle.build()
    nt = host_cores()
    # a bounded sample per step: ~3 s of work on all cores at C2 (the whole run stays within a minute or two)
    per_step = max(4 * nt, 8) if cfg["nmodel"] <= 1_000_000 else max(nt, 4)
    cfg_s = dict(cfg, nstar=per_step * (args.steps + args.warmup))
    stars = make_stars(args.config, cfg_s, grid, 0)
    times = []
    for it in range(args.steps + args.warmup):
        sl = slice(it * per_step, (it + 1) * per_step)
        sub = {k: stars[k][sl] for k in ("flux", "err", "mask", "parallax", "parallax_err")}
        rate, _, dt = cpu_sample(cfg, grid, sub, nt, per_step)
        if it >= args.warmup:
            times.append(dt)
    tot = float(np.sum(times))
    value = per_step * args.steps / tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "stars/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": CONFIG_NAMES[args.config], "nmodel": cfg["nmodel"],
                       "nfilt": cfg["nfilt"], "grid": "mock %s (brutus_b200/mock.py)" % GRID_KIND,
                       "stars_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "stars/s", "cores": nt, "kind": "port",
                             "sample": "%d stars per step x %d steps, OpenMP over stars, C port "
                                       "(oracle/loglike_ref.c) of the reference's numba loglike"
                                       % (per_step, args.steps)},
            "e2e": {"value": value, "unit": "stars/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    cfg = workload(args.config, args.nstar, rank)
    from brutus_b200 import _lib
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            # keep stdout to the one JSON line (an image-level nccl.conf may ask for the version banner;
            # the environment takes precedence over it)
            os.environ["NCCL_DEBUG"] = "WARN"
        # NCCL prints its version banner on stdout at the first communicator: keep fd 1 for the JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    h = _lib.Handle(local_rank, args.precision)
    # ---- stage the grid: rank 0 builds it; one NCCL broadcast; device-side re-tiling ----
    t0 = time.perf_counter()
    if world > 1:
        import torch
        from brutus_b200 import shard
        shape = (cfg["nmodel"], cfg["nfilt"], 3)
        grid, labels = make_inputs(args.config, cfg, 0) if rank == 0 else (None, None)
        # one NCCL broadcast GPU -> GPU, re-tiled on each device (bf_set_grid_device)
        dgrid = shard.broadcast_grid(grid, shape, dist=dist, src=0, handle=h,
                                     device=torch.device("cuda", local_rank))
        if grid is None:
            grid = dgrid.cpu().numpy()  # only to draw this rank's synthetic stars from
        del dgrid
        torch.cuda.empty_cache()
        # the per-model priors / labels of lnpost (3 x Nmodel float64), also one broadcast
        pri = model_priors(labels) if rank == 0 else None
        obj = [pri]
        dist.broadcast_object_list(obj, src=0)
        pri = obj[0]
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    else:
        grid, labels = make_inputs(args.config, cfg, 0)
        h.set_grid(grid)
        pri = model_priors(labels)
    h.set_model_priors(**pri)
    t_stage = time.perf_counter() - t0
    stars = make_stars(args.config, cfg, grid, rank)
    opts = _lib.make_options(avlim=cfg["avlim"])
    nstar = cfg["nstar"]

    def barrier():
        if world > 1:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    opts_dev = _lib.make_options(avlim=cfg["avlim"], skip_d2h=True)

    def step(e2e):
        """One pass over the catalogue.  e2e=False: records stay in HBM (device-timed `value`);
        e2e=True: host buffers in, pinned host records out (wall-clocked)."""
        h.flush_l2()
        t = time.perf_counter()
        res = h.sweep_batch(stars["flux"], stars["err"], stars["mask"], stars["parallax"],
                            stars["parallax_err"], opts=opts if e2e else opts_dev, rows=args.rows)
        wall = time.perf_counter() - t
        return res, wall, h.stats()

    def step_fit():
        """One pass of BruteForce.fit's per-object body (bf_fit_batch): host photometry in, posterior
        samples out; the generator is keyed by (seed, catalogue index)."""
        h.flush_l2()
        t = time.perf_counter()
        res = h.fit_batch(stars["flux"], stars["err"], stars["mask"], stars["parallax"],
                          stars["parallax_err"], coords=stars["coords"], opts=opts,
                          nmc_prior=args.nmc_prior, ndraws=args.ndraws, seed=12345, star_base=rank * nstar,
                          mem_lim=8000.,   # fit()'s default mem_lim (brutus/fitting.py:1436)
                          copy=False)      # the draws stay in the library's pinned arena (like the records)
        wall = time.perf_counter() - t
        return res, wall, h.stats()

    for _ in range(args.warmup):
        step(True)
        step(False)
        step_fit()
    tracing = bool(int(os.environ.get("BRUTUS_B200_TRACE", "0") or 0))
    if tracing:
        h.trace()   # drop the warm-up's entries
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region 1: K steps, device time from CUDA events (value, roofline) ----
    barrier()
    t_region = time.perf_counter()
    dev_ms = 0.0
    agg = {}
    for _ in range(args.steps):
        res, wall, st = step(False)
        dev_ms += st["ms_device"]
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
    barrier()
    if tracing:   # per-kernel CUDA-event times of the device-timed steps (stderr; the JSON line stays alone on stdout)
        for name, (cnt, ms) in sorted(h.trace().items(), key=lambda kv: -kv[1][1]):
            sys.stderr.write("trace value-steps  %-22s %6d launches %10.3f ms/step\n" % (name, cnt, ms / args.steps))
    # ---- timed region 2: K steps end to end through the C ABI with host buffers ----
    wall_s = 0.0
    agg_e = {}
    for _ in range(args.steps):
        res, wall, st = step(True)
        wall_s += wall
        for k, v in st.items():
            agg_e[k] = agg_e.get(k, 0) + v
    barrier()
    # ---- timed region 3: K steps of the fit-level call (device posterior), end to end ----
    wall_f = 0.0
    agg_f = {}
    for _ in range(args.steps):
        resf, wall, st = step_fit()
        wall_f += wall
        for k, v in st.items():
            agg_f[k] = agg_f.get(k, 0) + v
    barrier()
    if tracing:
        h.trace()
    t_region = time.perf_counter() - t_region
    clocks = sampler.summary()
    if world > 1:
        import torch
        t = torch.tensor([dev_ms, wall_s, wall_f], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_s, wall_f = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    total_stars = nstar * args.steps * world
    value = total_stars / (dev_ms * 1e-3)
    e2e = total_stars / wall_s
    e2e_fit = total_stars / wall_f
    peak, peak_src = read_peaks()
    bytes_per_star = cfg["nmodel"] * cfg["nfilt"] * 12
    launches = max(1, agg["magfit_launches"])
    ach = (bytes_per_star * agg["magfit_star_passes"] / launches) / (agg["ms_magfit"] / launches * 1e-3) / 1e9
    traffic = read_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "stars/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": CONFIG_NAMES[args.config], "nmodel": cfg["nmodel"], "nfilt": cfg["nfilt"],
                   "grid": "mock %s (brutus_b200/mock.py)" % GRID_KIND, "stars_per_step_per_gpu": nstar, "parallelism": "stars sharded x%d, grid replicated" % world,
                   "l2": "512 MB L2 flush before every step",
                   "grid_stage_s": round(t_stage, 3)},
        "e2e": {"value": e2e_fit, "unit": "stars/s", "ms_per_step": 1e3 * wall_f / args.steps,
                "h2d_bytes_per_step": int(agg_f["h2d_bytes"] / args.steps),
                "d2h_bytes_per_step": int(agg_f["d2h_bytes"] / args.steps),
                "call": "bf_fit_batch (the per-object body of BruteForce.fit)",
                "nmc_prior": args.nmc_prior, "ndraws": args.ndraws,
                "gpu_launches": int(agg_f["kernel_launches"]),
                "device_ms_per_step": agg_f["ms_device"] / args.steps,
                "posterior_ms_per_step": agg_f["ms_post"] / args.steps,
                "selected2_per_step": agg_f["selected2"] / args.steps,
                "clipped_stars_per_step": agg_f["clipped"] / args.steps,
                "finite_evidence_frac": float(np.mean(resf["levid"] > -1e299)),
                "note": "host float64 photometry in, Ndraws posterior samples per star out: the full-grid "
                        "sweep of `value`, then lnpost (default Galactic prior, Nmc_prior Monte Carlo draws per "
                        "selected model), evidence and resampling on the device -- a superset of the work the "
                        "reference arm (loglike only) is timed on"},
        "e2e_records": {"value": e2e, "unit": "stars/s", "ms_per_step": 1e3 * wall_s / args.steps,
                        "h2d_bytes_per_step": int(agg_e["h2d_bytes"] / args.steps),
                        "d2h_bytes_per_step": int(agg_e["d2h_bytes"] / args.steps),
                        "call": "bf_sweep_batch", "record_rows": args.rows,
                        "gpu_launches": int(agg_e["kernel_launches"]),
                        "note": "host float64 photometry in, every selected model's record (48 B) out to pinned "
                                "host memory for a host-side lnpost (user prior callables); the D2H of one star "
                                "batch overlaps the kernels of the next; PCIe-bound (~52 GB/s)"},
        "gpu_launches": int(agg["kernel_launches"]),
        "roofline": {"bound": "hbm", "kernel": "k_magfit", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "peak_source": peak_src,
                     "traffic": None if traffic is None else
                     traffic.get("dram_bytes_per_star_pass", 0) * agg["magfit_star_passes"] / launches,
                     "traffic_source": None if traffic is None else traffic.get("source"),
                     "note": "EFFECTIVE figure: algorithmic bytes = Nmodel*Nfilt*12 B per star per pass, per "
                             "launch = that x the stars of the launch; 32 stars reuse a grid tile held in "
                             "registers, so the DRAM traffic (`traffic`, ncu, same per-launch basis) is far "
                             "lower and the kernel is FP32-issue bound (DESIGN.md section 5)",
                     "issue": None if (traffic is None or cfg["nfilt"] != 8 or not clocks.get("sm_mhz")) else {
                         "note": "what actually bounds the kernel: FP32 instruction issue. warp instructions per 32 "
                                 "(model, star) pairs from the committed ncu capture x pairs swept / kernel time, "
                                 "against 4 warp-instructions per clock per SM at the SM clock sampled during the run",
                         "warp_inst_per_32_pairs": traffic.get("warp_instructions_per_32_model_star_pairs"),
                         "achieved_ginst_s": traffic.get("warp_instructions_per_32_model_star_pairs", 0)
                         * (cfg["nmodel"] / 32.0) * agg["magfit_star_passes"] / (agg["ms_magfit"] * 1e-3) / 1e9,
                         "peak_ginst_s": 148 * 4 * clocks["sm_mhz"] * 1e6 / 1e9,
                         "frac": traffic.get("warp_instructions_per_32_model_star_pairs", 0)
                         * (cfg["nmodel"] / 32.0) * agg["magfit_star_passes"] / (agg["ms_magfit"] * 1e-3)
                         / (148 * 4 * clocks["sm_mhz"] * 1e6)},
                     "kernel_share_of_step": agg["ms_magfit"] / dev_ms,
                     "launches": int(agg["magfit_launches"]),
                     "ms_per_launch": agg["ms_magfit"] / launches},
        "phases_ms_per_step": {"magfit": agg["ms_magfit"] / args.steps, "flux": agg["ms_flux"] / args.steps,
                               "select": agg["ms_select"] / args.steps},
        "counts_per_step": {"candidates": agg["candidates"] / args.steps, "survivors": agg["survivors"] / args.steps,
                            "selected": agg["selected"] / args.steps, "resweeps": agg["resweeps"] / args.steps,
                            "fallbacks": agg["fallbacks"] / args.steps},
        "clocks": clocks,
        "region_wall_s": t_region,
    }
    if world == 1 and not args.no_cpu:
        nt_probe = host_cores()
        # ~10-15 s of CPU work: 16 stars per core at C2 (0.7 s per star and core), 4 per core on the 3M-model grids
        nsample = max(8, 16 * nt_probe) if cfg["nmodel"] <= 1_000_000 else max(4, 4 * nt_probe)
        rate, nt, dt = cpu_sample(cfg, grid, stars, 0, min(nsample, nstar))
        line["cpu_baseline"] = {"value": rate, "unit": "stars/s", "cores": nt, "kind": "port",
                                "sample": "first %d stars of the same catalogue and grid, %.1f s, OpenMP over "
                                          "stars, C port of the reference loglike" % (min(nsample, nstar), dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 5])
    ap.add_argument("--nstar", type=int, default=0, help="stars per step per GPU (default: the config's; "
                    "capped at 1000 for configs 3 and 5)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--rows", type=int, default=11, choices=[3, 5, 11],
                    help="record rows shipped to the host per selected model (11 = everything)")
    ap.add_argument("--nmc-prior", type=int, default=50, help="Nmc_prior of fit() (brutus/fitting.py:1429)")
    ap.add_argument("--ndraws", type=int, default=250, help="Ndraws of fit() (brutus/fitting.py:1431)")
    ap.add_argument("--grid", default="locus", choices=["locus", "tilt"],
                    help="mock grid family (brutus_b200/mock.py): a stellar locus (default) or the degenerate "
                         "colour-tilt grid the golden vectors use")
    args = ap.parse_args()
    global GRID_KIND
    GRID_KIND = args.grid
    if not args.nstar and args.config in (3, 5):
        args.nstar = 1000
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
