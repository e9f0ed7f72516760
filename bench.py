#!/usr/bin/env python
"""bench.py -- stars/sec of the full-grid brute-force likelihood sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2] [--scaling weak|strong]

A step = one pass of the hot path (loglike + label priors + lnpost's first threshold, i.e.
``bf_sweep_batch``) over one synthetic catalogue against the whole grid.
N=1 workload: BASELINE.json configs[1] (1k stars, 8 bands, 1M-point grid).  N>1: stars sharded, grid replicated by
ONE NCCL broadcast issued inside the library, no collective in the hot loop, no PyTorch anywhere in this arm:
  * under torchrun (one rank per GPU; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment) every rank
    holds a one-device handle joined into an NCCL process group by the library (bf_nccl_init; the NCCL id
    travels over a TCP socket), timings are reduced with bf_allreduce_max;
  * started plainly with --gpus N the single process drives N devices through one handle (bf_create_multi).
``--scaling weak`` (default): every GPU sweeps its own ``nstar`` stars; ``strong``: ``nstar`` is the whole catalogue.

  value            stars/s from CUDA-event device time (first kernel -> last kernel of each call,
                   grid and scratch resident in HBM), max over ranks / devices
  e2e              stars/s through the call a user makes, BruteForce.fit's per-object body
                   (bf_fit_batch: the same sweep, then lnpost with the default Galactic prior, evidence
                   and resampling on the device; host float64 photometry in, Ndraws posterior samples
                   per star out): wall clock of the calls, host preparation, H2D and D2H included
  e2e_records      the same for bf_sweep_batch, which ships every selected model's record to the host
                   for a host-side lnpost (user-supplied prior callables): PCIe-bound
  e2e_fit_api      BruteForce.fit itself (setup, float32 output assembly, incremental writer), one call
  roofline         dominant kernel k_sweep: algorithmic bytes (Nmodel x Nfilt x 12 B per star) / its
                   CUDA-event time, against the measured HBM copy peak; `step_frac` = the same over the whole step
  cpu_baseline     the C oracle (port of the reference's loglike) on this box's host cores
--impl reference   times only that CPU path (the reference is Python+numba and cannot travel to
                   the GPU box; oracle/loglike_ref.c is its pinned restatement).
"""
import argparse
import json
import os
import sys
import tempfile
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
    4: "C4: NGC 2682 demo catalogue (1517 objects with >= 4 bands) x 40 896-model Bayestar-shaped lattice x 8 bands",
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
    """DRAM bytes / instruction counts of k_sweep from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML (in process: nothing is forked inside the timed region)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        while nv is not None and not self.stop_flag.is_set():
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetMaxClockInfo(self.dev, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev),
                                     nv.nvmlDeviceGetUtilizationRates(self.dev).gpu))
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=3)
        nv = self.nv
        if nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        busy = [s for s in self.samples if s[3] > 0] or self.samples
        reasons = set()
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        for s in busy:
            for k, bit in names.items():
                if s[2] & bit:
                    reasons.add(k)
        return {"sm_mhz": float(np.median([s[0] for s in busy])), "sm_max_mhz": float(max(s[1] for s in busy)),
                "reasons": sorted(reasons), "samples": len(busy), "source": "NVML, rank 0, 10 Hz, samples with the GPU busy"}


def workload(cfg_id, nstar):
    from brutus_b200 import mock
    cfg = dict(mock.CONFIGS[cfg_id])
    if nstar:
        cfg["nstar"] = nstar
    return cfg


def make_inputs(cfg_id, cfg):
    from brutus_b200 import mock
    if cfg_id == 4:
        return mock.make_grid_lattice(*cfg["lattice"], nfilt=cfg["nfilt"])
    return mock.make_grid(cfg["nmodel"], cfg["nfilt"], seed=1000 + cfg_id, kind=GRID_KIND)


def model_priors(labels):
    """Static inputs of lnpost for the mock grid: fit()'s default `lnprior` (IMF over 'mini', else the PS1
    luminosity function over 'Mr': brutus/fitting.py:1335-1341) and the 'feh' / 'loga' labels of the Galactic prior."""
    from brutus_b200 import pdf
    names = labels.dtype.names
    lnprior = pdf.imf_lnprior(labels["mini"]) if "mini" in names else pdf.ps1_MrLF_lnprior(labels["Mr"])
    return dict(lnprior=lnprior, feh=labels["feh"] if "feh" in names else None,
                loga=labels["loga"] if "loga" in names else None)


def make_stars(cfg_id, cfg, grid, seed_shift, nstar):
    from brutus_b200 import mock
    if cfg_id == 4:   # the real catalogue, tiled if more stars are asked for
        st = mock.load_ngc2682()
        idx = np.arange(nstar) % len(st["flux"])
        return {k: v[idx] for k, v in st.items()}
    return mock.make_stars(grid, nstar, seed=2000 + cfg_id + 100 * seed_shift, av_max=cfg["av_max"],
                           dropout=cfg["dropout"])


def config_block(cfg_id, cfg):
    """Identical in both arms (the driver compares the two `config` objects)."""
    return {"workload": CONFIG_NAMES[cfg_id], "nmodel": cfg["nmodel"], "nfilt": cfg["nfilt"],
            "grid": "Bayestar-shaped (Mr, [Fe/H]) lattice mock" if cfg_id == 4 else "mock %s (brutus_b200/mock.py)" % GRID_KIND,
            "avlim": list(cfg["avlim"]), "stars": "synthetic, seeded (brutus_b200/mock.py)" if cfg_id != 4
            else "demos/NGC_2682.fits (tests/golden/ngc2682.npz)"}


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
    sl = slice(0, nsample)
    t0 = time.perf_counter()
    oracle.loglike_batch(stars["flux"][sl], stars["err"][sl], stars["mask"][sl], grid,
                         parallax=stars["parallax"][sl], parallax_err=stars["parallax_err"][sl],
                         nthreads=nthreads, avlim=cfg["avlim"])
    dt = time.perf_counter() - t0
    return nsample / dt, dt


def cpu_baseline_block(cfg, grid, stars, nstar):
    nt = host_cores()
    # ~10-15 s of CPU work on all cores: 16 stars per core at C2 (0.7 s per star and core), fewer on the 3M grids
    per_core = max(1, int(16 * 1_000_000 * 8 / (cfg["nmodel"] * cfg["nfilt"])))
    nsample = min(nstar, max(8, min(per_core, 64) * nt))
    rate, dt = cpu_sample(cfg, grid, stars, nt, nsample)
    n1 = min(nstar, max(2, min(per_core, 4)))
    rate1, dt1 = cpu_sample(cfg, grid, stars, 1, n1)   # the reference itself is single-threaded (numba, no prange)
    return {"value": rate, "unit": "stars/s", "cores": nt, "kind": "port",
            "sample": "first %d stars of the same catalogue and grid, %.1f s, OpenMP over stars, C port of the "
                      "reference loglike" % (nsample, dt),
            "one_core": {"value": rate1, "unit": "stars/s", "cores": 1,
                         "sample": "first %d stars, %.1f s (the reference's own loop is serial)" % (n1, dt1)}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = workload(args.config, args.nstar)
    grid, _ = make_inputs(args.config, cfg)
    from oracle import oracle
    oracle.build()
    nt = host_cores()
    # a bounded sample per step: ~3 s of work on all cores at C2 (the whole run stays within a minute or two)
    per_step = max(4 * nt, 8) if cfg["nmodel"] * cfg["nfilt"] <= 8_000_000 else max(nt, 4)
    stars = make_stars(args.config, cfg, grid, 0, per_step * (args.steps + args.warmup))
    times = []
    for it in range(args.steps + args.warmup):
        sl = slice(it * per_step, (it + 1) * per_step)
        sub = {k: stars[k][sl] for k in ("flux", "err", "mask", "parallax", "parallax_err")}
        _, dt = cpu_sample(cfg, grid, sub, nt, per_step)
        if it >= args.warmup:
            times.append(dt)
    tot = float(np.sum(times))
    value = per_step * args.steps / tot
    rate1, dt1 = cpu_sample(cfg, grid, stars, 1, min(4, per_step))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "stars/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_block(args.config, cfg),
            "cpu_baseline": {"value": value, "unit": "stars/s", "cores": nt, "kind": "port",
                             "sample": "%d stars per step x %d steps, OpenMP over stars, C port "
                                       "(oracle/loglike_ref.c) of the reference's numba loglike"
                                       % (per_step, args.steps),
                             "one_core": {"value": rate1, "unit": "stars/s", "cores": 1,
                                          "sample": "%d stars, %.1f s (the reference's own loop is serial)" % (min(4, per_step), dt1)}},
            "e2e": {"value": value, "unit": "stars/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def fit_api_e2e(cfg, grid, labels, stars, device, precision, nmc_prior, ndraws):
    """BruteForce.fit end to end (the reference's entry point: setup, the device fit, float32 output assembly,
    incremental writer) on the same catalogue; one warm call, one timed call."""
    from brutus_b200 import fitting
    lmask = np.ones(1, dtype=[(n, bool) for n in labels.dtype.names])
    bf = fitting.BruteForce(grid, labels, lmask, precision=precision, device=device)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for tag in ("warm", "timed"):
            t0 = time.perf_counter()
            bf.fit(stars["flux"], stars["err"], stars["mask"], np.arange(len(stars["flux"])), os.path.join(td, tag),
                   parallax=stars["parallax"], parallax_err=stars["parallax_err"], data_coords=stars["coords"],
                   avlim=cfg["avlim"], Nmc_prior=nmc_prior, Ndraws=ndraws, apply_agewt=False, apply_grad=False,
                   rstate=np.random.RandomState(7), verbose=False)
            out[tag] = time.perf_counter() - t0
    bf.close()
    n = len(stars["flux"])
    return {"value": n / out["timed"], "unit": "stars/s", "seconds": out["timed"], "stars": n,
            "call": "BruteForce.fit (setup + bf_fit_batch per batch + float32 assembly + incremental writer)"}


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ndev = args.gpus if world == 1 else 1          # devices driven by THIS process
    ngpu = world if world > 1 else ndev
    cfg = workload(args.config, args.nstar)
    from brutus_b200 import _lib, shard
    comm = None
    # NCCL may print a version banner on stdout at the first communicator: keep fd 1 for the one JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if world > 1:
            comm = shard.SocketComm.from_env()
            h = _lib.Handle(local_rank, args.precision)
            shard.init_process_group(h, comm)          # ncclCommInitRank inside the library
        else:
            h = _lib.Handle(list(range(ndev)) if ndev > 1 else 0, args.precision)
        # ---- stage the grid: built from its seed on every rank (each needs it to draw its synthetic stars), but only
        # rank 0's copy goes to a GPU: one H2D, ONE ncclBroadcast inside the library, device-side re-tiling ----
        grid, labels = make_inputs(args.config, cfg)
        t0 = time.perf_counter()
        shard.broadcast_grid(h, grid if rank == 0 else None, grid.shape)
        shard.broadcast_model_priors(h, cfg["nmodel"], **(model_priors(labels) if rank == 0 else {}))
        h.allreduce_max(np.zeros(1))
        t_stage = time.perf_counter() - t0
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    # ---- this process's stars ----
    strong = args.scaling == "strong"
    if strong:
        total = cfg["nstar"]
        allst = make_stars(args.config, cfg, grid, 0, total)
        lo, hi = shard.shard_bounds(total, world, rank)
        stars = {k: v[lo:hi] for k, v in allst.items() if isinstance(v, np.ndarray)}
        base0 = lo
    else:
        nper = cfg["nstar"] * ndev
        stars = make_stars(args.config, cfg, grid, rank, nper)
        total = nper * world
        base0 = rank * nper
    nloc = len(stars["flux"])
    opts = _lib.make_options(avlim=cfg["avlim"])
    opts_dev = _lib.make_options(avlim=cfg["avlim"], skip_d2h=True)

    def barrier():
        h.allreduce_max(np.zeros(1))

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
                          nmc_prior=args.nmc_prior, ndraws=args.ndraws, seed=12345, star_base=base0,
                          mem_lim=8000.,   # fit()'s default mem_lim (brutus/fitting.py:1436)
                          copy=False)      # the draws stay in the library's pinned arena (like the records)
        wall = time.perf_counter() - t
        return res, wall, h.stats()

    # the records-out leg ships ~7 MB per star to pinned host memory: bounded to catalogues of <= 2 000 stars per process
    do_records = not args.no_records and nloc <= 2000
    for _ in range(args.warmup):
        if do_records:
            step(True)
        step(False)
        step_fit()
    tracing = bool(int(os.environ.get("BRUTUS_B200_TRACE", "0") or 0))
    if tracing:
        h.trace()   # drop the warm-up's entries
    sampler = ClockSampler(local_rank if world > 1 else 0) if rank == 0 else None
    if sampler is not None:
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
    n_iter_val, n_surv_val = np.array(res["n_iter"]), np.array(res["n_surv"])   # of the last device-timed step
    if tracing:   # per-kernel CUDA-event times of the device-timed steps (stderr; the JSON line stays alone on stdout)
        for name, (cnt, ms) in sorted(h.trace().items(), key=lambda kv: -kv[1][1]):
            sys.stderr.write("trace value-steps  %-22s %6d launches %10.3f ms/step\n" % (name, cnt, ms / args.steps))
    # ---- timed region 2: K steps end to end through the C ABI with host buffers (records out) ----
    wall_s = 0.0
    agg_e = {}
    if do_records:
        for _ in range(args.steps):
            res, wall, st = step(True)
            wall_s += wall
            for k, v in st.items():
                agg_e[k] = agg_e.get(k, 0) + v
        barrier()
        if tracing:
            h.trace()
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
        for name, (cnt, ms) in sorted(h.trace().items(), key=lambda kv: -kv[1][1]):
            sys.stderr.write("trace fit-steps    %-22s %6d launches %10.3f ms/step\n" % (name, cnt, ms / args.steps))
    t_region = time.perf_counter() - t_region
    clocks = sampler.summary() if sampler is not None else None
    dev_ms, wall_s, wall_f = [float(x) for x in h.allreduce_max(np.array([dev_ms, wall_s, wall_f]))]
    if rank != 0:
        h.close()
        if comm is not None:
            comm.close()
        return 0

    steps_stars = total * args.steps
    value = steps_stars / (dev_ms * 1e-3)
    e2e_fit = steps_stars / wall_f
    peak, peak_src = read_peaks()
    bytes_per_star = cfg["nmodel"] * cfg["nfilt"] * 12
    launches = max(1, agg["magfit_launches"])
    # `agg` is this process's share: its stars (all devices of the handle), device times = slowest device
    loc_stars = nloc * args.steps
    ach = bytes_per_star * (loc_stars / ndev) / (agg["ms_magfit"] * 1e-3) / 1e9
    ach_pass = bytes_per_star * (agg["magfit_star_passes"] / ndev) / (agg["ms_magfit"] * 1e-3) / 1e9
    ach_step = bytes_per_star * (loc_stars / ndev) / (agg["ms_device"] * 1e-3) / 1e9
    traffic = read_traffic()
    issue = None
    if traffic is not None and cfg["nfilt"] == traffic.get("nfilt") and clocks and clocks.get("sm_mhz"):
        wi = traffic["warp_instructions_per_32_model_star_pairs"]
        g = wi * (cfg["nmodel"] / 32.0) * (agg["magfit_star_passes"] / ndev) / (agg["ms_magfit"] * 1e-3)
        issue = {"note": "what actually bounds the kernel: FP32 instruction issue / the FMA pipe.  Warp instructions "
                         "per 32 (model, star) pairs from the committed ncu capture x pairs swept / kernel time, "
                         "against 4 warp-instructions per clock per SM at the SM clock sampled during the run",
                 "warp_inst_per_32_pairs": wi, "achieved_ginst_s": g / 1e9,
                 "peak_ginst_s": 148 * 4 * clocks["sm_mhz"] * 1e6 / 1e9,
                 "frac": g / (148 * 4 * clocks["sm_mhz"] * 1e6)}
    # algorithmic flop-instructions per star = Nmodel x Nb x (5 + 15 K_mag + 35 + Nsurv/Nmodel x 43 K_flux),
    # with the iteration counts and survivor fractions the stars of this rank actually had (Nb taken as Nfilt)
    k_mag, k_flux = n_iter_val[:, 0].astype(float), n_iter_val[:, 1].astype(float)
    flop_star = cfg["nmodel"] * cfg["nfilt"] * (5. + 15. * k_mag + 35. + n_surv_val / cfg["nmodel"] * 43. * k_flux)
    algo_flop = {"per_star_mean": float(flop_star.mean()), "k_mag_mean": float(k_mag.mean()), "k_flux_mean": float(k_flux.mean()),
                 "survivor_frac_mean": float(n_surv_val.mean() / cfg["nmodel"]),
                 "achieved_tflop_instr_s": float(flop_star.mean()) * (loc_stars / ndev) / (agg["ms_device"] * 1e-3) / 1e12,
                 "note": "flop-instructions of the reference's arithmetic (an FMA counts once) over the whole device step, per GPU; "
                         "a B200 issues 148 SM x 128 FP32 lanes x ~1.9 GHz = 36 T lane-instructions/s"}
    line = {
        "metric": METRIC, "value": value, "unit": "stars/s", "n_gpus": ngpu, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": config_block(args.config, cfg),
        "run": {"stars_per_step": total, "stars_per_step_per_gpu": total / ngpu,
                "parallelism": "stars sharded x%d (%s), grid replicated by one in-library ncclBroadcast"
                               % (ngpu, "one process per GPU" if world > 1 else "one process"),
                "l2": "512 MB L2 flush before every step", "grid_stage_s": round(t_stage, 3)},
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
                        "reference arm (loglike only) is timed on; bytes and launches are rank 0's"},
        "gpu_launches": int(agg["kernel_launches"]),
        "roofline": {"bound": "hbm", "limiter": "fp32-issue", "kernel": "k_sweep", "achieved": ach, "peak": peak,
                     "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                     "frac_per_pass": ach_pass / peak, "step_frac": ach_step / peak,
                     "traffic": None if traffic is None else traffic.get("dram_bytes_per_star_pass", 0) * agg["magfit_star_passes"] / ndev / launches,
                     "traffic_source": None if traffic is None else traffic.get("source"),
                     "note": "`bound` names the roofline the contract asks for (HBM copy peak); what limits the kernel is FP32 "
                             "instruction issue / the FMA pipe (`limiter`, `issue`).  EFFECTIVE figures: achieved = Nmodel*Nfilt*12 B per star x the "
                             "stars of the step / ALL the time spent in k_sweep (re-sweeps of stars whose iteration count "
                             "was mis-speculated are NOT credited; frac_per_pass credits them); step_frac = the same "
                             "bytes / the whole device step (sweep + cull + flux continuation + selection + ordered "
                             "records).  32 stars reuse a grid tile held in registers and shared memory, so the DRAM "
                             "traffic (`traffic`, ncu, per launch) is far below the algorithmic bytes: the kernel is "
                             "bound by FP32 issue / the FMA pipe (`issue`; ncu: issue active 70 %, FMA pipe 58 %, 24 warps "
                             "per SM), DESIGN.md section 5.  A sweep round is two launches (a strided subsample of the "
                             "model tiles, then the rest): `launches` counts both",
                     "issue": issue,
                     "algorithmic_flop_instr": algo_flop,
                     "kernel_share_of_step": agg["ms_magfit"] / agg["ms_device"],
                     "launches": int(agg["magfit_launches"]),
                     "ms_per_launch": agg["ms_magfit"] / launches},
        "phases_ms_per_step": {"sweep": agg["ms_magfit"] / args.steps, "flux": agg["ms_flux"] / args.steps,
                               "select": agg["ms_select"] / args.steps},
        "counts_per_step": {"candidates": agg["candidates"] / args.steps, "survivors": agg["survivors"] / args.steps,
                            "selected": agg["selected"] / args.steps, "resweeps": agg["resweeps"] / args.steps,
                            "fixups": agg["fixups"] / args.steps, "fallbacks": agg["fallbacks"] / args.steps,
                            "regroups": agg["regroups"] / args.steps,
                            "host_syncs": agg.get("host_syncs", 0) / args.steps,
                            "host_syncs_fit_call": agg_f.get("host_syncs", 0) / args.steps},
        "clocks": clocks,
        "region_wall_s": t_region,
    }
    if do_records:
        line["e2e_records"] = {"value": steps_stars / wall_s, "unit": "stars/s", "ms_per_step": 1e3 * wall_s / args.steps,
                               "h2d_bytes_per_step": int(agg_e["h2d_bytes"] / args.steps),
                               "d2h_bytes_per_step": int(agg_e["d2h_bytes"] / args.steps),
                               "call": "bf_sweep_batch", "record_rows": args.rows,
                               "gpu_launches": int(agg_e["kernel_launches"]),
                               "note": "host float64 photometry in, every selected model's record (48 B) out to pinned "
                                       "host memory for a host-side lnpost (user prior callables); the D2H of one star "
                                       "batch overlaps the kernels of the next; PCIe-bound (~52 GB/s)"}
    if world == 1 and ndev == 1 and not args.no_fit_api:
        line["e2e_fit_api"] = fit_api_e2e(cfg, grid, labels, stars, 0, args.precision, args.nmc_prior, args.ndraws)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_block(cfg, grid, stars, nloc)
    print(json.dumps(line))
    h.close()
    if comm is not None:
        comm.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--nstar", type=int, default=0, help="stars per step: per GPU with --scaling weak, the whole "
                    "catalogue with --scaling strong (default: the config's, capped at 1000 for configs 3 and 5)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-records", action="store_true", help="skip the records-out (bf_sweep_batch to host) leg")
    ap.add_argument("--no-fit-api", action="store_true", help="skip the BruteForce.fit leg")
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
