"""CPU: the host posterior path (brutus_b200.fitting.lnpost_selected) against golden outputs of the unmodified
reference `lnpost` (brutus/fitting.py:823-1107; tests/golden/lnpost.npz from tests/gen_golden.py::gen_lnpost), with
the reference's own `loglike` outputs (tests/golden/loglike_mixed_9band.npz) as input and the same RandomState.
This function is also the checker of the device posterior (tests/test_posterior_gpu.py), so it is pinned here
without a GPU, memory clip (brutus/fitting.py:1029-1036) included."""
import os

import numpy as np
import pytest

import gen_golden
import golden_cases as gc
from brutus_b200 import fitting


@pytest.mark.parametrize("k", [0, 1])
def test_lnpost_selected_matches_reference(k):
    c = gen_golden.LNPOST_CASE
    grid, labels, st, kw = gc.build_case(c["name"])
    ll = gc.load_loglike(c["name"])
    gold = np.load(os.path.join(gc.GOLD, "lnpost.npz"))
    lnprior = -0.1 * (labels["Mr"] - 5.) ** 2
    mem = c["mem_lims"][k]
    clipped = 0
    for i in range(len(st["flux"])):
        lnl, scale, av, rv, icov = (ll["%s_%d" % (n, i)] for n in ("lnl", "scale", "av", "rv", "icov"))
        par, perr = st["parallax"][i], st["parallax_err"][i]
        # lnpost's first stage (brutus/fitting.py:976-991), which bf_sweep_batch performs on the GPU
        lnprob = lnl.copy()
        if np.isfinite(par) and np.isfinite(perr):
            lnprob = lnl + fitting.scale_parallax_lnprior(scale, 1. / np.sqrt(np.abs(icov[:, 0, 0])), par, perr)
        lnprob[~np.isfinite(lnprob)] = -1e300
        sel = np.where(lnprob > np.log(1e-3) + np.max(lnprob))[0]
        out = fitting.lnpost_selected(sel, lnl[sel], scale[sel], av[sel], rv[sel], icov[sel], parallax=par,
                                      parallax_err=perr, coord=np.zeros(2), Nmc_prior=c["Nmc_prior"], lnprior=lnprior,
                                      lngalprior=gc.toy_galprior, dlabels=labels, mem_lim=mem,
                                      rstate=np.random.RandomState(c["rseed"]), apply_av_prior=False)
        sel2, keep, cov, lnp, dists, reds, dreds, logwts = out
        g = {n: gold["%s_%d_%d" % (n, i, k)] for n in ("sel", "cov_sar", "lnp", "dists", "reds", "dreds", "logwts")}
        assert np.array_equal(sel2, g["sel"]), (i, k)
        nsel_max = int(mem / c["Nmc_prior"] / 4.0e-4)
        assert len(sel2) <= nsel_max
        clipped += int(len(sel2) == nsel_max)
        for name, a in (("cov_sar", cov), ("lnp", lnp), ("dists", dists), ("reds", reds), ("dreds", dreds),
                        ("logwts", logwts)):
            b = g[name]
            assert a.shape == b.shape, (i, k, name)
            fin = np.isfinite(b)
            assert np.array_equal(fin, np.isfinite(a)), (i, k, name)
            assert np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-12)) < 1e-9, (i, k, name)
    assert clipped >= (1 if k == 1 else 0)      # with the small mem_lim the clip did bite


def test_lnpost_cdf_thresholding_matches_reference():
    """wt_thresh=None: both selections by cumulative probability (brutus/fitting.py:992-997, :1017-1022), memory
    clip included; the device ships every model in this mode and the host takes both selections."""
    c = gen_golden.LNPOST_CASE
    grid, labels, st, kw = gc.build_case(c["name"])
    ll = gc.load_loglike(c["name"])
    gold = np.load(os.path.join(gc.GOLD, "lnpost_cdf.npz"))
    lnprior = -0.1 * (labels["Mr"] - 5.) ** 2
    for i in range(2):
        lnl, scale, av, rv, icov = (ll["%s_%d" % (n, i)] for n in ("lnl", "scale", "av", "rv", "icov"))
        sel = np.arange(len(lnl))
        out = fitting.lnpost_selected(sel, lnl, scale, av, rv, icov, parallax=st["parallax"][i],
                                      parallax_err=st["parallax_err"][i], coord=np.zeros(2), Nmc_prior=4,
                                      lnprior=lnprior, wt_thresh=None, cdf_thresh=2e-3, lngalprior=gc.toy_galprior,
                                      dlabels=labels, mem_lim=20., rstate=np.random.RandomState(c["rseed"]),
                                      apply_av_prior=False)
        sel2, keep, cov, lnp, dists, reds, dreds, logwts = out
        assert np.array_equal(sel2, gold["sel_%d" % i]), i
        assert 0 < len(sel2) <= int(20. / 4 / 4.0e-4)
        assert not np.all(np.diff(sel2) > 0)               # the reference returns this selection in probability order
        for name, a in (("cov_sar", cov), ("lnp", lnp), ("dists", dists), ("reds", reds), ("dreds", dreds),
                        ("logwts", logwts)):
            b = gold["%s_%d" % (name, i)]
            fin = np.isfinite(b)
            assert a.shape == b.shape and np.array_equal(fin, np.isfinite(a)), (i, name)
            assert np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-12)) < 1e-9, (i, name)
