"""GPU parity tests of the device posterior (bf_fit_batch: lnpost after its first selection, evidence,
resampling).

The checker is the host path of brutus_b200.fitting (`lnpost_selected` + the tail of `_fit`), which
tests/test_fit_gpu.py holds to the golden 13-tuples of the unmodified reference, evaluated with the
NumPy restatement of the reference's default Galactic prior (oracle/galprior.py, pinned to golden values
of brutus.pdf.gal_lnprior).  Both sides consume the SAME random numbers: the device through the
z_override / u_override hooks, the host through oracle.galprior.ReplayRState, so every member of the
13-tuple is comparable draw for draw.

Tolerances.  float64 kernels: 1e-7 relative (observed ~1e-12), identical second selection and identical
drawn model indices.  float32 kernels: lnprob / levid |d| <= 5e-3, dists rel 1e-3, the second
selection may differ only within 5e-3 of its threshold, and a drawn index may differ only where the
uniform falls within 1e-4 of a CDF step.
"""
import numpy as np
import pytest

from brutus_b200 import mock
from oracle import galprior as gp

pytestmark = pytest.mark.gpu

NMC, NDRAWS = 20, 40


def _case(nmodel=3000, nstar=5, seed=1030):
    grid, labels = mock.make_grid(nmodel, 8, seed=seed, kind="locus")
    st = mock.make_stars(grid, nstar, seed=seed + 1000)
    lab = np.zeros(nmodel, dtype=[("mini", "f8"), ("feh", "f8"), ("loga", "f8")])
    lab["mini"], lab["feh"] = labels["mini"], labels["feh"]
    # main-sequence-lifetime-like ages; a few models older than 13.8 Gyr get a -inf age prior
    lab["loga"] = np.clip(10.0 - 2.5 * np.log10(labels["mini"]) + 0.3 * (labels["eep"] - 500.) / 300., 6.5, 10.2)
    rs = np.random.RandomState(seed + 7)
    coords = np.stack([rs.uniform(0., 360., nstar), rs.uniform(-80., 80., nstar)], axis=1)
    lnprior = -2.3 * np.log(labels["mini"])
    return grid, lab, st, coords, lnprior


def _galprior(dist, coord, labels=None):
    return gp.gal_lnprior(dist, coord, labels=labels)


def _run_both(precision, case, nmc=NMC, ndraws=NDRAWS):
    from brutus_b200 import fitting
    grid, lab, st, coords, lnprior = case
    nstar = len(st["flux"])
    lmask = np.ones(1, dtype=[(n, bool) for n in lab.dtype.names])
    bf = fitting.BruteForce(grid, lab, lmask, precision=precision)
    rs = np.random.RandomState(99)
    z = rs.normal(size=(grid.shape[0], 3, nmc))
    u = rs.uniform(size=(nstar, 2, ndraws))
    kw = dict(parallax=st["parallax"], parallax_err=st["parallax_err"], Nmc_prior=nmc, lnprior=lnprior,
              Ndraws=ndraws, dustfile=None, data_coords=coords)
    try:
        # --- device posterior, host-supplied random numbers ---
        bf._z_override, bf._u_override = z, u
        dev = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), lngalprior=None, **kw))
        nsel_dev = bf._get_handle().stats()["selected2"]
        bf._z_override = bf._u_override = None
        # --- host posterior: pass 1 finds the second selection of every star (it does not depend on the
        # generator), pass 2 replays the same normals / uniforms in the reference's order ---
        h = bf._get_handle()
        res = h.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], copy=True)
        sels = []
        for i in range(nstar):
            lo, hi = res["offsets"][i], res["offsets"][i + 1]
            out = fitting.lnpost_selected(
                res["model_idx"][lo:hi], res["lnl"][lo:hi], res["scale"][lo:hi], res["av"][lo:hi],
                res["rv"][lo:hi], fitting._unpack_icov(res["icov6"][:, lo:hi]), parallax=st["parallax"][i],
                parallax_err=st["parallax_err"][i], coord=coords[i], Nmc_prior=nmc, lnprior=lnprior,
                lngalprior=_galprior, dlabels=lab, rstate=np.random.RandomState(1), apply_av_prior=False)
            sels.append(out[0])
        replay = gp.ReplayRState(z, sels, u[:, 0], u[:, 1])
        host = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), lngalprior=_galprior, rstate=replay, **kw))
    finally:
        bf.close()
    return dev, host, sels, nsel_dev


NAMES = ("sidxs", "scales", "avs", "rvs", "cov_sar", "Ndim", "lnprob", "levid", "chi2min", "dists", "reds",
         "dreds", "logwts")



def test_device_posterior_replay_f64():
    dev, host, sels, nsel_dev = _run_both("f64", _case())
    assert nsel_dev == sum(len(s) for s in sels)
    for i, (d, h) in enumerate(zip(dev, host)):
        assert np.array_equal(d[0], h[0]), (i, "sidxs")
        assert d[5] == h[5], (i, "Ndim")
        for k in (1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12):
            a, b = np.asarray(d[k], dtype=np.float64), np.asarray(h[k], dtype=np.float64)
            err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))
            assert err < 1e-7, (i, NAMES[k], float(err))


def test_device_posterior_replay_f32():
    dev, host, sels, nsel_dev = _run_both("f32", _case())
    n_host = sum(len(s) for s in sels)
    assert abs(nsel_dev - n_host) <= max(3, n_host // 300)
    nmis = 0
    for i, (d, h) in enumerate(zip(dev, host)):
        same = d[0] == h[0]
        nmis += int((~same).sum())
        assert d[5] == h[5]
        assert abs(d[7] - h[7]) < 5e-3, (i, "levid", d[7], h[7])
        assert abs(d[8] - h[8]) < 2e-3 + 2e-5 * abs(h[8]), (i, "chi2min")
        for k, tol in ((1, 1e-4), (6, None), (9, 1e-3)):
            a, b = np.asarray(d[k])[same], np.asarray(h[k])[same]
            if tol is None:
                assert np.max(np.abs(a - b)) < 5e-3, (i, NAMES[k])
            else:
                assert np.max(np.abs(a - b) / np.abs(b)) < tol + 1.3 * 2e-4, (i, NAMES[k])
        for k in (2, 3, 10, 11):
            a, b = np.asarray(d[k])[same], np.asarray(h[k])[same]
            assert np.max(np.abs(a - b)) < 2e-3, (i, NAMES[k])
        cs = np.sqrt(np.abs(np.einsum("nii->ni", h[4])))
        rel = np.abs(d[4] - h[4]) / (cs[:, :, None] * cs[:, None, :])
        assert rel[same].max() < 5e-3, (i, "cov_sar")
    assert nmis <= 2, nmis


def test_device_posterior_philox_distribution():
    """Production mode (counter-based generator): the evidence agrees with the host Monte Carlo run with
    NumPy's generator within the Monte Carlo error, the draws are records of selected models, the result
    is reproducible, and it does not depend on how the catalogue is batched."""
    from brutus_b200 import fitting
    grid, lab, st, coords, lnprior = _case(nstar=6)
    lmask = np.ones(1, dtype=[(n, bool) for n in lab.dtype.names])
    bf = fitting.BruteForce(grid, lab, lmask, precision="f32")
    kw = dict(parallax=st["parallax"], parallax_err=st["parallax_err"], Nmc_prior=400, lnprior=lnprior,
              Ndraws=200, dustfile=None, data_coords=coords)
    try:
        a = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), rstate=np.random.RandomState(5), **kw))
        b = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), rstate=np.random.RandomState(5), **kw))
        c = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), rstate=np.random.RandomState(5), batch=4, **kw))
        host = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), lngalprior=_galprior,
                            rstate=np.random.RandomState(6), **kw))
    finally:
        bf.close()
    for i in range(len(a)):
        for k in range(13):
            assert np.array_equal(np.asarray(a[i][k]), np.asarray(b[i][k])), "not reproducible"
            assert np.array_equal(np.asarray(a[i][k]), np.asarray(c[i][k])), "depends on batching"
        assert abs(a[i][7] - host[i][7]) < 0.05, (i, a[i][7], host[i][7])
        assert a[i][5] == host[i][5]
        assert abs(a[i][8] - host[i][8]) < 2e-3 + 2e-5 * abs(host[i][8])
        # posterior means of distance agree within the resampling noise
        sd = np.std(np.log(host[i][9])) / np.sqrt(200.) * 5. + 0.02
        assert abs(np.mean(np.log(a[i][9])) - np.mean(np.log(host[i][9]))) < sd, (i, "dist")
        assert np.all(a[i][0] >= 0) and np.all(np.isfinite(a[i][9])) and np.all(a[i][9] > 0)


def test_fit_batch_argument_errors():
    from brutus_b200 import _lib
    grid, lab, st, coords, lnprior = _case(nmodel=500, nstar=2)
    h = _lib.Handle(0, "f32")
    try:
        h.set_grid(grid)
        with pytest.raises(ValueError):   # brutus/fitting.py:1362-1365
            h.fit_batch(st["flux"], st["err"], st["mask"], coords=None, use_gal_prior=True)
        with pytest.raises(ValueError):
            h.fit_batch(st["flux"], st["err"], st["mask"], coords=coords, nmc_prior=0)
        out = h.fit_batch(st["flux"], st["err"], st["mask"], coords=None, use_gal_prior=False, nmc_prior=5, ndraws=7)
        assert out["sidxs"].shape == (2, 7) and np.all(out["sidxs"] >= 0)
    finally:
        h.close()


def test_fit_batch_edge_cases():
    """Empty catalogue; more draws than a CTA has threads with an odd Nmc_prior; a prior that excludes every model
    (the reference would fail on an empty selection: the device reports sidxs = -99 and levid = -1e300)."""
    from brutus_b200 import _lib
    grid, lab, st, coords, lnprior = _case(nmodel=800, nstar=3)
    h = _lib.Handle(0, "f32")
    try:
        h.set_grid(grid)
        h.set_model_priors(lnprior=lnprior, feh=lab["feh"], loga=lab["loga"])
        out = h.fit_batch(st["flux"][:0], st["err"][:0], st["mask"][:0], coords=coords[:0], nmc_prior=4, ndraws=5)
        assert out["sidxs"].shape == (0, 5) and out["levid"].shape == (0,)
        out = h.fit_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], coords=coords,
                          nmc_prior=3, ndraws=1500, seed=9)
        assert out["sidxs"].shape == (3, 1500)
        ok = out["nsel"] > 0       # (this case's labels put some models beyond 13.8 Gyr: a star may select none)
        assert ok.sum() >= 2 and np.all(out["sidxs"][ok] >= 0) and np.all(out["sidxs"][ok] < 800)
        assert np.all(out["sidxs"][~ok] == -99) and np.all(out["levid"][~ok] <= -1e299)
        assert np.all(np.isfinite(out["dists"][ok])) and np.all(np.isfinite(out["levid"][ok]))
        # every drawn (scale, av, rv) is the record of its model: same model -> same values within a star
        for k in np.where(ok)[0]:
            _, first = np.unique(out["sidxs"][k], return_index=True)
            for j in first[:20]:
                same = out["sidxs"][k] == out["sidxs"][k][j]
                assert np.all(out["scales"][k][same] == out["scales"][k][j])
        h.set_model_priors(lnprior=np.full(800, -np.inf))
        out = h.fit_batch(st["flux"], st["err"], st["mask"], coords=coords, nmc_prior=4, ndraws=6)
        assert np.all(out["nsel"] == 0) and np.all(out["sidxs"] == -99) and np.all(out["levid"] <= -1e299)
    finally:
        h.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_memory_clip_matches_host(precision):
    """lnpost's mem_lim clip (brutus/fitting.py:1029-1036): a star whose second selection exceeds
    Nsel_max = int(mem_lim / Nmc_prior / 4e-4) keeps its Nsel_max best models by lnlike + lnprior.  The host path
    re-orders the kept models by rank, so draws are not comparable one to one; the kept set, the evidence (same
    normals on both sides) and chi2min are."""
    from brutus_b200 import fitting
    grid, lab, st, coords, lnprior = _case()
    nstar, nmc, ndraws = len(st["flux"]), 20, 30
    nsel_max = 150
    mem_lim = (nsel_max + 0.5) * nmc * 4.0e-4
    lmask = np.ones(1, dtype=[(n, bool) for n in lab.dtype.names])
    bf = fitting.BruteForce(grid, lab, lmask, precision=precision)
    rs = np.random.RandomState(7)
    z = rs.normal(size=(grid.shape[0], 3, nmc))
    u = rs.uniform(size=(nstar, 2, ndraws))
    kw = dict(parallax=st["parallax"], parallax_err=st["parallax_err"], Nmc_prior=nmc, lnprior=lnprior, Ndraws=ndraws,
              dustfile=None, data_coords=coords, mem_lim=mem_lim)
    try:
        bf._z_override, bf._u_override = z, u
        dev = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), **kw))
        stats = bf._get_handle().stats()
        bf._z_override = bf._u_override = None
        h = bf._get_handle()
        res = h.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], copy=True)
        sels = []
        for i in range(nstar):
            lo, hi = res["offsets"][i], res["offsets"][i + 1]
            sels.append(fitting.lnpost_selected(
                res["model_idx"][lo:hi], res["lnl"][lo:hi], res["scale"][lo:hi], res["av"][lo:hi], res["rv"][lo:hi],
                fitting._unpack_icov(res["icov6"][:, lo:hi]), parallax=st["parallax"][i], parallax_err=st["parallax_err"][i],
                coord=coords[i], Nmc_prior=nmc, lnprior=lnprior, lngalprior=_galprior, dlabels=lab, mem_lim=mem_lim,
                rstate=np.random.RandomState(1), apply_av_prior=False)[0])
        host = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), lngalprior=_galprior,
                            rstate=gp.ReplayRState(z, sels, u[:, 0], u[:, 1]), **kw))
    finally:
        bf.close()
    assert max(len(s) for s in sels) == nsel_max and stats["clipped"] >= 2
    assert abs(stats["selected2"] - sum(len(s) for s in sels)) <= (0 if precision == "f64" else 3)
    for i, (d, hh) in enumerate(zip(dev, host)):
        tol = 1e-7 if precision == "f64" else 5e-3
        assert abs(d[7] - hh[7]) < tol * max(1., abs(hh[7])), (i, "levid", d[7], hh[7])
        assert abs(d[8] - hh[8]) < (1e-7 if precision == "f64" else 2e-3) * max(1., abs(hh[8])), (i, "chi2min")
        assert set(d[0]) <= set(sels[i]) or precision == "f32"      # every drawn model belongs to the kept set


def test_device_posterior_replay_ext_prior_f64():
    """lnprior_ext (Gaussian priors on label columns, brutus/fitting.py:1995-2009) enters lnlike inside the sweep
    and therefore both selections, the evidence and the draws: device and host paths agree draw for draw."""
    from brutus_b200 import fitting
    grid, lab, st, coords, lnprior = _case(nstar=4)
    nstar, nmc, ndraws = len(st["flux"]), 10, 20
    lmask = np.ones(1, dtype=[(n, bool) for n in lab.dtype.names])
    bf = fitting.BruteForce(grid, lab, lmask, precision="f64")
    rs = np.random.RandomState(17)
    z = rs.normal(size=(grid.shape[0], 3, nmc))
    u = rs.uniform(size=(nstar, 2, ndraws))
    ext = {"feh": np.array([[-0.3, 0.4], [np.nan, 0.2], [-1.0, 0.3], [0.1, 0.5]]),
           "mini": np.array([[1.0, 0.8], [0.7, 0.5], [np.nan, np.nan], [1.5, 1.0]])}
    kw = dict(parallax=st["parallax"], parallax_err=st["parallax_err"], Nmc_prior=nmc, lnprior=lnprior, Ndraws=ndraws,
              dustfile=None, data_coords=coords, lnprior_ext=ext)
    try:
        bf._z_override, bf._u_override = z, u
        dev = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), **kw))
        bf._z_override = bf._u_override = None
        # selections of the host path (pass 1), then the replay (pass 2)
        sels, orig = [], fitting.lnpost_selected

        def recording(*a, **k):
            out = orig(*a, **k)
            sels.append(out[0])
            return out
        fitting.lnpost_selected = recording
        try:
            first = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), lngalprior=_galprior,
                                 rstate=np.random.RandomState(1), **kw))
        finally:
            fitting.lnpost_selected = orig
        sels = list(sels)
        host = list(bf._fit(st["flux"], st["err"], st["mask"].copy(), lngalprior=_galprior,
                            rstate=gp.ReplayRState(z, sels, u[:, 0], u[:, 1]), **kw))
    finally:
        bf.close()
    assert len(first) == nstar
    for i, (d, h) in enumerate(zip(dev, host)):
        assert np.array_equal(d[0], h[0]), (i, "sidxs")
        for k in (1, 2, 3, 6, 7, 8, 9, 10, 11, 12):
            a, b = np.asarray(d[k], dtype=np.float64), np.asarray(h[k], dtype=np.float64)
            assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12)) < 1e-7, (i, NAMES[k])


def test_fit_shards_equal_single_call():
    """Star-sharded fit (brutus_b200.shard.fit_shard: star_base = first catalogue index of the shard) gives every
    star the same posterior samples as one call over the whole catalogue -- here three shards on one GPU."""
    from brutus_b200 import _lib, shard
    grid, lab, st, coords, lnprior = _case(nmodel=2000, nstar=7)
    h = _lib.Handle(0, "f32")
    try:
        h.set_grid(grid)
        h.set_model_priors(lnprior=lnprior, feh=lab["feh"], loga=np.minimum(lab["loga"], 10.1))
        kw = dict(nmc_prior=12, ndraws=20, seed=77)
        one = h.fit_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], coords=coords, **kw)
        parts = [shard.fit_shard(h, st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"],
                                 coords=coords, world=3, rank=r, **kw) for r in range(3)]
    finally:
        h.close()
    assert [p[:2] for p in parts] == [shard.shard_bounds(7, 3, r) for r in range(3)]
    for k in one:
        assert np.array_equal(np.concatenate([p[2][k] for p in parts]), one[k]), k
