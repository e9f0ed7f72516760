"""Generates tests/golden/*.npz by running the UNMODIFIED reference (numba/NumPy) from
/root/reference.  Run in the build container only:  python tests/gen_golden.py

The fixtures pin (a) the C oracle (tests/test_oracle.py) and (b) the CUDA path (-m gpu tests).
Inputs are regenerated from seeds by brutus_b200.mock; only the reference's outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402
from brutus_b200 import mock  # noqa: E402

GOLD = os.path.join(HERE, "golden")


def toy_galprior(dist, coord, labels=None):
    """User-supplied Galactic prior (the hook at brutus/fitting.py:870-878): volume element times
    an exponential fall-off.  Shared with the tests via tests/golden_cases.py."""
    return 2. * np.log(dist) - dist / 2.


# name -> (grid kwargs, star kwargs, per-star tweaks, loglike kwargs)
LOGLIKE_CASES = {
    # BASELINE.json configs[0]: 1 star, 5 bands, 10k-model grid
    "c1_plumbing": dict(grid=dict(nmodel=10_000, nfilt=5, seed=1001),
                        stars=dict(nstar=1, seed=2001, par_nan_frac=0.0), kw={}),
    "mixed_9band": dict(grid=dict(nmodel=2_000, nfilt=9, seed=1010),
                        stars=dict(nstar=6, seed=2010, dropout=0.15), kw={}, negflux=[(2, 3)]),
    "nodimprior": dict(grid=dict(nmodel=2_000, nfilt=8, seed=1011),
                       stars=dict(nstar=3, seed=2011), kw=dict(dim_prior=False)),
    "avlim6_rvprior": dict(grid=dict(nmodel=2_000, nfilt=12, seed=1012),
                           stars=dict(nstar=3, seed=2012, av_max=6.0, dropout=0.1),
                           kw=dict(avlim=(0., 6.), rv_gauss=(3.1, 0.3), av_gauss=(1.0, 2.0))),
    # stellar-locus mock (the bench workload's grid family): K_mag = 1 and 2 both occur
    "locus_8band": dict(grid=dict(nmodel=3_000, nfilt=8, seed=1014, kind="locus"),
                        stars=dict(nstar=5, seed=2014, dropout=0.05), kw={}),
    "loose_tol": dict(grid=dict(nmodel=2_000, nfilt=6, seed=1013),
                      stars=dict(nstar=3, seed=2013, snr_range=(5., 20.)),
                      kw=dict(ltol=1e-3, ltol_subthresh=5e-2, init_thresh=1e-4)),
    # caller-supplied per-model start of the magnitude fit (brutus/fitting.py:583, :700-703)
    "av_rv_init": dict(grid=dict(nmodel=2_000, nfilt=7, seed=1015, kind="locus"),
                       stars=dict(nstar=4, seed=2015, dropout=0.1), kw={}, init=dict(seed=3015)),
}


def build_case(name):
    spec = LOGLIKE_CASES[name]
    grid, labels = mock.make_grid(**spec["grid"])
    st = mock.make_stars(grid, **spec["stars"])
    for (i, j) in spec.get("negflux", []):
        st["flux"][i, j] = -0.1 * abs(st["flux"][i, j])
    kw = dict(spec["kw"])
    if "init" in spec:
        rs = np.random.RandomState(spec["init"]["seed"])
        kw["av_init"] = rs.uniform(0., 3., grid.shape[0])
        kw["rv_init"] = rs.uniform(2.6, 4.2, grid.shape[0])
    return grid, labels, st, kw


def gen_loglike(fit, only=None):
    for name in LOGLIKE_CASES:
        if only and name not in only:
            continue
        grid, labels, st, kw = build_case(name)
        gF = np.array(grid, order="F")  # as _fit does, brutus/fitting.py:1964
        out = {}
        for i in range(len(st["flux"])):
            m = st["mask"][i].copy()
            # array-valued keywords (av_init, rv_init) are the reference's working arrays: it updates them in place and
            # returns them as av / rv (brutus/fitting.py:202, :232, :809), so every star gets fresh copies
            kwi = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            r = fit.loglike(st["flux"][i], st["err"][i], m, gF, return_vals=True,
                            parallax=st["parallax"][i], parallax_err=st["parallax_err"][i], **kwi)
            for key, val in zip(("lnl", "ndim", "chi2", "scale", "av", "rv", "icov"), r):
                out["%s_%d" % (key, i)] = np.array(val)
            for k, v in kwi.items():
                if isinstance(v, np.ndarray):
                    out["%s_after_%d" % (k, i)] = v
            out["mask_%d" % i] = m
        np.savez_compressed(os.path.join(GOLD, "loglike_%s.npz" % name), **out)
        print("wrote", name)


FIT_CASE = dict(grid=dict(nmodel=3_000, nfilt=8, seed=1020), stars=dict(nstar=5, seed=2020),
                Nmc_prior=20, Ndraws=40, rseed=77)


def gen_fit(fit):
    grid, labels, st = None, None, None
    grid, labels = mock.make_grid(**FIT_CASE["grid"])
    st = mock.make_stars(grid, **FIT_CASE["stars"])
    lmask = np.ones(1, dtype=[("Mr", bool), ("feh", bool)])
    bf = fit.BruteForce(grid, labels, lmask)
    lnprior = -0.1 * (labels["Mr"] - 5.) ** 2
    out = {}
    names = ("sidxs", "scales", "avs", "rvs", "cov_sar", "Ndim", "lnprob", "levid", "chi2min",
             "dists", "reds", "dreds", "logwts")
    gen = bf._fit(st["flux"], st["err"], st["mask"].copy(), parallax=st["parallax"],
                  parallax_err=st["parallax_err"], Nmc_prior=FIT_CASE["Nmc_prior"],
                  lnprior=lnprior, Ndraws=FIT_CASE["Ndraws"], lngalprior=toy_galprior,
                  dustfile=None, data_coords=np.zeros((len(st["flux"]), 2)),
                  rstate=np.random.RandomState(FIT_CASE["rseed"]))
    for i, res in enumerate(gen):
        for key, val in zip(names, res):
            out["%s_%d" % (key, i)] = np.asarray(val)
    np.savez_compressed(os.path.join(GOLD, "fit_generator.npz"), **out)
    print("wrote fit_generator")


def gen_galprior(fit):
    """Golden values of the reference's default Galactic prior (brutus/pdf.py:476-749).  astropy is not
    installed, so `SkyCoord(...).galactocentric...represent_as(CylRep)` (:630-635) is served by a
    stand-in that applies oracle.galprior.galactic_to_cyl; every other line is the reference's own."""
    import types
    from brutus import pdf as rpdf  # the reference module
    from oracle import galprior as gp

    class _Val(object):
        def __init__(self, v):
            self.value = v

    class _Sky(object):
        def __init__(self, l=None, b=None, distance=None, frame=None):
            self.R, self.Z = gp.galactic_to_cyl(np.asarray(distance), (np.asarray(l).flat[0], np.asarray(b).flat[0]))
            self.galactocentric = self
            self.cartesian = self

        def represent_as(self, _):
            return types.SimpleNamespace(rho=_Val(self.R), z=_Val(self.Z))

    rpdf.SkyCoord = _Sky
    rpdf.units = types.SimpleNamespace(deg=1.0, kpc=1.0)
    rs = np.random.RandomState(4242)
    n = 400
    dists = 10. ** rs.uniform(-2., 1.7, n)
    lab = np.zeros(n, dtype=[("feh", "f8"), ("loga", "f8")])
    lab["feh"] = rs.uniform(-3., 0.6, n)
    lab["loga"] = rs.uniform(6.5, 10.2, n)          # a few ages beyond 13.8 Gyr -> -inf
    out = dict(dists=dists, feh=lab["feh"], loga=lab["loga"])
    coords = [(0., 0.), (30., 5.), (180., -60.), (271.3, 89.), (95., -1.5)]
    out["coords"] = np.array(coords)
    for k, c in enumerate(coords):
        out["full_%d" % k] = rpdf.gal_lnprior(dists, c, labels=lab)
        out["feh_only_%d" % k] = rpdf.gal_lnprior(dists, c, labels=lab[["feh"]])
        out["nolabels_%d" % k] = rpdf.gal_lnprior(dists, c)
    np.savez_compressed(os.path.join(GOLD, "galprior.npz"), **out)
    print("wrote galprior")


OFFSETS_CASE = dict(grid=dict(nmodel=3_000, nfilt=8, seed=1040), nobj=40, nsamps=25, Nmc=30, seed=4040, rseed=123)


def build_offsets_case():
    """Inputs of the photometric_offsets / get_seds golden case (regenerated from seeds by the tests)."""
    c = OFFSETS_CASE
    grid, labels = mock.make_grid(**c["grid"])
    rs = np.random.RandomState(c["seed"])
    nobj, nsamps, nfilt = c["nobj"], c["nsamps"], grid.shape[1]
    idxs = rs.randint(0, grid.shape[0], (nobj, nsamps))
    reds = rs.uniform(0., 2., (nobj, nsamps))
    dreds = rs.normal(3.3, 0.2, (nobj, nsamps))
    dists = 10. ** rs.uniform(-0.5, 0.5, (nobj, nsamps))
    # photometry: the first sample's SED, offset per band, plus noise
    co = grid[idxs[:, 0]].astype(np.float64)
    mag = co[:, :, 0] + reds[:, :1] * (co[:, :, 1] + dreds[:, :1] * co[:, :, 2])
    true_off = 1. + 0.05 * np.linspace(-1., 1., nfilt)
    phot = 10. ** (-0.4 * mag) / dists[:, :1] ** 2 / true_off[None, :]
    err = 0.03 * phot
    phot = phot + rs.normal(size=phot.shape) * err
    mask = rs.uniform(size=phot.shape) > 0.15
    mask[:, 0] |= mask.sum(axis=1) < 5
    weights = rs.uniform(0.2, 1., (nobj, nsamps))
    weights[3] = 0.
    sel = rs.uniform(size=nobj) > 0.1
    mask_fit = np.ones(nfilt, dtype=bool)
    mask_fit[-1] = False
    old = 1. + 0.01 * rs.normal(size=nfilt)
    return dict(grid=grid, phot=phot, err=err, mask=mask, idxs=idxs, reds=reds, dreds=dreds, dists=dists,
                sel=sel, weights=weights, mask_fit=mask_fit, old_offsets=old)


def gen_offsets(fit):
    """Golden outputs of the reference's get_seds and photometric_offsets (brutus/utils.py:1089-1400)."""
    from brutus import utils as rutils  # the reference module
    c = build_offsets_case()
    out = {}
    n = 200
    rs = np.random.RandomState(5)
    av, rv = rs.uniform(0., 3., n), rs.normal(3.3, 0.3, n)
    for flux in (False, True):
        s, r, d = rutils.get_seds(c["grid"][:n], av=av, rv=rv, return_flux=flux, return_rvec=True, return_drvec=True)
        out["seds_%d" % flux], out["rvecs_%d" % flux], out["drvecs_%d" % flux] = s, r, d
    out["seds_default"] = rutils.get_seds(c["grid"][:n])
    out["av"], out["rv"] = av, rv
    for name, kw in (("plain", {}), ("prior", dict(prior_mean=np.ones(8), prior_std=np.full(8, 0.02), dim_prior=False))):
        r = rutils.photometric_offsets(c["phot"], c["err"], c["mask"], c["grid"], c["idxs"], c["reds"], c["dreds"],
                                       c["dists"], sel=c["sel"], weights=c["weights"], mask_fit=c["mask_fit"],
                                       Nmc=OFFSETS_CASE["Nmc"], old_offsets=c["old_offsets"], verbose=False,
                                       rstate=np.random.RandomState(OFFSETS_CASE["rseed"]), **kw)
        out["ratios_" + name], out["ratios_err_" + name], out["nratio_" + name] = r
    np.savez_compressed(os.path.join(GOLD, "offsets.npz"), **out)
    print("wrote offsets")


LNPOST_CASE = dict(name="mixed_9band", Nmc_prior=15, rseed=5, mem_lims=(8000., 0.75))


def gen_lnpost(fit):
    """Golden outputs of the reference's `lnpost` (brutus/fitting.py:823-1107) on the `loglike` results of the
    mixed_9band case: pins brutus_b200.fitting.lnpost_selected (the host posterior path, and the checker of the
    device posterior) on the CPU.  The second mem_lim makes the memory clip (:1029-1036) bite."""
    c = LNPOST_CASE
    grid, labels, st, kw = build_case(c["name"])
    gF = np.array(grid, order="F")
    lnprior = -0.1 * (labels["Mr"] - 5.) ** 2
    out = {}
    for i in range(len(st["flux"])):
        m = st["mask"][i].copy()
        res = fit.loglike(st["flux"][i], st["err"][i], m, gF, return_vals=True, parallax=st["parallax"][i],
                          parallax_err=st["parallax_err"][i], **kw)
        for k, mem in enumerate(c["mem_lims"]):
            r = fit.lnpost(tuple(np.array(x) if np.ndim(x) else x for x in res), parallax=st["parallax"][i],
                           parallax_err=st["parallax_err"][i], coord=np.zeros(2), Nmc_prior=c["Nmc_prior"],
                           lnprior=lnprior, wt_thresh=1e-3, lngalprior=toy_galprior, apply_av_prior=False,
                           dlabels=labels, avlim=(0., 20.), rvlim=(1., 8.), mem_lim=mem,
                           rstate=np.random.RandomState(c["rseed"]))
            for key, val in zip(("sel", "cov_sar", "lnp", "dists", "reds", "dreds", "logwts"), r):
                out["%s_%d_%d" % (key, i, k)] = np.asarray(val)
    np.savez_compressed(os.path.join(GOLD, "lnpost.npz"), **out)
    print("wrote lnpost")


def gen_lnpost_cdf(fit):
    """`lnpost` with CDF thresholding (wt_thresh=None, cdf_thresh=2e-3: brutus/fitting.py:992-997, :1017-1022) on the
    first two stars of the lnpost case; the selection comes back in ascending order of probability."""
    c = LNPOST_CASE
    grid, labels, st, kw = build_case(c["name"])
    gF = np.array(grid, order="F")
    lnprior = -0.1 * (labels["Mr"] - 5.) ** 2
    out = {}
    for i in range(2):
        m = st["mask"][i].copy()
        res = fit.loglike(st["flux"][i], st["err"][i], m, gF, return_vals=True, parallax=st["parallax"][i],
                          parallax_err=st["parallax_err"][i], **kw)
        r = fit.lnpost(tuple(np.array(x) if np.ndim(x) else x for x in res), parallax=st["parallax"][i],
                       parallax_err=st["parallax_err"][i], coord=np.zeros(2), Nmc_prior=4,
                       lnprior=lnprior, wt_thresh=None, cdf_thresh=2e-3, lngalprior=toy_galprior, apply_av_prior=False,
                       dlabels=labels, avlim=(0., 20.), rvlim=(1., 8.), mem_lim=20.,
                       rstate=np.random.RandomState(c["rseed"]))
        for key, val in zip(("sel", "cov_sar", "lnp", "dists", "reds", "dreds", "logwts"), r):
            out["%s_%d" % (key, i)] = np.asarray(val)
    np.savez_compressed(os.path.join(GOLD, "lnpost_cdf.npz"), **out)
    print("wrote lnpost_cdf", [len(out["sel_%d" % i]) for i in range(2)])


NGC2682_FITS = os.path.join(ref_import.REFERENCE_ROOT, "demos", "NGC_2682.fits")
NGC2682_GRID = dict(nmodel=4_000, nfilt=8, seed=1050, kind="locus")


def read_fits_bintable(path):
    """Minimal reader of a FITS binary-table extension (no astropy here): header cards -> big-endian NumPy
    structured dtype."""
    import re
    raw = open(path, "rb").read()

    def header(off):
        cards = {}
        while True:
            blk = raw[off:off + 2880]
            off += 2880
            for i in range(36):
                c = blk[i * 80:(i + 1) * 80].decode("ascii")
                if c[:8].strip() == "END":
                    return cards, off
                if c[8:10] == "= ":
                    v = c[10:].split("/")[0].strip()
                    cards[c[:8].strip()] = v.strip("'").strip() if v.startswith("'") else v

    _, off = header(0)
    h, off = header(off)
    code = {"L": "i1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8"}
    fields = []
    for i in range(1, int(h["TFIELDS"]) + 1):
        m = re.match(r"(\d*)([ALBIJKED])", h["TFORM%d" % i])
        rep, c = int(m.group(1) or 1), m.group(2)
        if c == "A":
            fields.append((h["TTYPE%d" % i], "S%d" % rep))
        elif rep == 1:
            fields.append((h["TTYPE%d" % i], code[c]))
        else:
            fields.append((h["TTYPE%d" % i], code[c], (rep,)))
    dt = np.dtype(fields)
    assert dt.itemsize == int(h["NAXIS1"])
    return np.frombuffer(raw, dtype=dt, count=int(h["NAXIS2"]), offset=off)


def ngc2682_grid():
    """Mock grid for the NGC 2682 catalogue: the locus mock with magnitudes placed at 1 kpc (+10 mag), so that
    scale = parallax^2 holds for real parallaxes (the real Bayestar grid cannot be fetched offline)."""
    grid, labels = mock.make_grid(**NGC2682_GRID)
    grid[:, :, 0] += 10.
    return grid, labels


def gen_ngc2682(fit):
    """The 8 bands of demos/NGC_2682.fits that the Bayestar grid covers (PS grizy + 2MASS JHKs), assembled as in
    demos/Overview 5 cell 13: PS fluxes in maggies, 2MASS magnitudes -> maggies, 2 % / 3 %
    systematic floors, parallax zero-point +0.054 mas and 0.043 mas floor.  Stored with the reference's loglike
    outputs for three of its stars against the mock grid."""
    t = read_fits_bintable(NGC2682_FITS)
    n = len(t)
    phot = np.full((n, 8), np.nan)
    err = np.full((n, 8), np.nan)
    phot[:, :5] = t["ucal_fluxqz.median"]
    err[:, :5] = t["ucal_fluxqz.err"]
    for k, b in enumerate(("J", "H", "Ks")):
        m, me = t["2MASS_%s" % b].astype(float), t["2MASS_%s_Err" % b].astype(float)
        f = 10. ** (-0.4 * m)
        phot[:, 5 + k], err[:, 5 + k] = f, 0.4 * np.log(10.) * f * me
    floor = np.array([0.02] * 5 + [0.03] * 3)
    with np.errstate(all="ignore"):
        err = np.sqrt(err ** 2 + (floor[None, :] * phot) ** 2)
        mask = np.isfinite(phot) & np.isfinite(err) & (err > 0.) & (phot > 0.) & (err / phot < 0.5)
    par = t["Parallax"].astype(float) + 0.054
    perr = np.sqrt(t["Parallax_Err"].astype(float) ** 2 + 0.043 ** 2)
    coords = np.stack([t["l"].astype(float), t["b"].astype(float)], axis=1)
    out = dict(phot=phot, err=err, mask=mask, parallax=par, parallax_err=perr, coords=coords,
               memprob=t["HDBscan_MemProb"].astype(float))
    grid, _ = ngc2682_grid()
    gF = np.array(grid, order="F")
    nb = mask.sum(axis=1)
    picks = [int(np.where(nb == 8)[0][0]), int(np.where((nb >= 4) & (nb < 8))[0][0]),
             int(np.where((nb >= 5) & ~np.isfinite(par))[0][0]) if np.any((nb >= 5) & ~np.isfinite(par))
             else int(np.where(nb == 8)[0][7])]
    out["picks"] = np.array(picks)
    for i in picks:
        m = mask[i].copy()
        r = fit.loglike(phot[i], err[i], m, gF, return_vals=True, parallax=par[i], parallax_err=perr[i])
        for key, val in zip(("lnl", "ndim", "chi2", "scale", "av", "rv", "icov"), r):
            out["%s_%d" % (key, i)] = np.asarray(val)
        out["mask_%d" % i] = m
    np.savez_compressed(os.path.join(GOLD, "ngc2682.npz"), **out)
    print("wrote ngc2682: %d objects, bands-per-object %s, picks %s" % (n, np.bincount(nb).tolist(), picks))


def gen_priors():
    """Static priors of the fit path (brutus/pdf.py:38-260) on seeded inputs."""
    ref_import.import_reference()
    from brutus import pdf as rpdf
    rs = np.random.RandomState(41)
    out = dict(mini=10. ** rs.uniform(-1.3, 1., 400), mini2=10. ** rs.uniform(-1.3, 1., 400),
               Mr=rs.uniform(-3., 24., 400), parallaxes=rs.uniform(0.01, 5., 300),
               scales=10. ** rs.uniform(-3, 1, 300), scale_errs=10. ** rs.uniform(-4, 0, 300),
               par_cases=np.array([[1.3, 0.1], [0.2, 0.1], [np.nan, 0.1], [-0.1, 0.3], [2.0, np.nan]]))
    out["imf"] = rpdf.imf_lnprior(out["mini"])
    out["imf_binary"] = rpdf.imf_lnprior(out["mini"], mgrid2=out["mini2"])
    out["ps1"] = rpdf.ps1_MrLF_lnprior(out["Mr"])
    for k, (pm, pe) in enumerate(out["par_cases"]):
        out["parallax_lnprior_%d" % k] = rpdf.parallax_lnprior(out["parallaxes"], pm, pe)
        out["scale_parallax_lnprior_%d" % k] = rpdf.scale_parallax_lnprior(out["scales"], out["scale_errs"], pm, pe)
    np.savez_compressed(os.path.join(GOLD, "priors.npz"), **out)
    print("wrote priors")


if __name__ == "__main__":
    fit = ref_import.import_reference()
    os.makedirs(GOLD, exist_ok=True)
    only = [a for a in sys.argv[1:] if a in LOGLIKE_CASES or a in ("galprior", "offsets", "ngc2682", "lnpost", "lnpost_cdf", "priors", "fitopts")]
    gen_loglike(fit, only=only)
    if not only:
        gen_fit(fit)
    if not only or "galprior" in sys.argv[1:]:
        gen_galprior(fit)
    if not only or "offsets" in sys.argv[1:]:
        gen_offsets(fit)
    if not only or "lnpost" in sys.argv[1:]:
        gen_lnpost(fit)
    if (not only or "ngc2682" in sys.argv[1:]) and os.path.exists(NGC2682_FITS):
        gen_ngc2682(fit)
    if not only or "priors" in sys.argv[1:]:
        gen_priors()
    if not only or "lnpost_cdf" in sys.argv[1:]:
        gen_lnpost_cdf(fit)
