"""Seeded synthetic grids and star catalogues.

Used by the tests, ``bench.py`` and ``__graft_entry__.smoke()``; there is no network, so the
reference's real grids (``grid_mist_v9.h5`` ...) cannot be fetched.  The grid has the reference's
format: float32 ``(Nmodel, Nfilt, 3)`` holding ``(mag0 @ 1 kpc, R0, dR/dRv)`` per band
(produced at brutus/seds.py:828-832, consumed at brutus/utils.py:293-298).
"""
import numpy as np

__all__ = ["make_grid", "make_grid_lattice", "make_stars", "load_ngc2682", "CONFIGS"]

# BASELINE.json configs (index -> workload shape)
CONFIGS = {
    1: dict(nmodel=10_000, nfilt=5, nstar=1, avlim=(0., 20.), av_max=2.0, dropout=0.0),
    2: dict(nmodel=1_000_000, nfilt=8, nstar=1_000, avlim=(0., 20.), av_max=2.0, dropout=0.0),
    3: dict(nmodel=3_000_000, nfilt=12, nstar=100_000, avlim=(0., 6.), av_max=6.0, dropout=0.1),
    # NGC 2682 demo catalogue against a Bayestar-shaped (Mr, [Fe/H]) lattice: 568 x 72 = 40 896 models, PS grizy +
    # 2MASS JHKs (demos/Overview 1 cell 39, brutus/filters.py:9-10); 3.9 MB grid: L2-resident, launch/latency-bound
    4: dict(nmodel=40_896, nfilt=8, nstar=1_585, avlim=(0., 20.), av_max=2.0, dropout=0.0, lattice=(568, 72)),
    5: dict(nmodel=3_000_000, nfilt=8, nstar=1_000_000, avlim=(0., 20.), av_max=2.0, dropout=0.0),
}


# effective wavelengths (micron) of the mock filters: PS1 grizy, 2MASS JHKs, then Gaia G/BP/RP, W1, ...
_WAVE = np.array([0.481, 0.617, 0.752, 0.866, 0.962, 1.235, 1.662, 2.159, 0.622, 0.511, 0.777, 3.353,
                  4.603, 0.354, 0.440, 0.550])


def make_grid_locus(nmodel, nfilt, seed=1000):
    """Mock grid on a stellar locus, the shape of a real MIST grid: every model is a point of a
    3-parameter family (initial mass, evolutionary phase, [Fe/H]) whose colours come from a
    blackbody at the model's Teff plus a metallicity-dependent blanketing term, with absolute
    magnitudes from L(mass, phase); the reddening vector follows a power-law extinction curve
    evaluated at the mock filters' effective wavelengths.  Unlike :func:`make_grid` (whose single
    colour tilt is collinear with the reddening vector, so that a sixth of the grid fits any star),
    only the models near the star's Teff-A(V) degeneracy track survive, as with real grids.
    Labels: 'mini', 'eep', 'feh', 'Mr', 'loga' (``load_models``-like, brutus/utils.py:608-609; 'loga' is a
    main-sequence-lifetime-like log10 age in years, a deterministic function of mass and phase)."""
    rs = np.random.RandomState(seed)
    lam = _WAVE[:nfilt]
    u1, u2 = rs.uniform(size=nmodel), rs.uniform(size=nmodel)
    feh = rs.uniform(-2., 0.5, nmodel)
    mini = 10. ** (-1. + 1.9 * u1)                               # 0.1 .. 8 Msun
    teff = np.clip(5772. * mini ** 0.57, 2800., 15000.)
    logl = 3.5 * np.log10(mini) + 0.3 * np.minimum(u2, 0.7) / 0.7   # main-sequence brightening
    giant = (u2 > 0.7) & (mini > 0.8)
    x = np.where(giant, (u2 - 0.7) / 0.3, 0.)
    teff = np.where(giant, np.minimum(teff, 5200.) - 1500. * x, teff)
    logl = logl + 2.5 * x
    teff = teff * 10. ** (-0.02 * feh)                           # metal-poor stars are hotter
    grid = np.empty((nmodel, nfilt, 3), dtype=np.float32)
    hck = 14387.77  # micron K
    r0 = (0.55 / lam) ** 1.6                                      # A_lambda / A_V at R(V) = 3.3
    dr = 0.06 * (1. - (0.55 / lam) ** 0.8)                        # d(A_lambda/A_V)/dR(V)
    chunk = 1 << 18
    for lo in range(0, nmodel, chunk):
        hi = min(nmodel, lo + chunk)
        t = teff[lo:hi, None]
        bb = -2.5 * np.log10(lam[None, :] ** -5 / np.expm1(hck / (lam[None, :] * t)))
        bbv = -2.5 * np.log10(0.55 ** -5 / np.expm1(hck / (0.55 * t)))
        mv = 4.81 - 2.5 * logl[lo:hi, None]
        blanket = -0.12 * feh[lo:hi, None] * (0.55 / lam[None, :]) ** 2 * (lam[None, :] < 0.7)
        grid[lo:hi, :, 0] = mv + (bb - bbv) + blanket
        grid[lo:hi, :, 1] = (r0 - 3.3 * dr)[None, :]
        grid[lo:hi, :, 2] = dr[None, :]
    labels = np.zeros(nmodel, dtype=[("mini", "f8"), ("eep", "f8"), ("feh", "f8"), ("Mr", "f8"), ("loga", "f8")])
    labels["mini"], labels["eep"], labels["feh"] = mini, 200. + 600. * u2, feh
    labels["Mr"] = grid[:, min(1, nfilt - 1), 0]
    labels["loga"] = np.clip(10.0 - 2.5 * np.log10(mini) + 0.3 * (u2 - 0.5), 6.5, 10.13)   # <= 13.5 Gyr
    return grid, labels


def make_grid_lattice(n_mr=568, n_feh=72, nfilt=8, dist_mod=10.):
    """Bayestar-shaped mock grid (BASELINE.json configs[3]): a regular (Mr, [Fe/H]) lattice, Mr fastest -- the real
    ``grid_bayestar_v5.h5`` has 40 896 models on such a lattice and cannot be fetched offline.  A main sequence
    below Mr = 3.5 and a giant branch above it; blackbody colours + blanketing as in :func:`make_grid_locus`;
    magnitudes at 1 kpc (``dist_mod`` = 10, brutus/seds.py:720), so that scale = parallax^2 holds for real
    parallaxes.  Labels 'Mr', 'feh' only: no 'mini', so the default prior is the PS1 luminosity function
    (brutus/fitting.py:1335-1341)."""
    lam = _WAVE[:nfilt]
    mr = np.repeat(np.linspace(-1., 18., n_mr)[None, :], n_feh, axis=0).ravel()
    feh = np.repeat(np.linspace(-2.5, 0.5, n_feh)[:, None], n_mr, axis=1).ravel()
    nmodel = mr.size
    mini = np.clip(10. ** ((4.81 - mr) / 8.75), 0.08, 10.)
    teff = np.clip(5772. * mini ** 0.57, 2800., 15000.)
    giant = mr < 3.5
    teff = np.where(giant, 5200. - 1500. * (3.5 - mr) / 4.5, teff) * 10. ** (-0.02 * feh)
    hck = 14387.77
    r0 = (0.55 / lam) ** 1.6
    dr = 0.06 * (1. - (0.55 / lam) ** 0.8)
    t = teff[:, None]
    bb = -2.5 * np.log10(lam[None, :] ** -5 / np.expm1(hck / (lam[None, :] * t)))
    bbr = -2.5 * np.log10(lam[1] ** -5 / np.expm1(hck / (lam[1] * t)))
    blanket = -0.12 * feh[:, None] * (0.55 / lam[None, :]) ** 2 * (lam[None, :] < 0.7)
    grid = np.empty((nmodel, nfilt, 3), dtype=np.float32)
    grid[:, :, 0] = mr[:, None] + (bb - bbr) + (blanket - blanket[:, 1:2]) + dist_mod   # band 1 (PS r) carries Mr
    grid[:, :, 1] = (r0 - 3.3 * dr)[None, :]
    grid[:, :, 2] = dr[None, :]
    labels = np.zeros(nmodel, dtype=[("Mr", "f8"), ("feh", "f8")])
    labels["Mr"], labels["feh"] = mr, feh
    return grid, labels


def load_ngc2682(path=None):
    """The NGC 2682 (M67) demo catalogue of the reference (demos/NGC_2682.fits, 1 585 objects) restricted to the 8
    bands a Bayestar-type grid covers, as assembled by tests/gen_golden.py::gen_ngc2682 and
    kept in tests/golden/ngc2682.npz.  Returns the ``make_stars`` dictionary for the objects with >= 4 usable bands
    (``BruteForce`` refuses the others, brutus/fitting.py:1413-1420) plus ``index`` into the catalogue."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ngc2682.npz")
    d = np.load(path)
    ok = np.where(d["mask"].sum(axis=1) >= 4)[0]
    return dict(flux=d["phot"][ok], err=d["err"][ok], mask=d["mask"][ok], parallax=d["parallax"][ok],
                parallax_err=d["parallax_err"][ok], coords=d["coords"][ok], index=ok)


def make_grid(nmodel, nfilt, seed=1000, kind="tilt"):
    """Mock SED grid: absolute magnitude ~U(-2,12), a one-parameter colour family plus 0.03 mag
    scatter, reddening vector falling from 1.2 to 0.2 across the bands, small dR/dRv.
    Also returns labels (structured: 'Mr', 'feh') like ``load_models`` does
    (brutus/utils.py:608-609).  ``kind="locus"`` gives :func:`make_grid_locus` instead."""
    if kind == "locus":
        return make_grid_locus(nmodel, nfilt, seed=seed)
    rs = np.random.RandomState(seed)
    mabs = rs.uniform(-2., 12., nmodel)
    col = rs.normal(0., 0.5, nmodel)
    x = np.linspace(-1., 1., nfilt)
    grid = np.empty((nmodel, nfilt, 3), dtype=np.float32)
    chunk = 1 << 18
    for lo in range(0, nmodel, chunk):  # chunked to bound temporaries at 3M models
        hi = min(nmodel, lo + chunk)
        n = hi - lo
        grid[lo:hi, :, 0] = (mabs[lo:hi, None] + col[lo:hi, None] * x[None, :]
                             + rs.normal(0., 0.03, (n, nfilt)))
        grid[lo:hi, :, 1] = np.linspace(1.2, 0.2, nfilt)[None, :] + rs.normal(0., 0.01, (n, nfilt))
        grid[lo:hi, :, 2] = (np.linspace(0.05, -0.02, nfilt)[None, :]
                             + rs.normal(0., 0.002, (n, nfilt)))
    labels = np.zeros(nmodel, dtype=[("Mr", "f8"), ("feh", "f8")])
    labels["Mr"] = grid[:, min(1, nfilt - 1), 0]
    labels["feh"] = rs.uniform(-2., 0.5, nmodel)
    return grid, labels


def make_stars(grid, nstar, seed=2000, av_max=2.0, dropout=0.0, par_nan_frac=0.3,
               snr_range=(10., 100.)):
    """Mock catalogue drawn from the grid: returns dict(flux, err, mask, parallax, parallax_err,
    coords, truth=(idx, av, rv, dist))."""
    rs = np.random.RandomState(seed)
    nmodel, nfilt, _ = grid.shape
    idx = rs.randint(0, nmodel, nstar)
    av = rs.uniform(0., av_max, nstar)
    rv = np.clip(rs.normal(3.32, 0.18, nstar), 2., 5.)
    dist = 10. ** rs.uniform(-1., 1., nstar)  # kpc
    co = grid[idx].astype(np.float64)
    mag = co[:, :, 0] + av[:, None] * (co[:, :, 1] + rv[:, None] * co[:, :, 2])
    flux = 10. ** (-0.4 * mag) / dist[:, None] ** 2
    snr = rs.uniform(snr_range[0], snr_range[1], (nstar, nfilt))
    err = flux / snr
    flux = flux + rs.normal(0., 1., (nstar, nfilt)) * err
    perr = rs.uniform(0.02, 0.3, nstar)
    par = 1. / dist + rs.normal(0., 1., nstar) * perr
    nan = rs.uniform(size=nstar) < par_nan_frac
    par[nan] = np.nan
    perr[nan] = np.nan
    mask = np.ones((nstar, nfilt), dtype=bool)
    if dropout > 0:
        drop = rs.uniform(size=(nstar, nfilt)) < dropout
        # keep at least 4 bands (brutus/fitting.py:1413-1420)
        for i in np.where((~drop).sum(axis=1) < 4)[0]:
            drop[i] = False
        mask &= ~drop
    # Galactic (l, b) in degrees, from an independent stream so that the photometry above is unchanged
    rc = np.random.RandomState(seed + 77_000)
    coords = np.stack([rc.uniform(0., 360., nstar), np.rad2deg(np.arcsin(rc.uniform(-1., 1., nstar)))], axis=1)
    return dict(flux=flux, err=err, mask=mask, parallax=par, parallax_err=perr, coords=coords,
                truth=dict(idx=idx, av=av, rv=rv, dist=dist))
