"""Host-side mirror of the priors ``BruteForce.fit`` uses (reference: brutus/pdf.py), for the parts of the fit path
that run in NumPy: the static per-model prior staged on the device once, and the host posterior path (user prior
callables, CDF thresholding).  Same names, arguments and return values as the reference.

The default Galactic prior is also built into the device posterior (csrc/posterior.cuh); :func:`gal_lnprior` here is
its NumPy twin.  Its coordinate transform replaces the reference's astropy call (``SkyCoord(...).galactocentric``,
brutus/pdf.py:630-635; astropy >= 4.0 frame defaults: Sun 8.122 kpc from the centre and 20.8 pc above the plane) --
parity of that transform with astropy is unpinned (no astropy offline), everything else is held to golden values of
the reference (tests/test_pdf_host.py).  The 3-D dust prior needs the 2 GB Bayestar map and healpy: not bundled.
"""
import os

import numpy as np
from scipy.special import erf, logsumexp

__all__ = ["imf_lnprior", "ps1_MrLF_lnprior", "parallax_lnprior", "scale_parallax_lnprior", "parallax_to_scale",
           "logn_disk", "logn_halo", "logp_feh", "logp_age_from_feh", "gal_lnprior", "dust_lnprior"]

GALCEN_DISTANCE = 8.122   # kpc
Z_SUN = 0.0208            # kpc


def _imf_piece(m, alpha_low, alpha_high, mass_break):
    m = np.asarray(m, dtype=float)
    out = np.full_like(m, -np.inf)
    lo = (m > 0.08) & (m <= mass_break)
    hi = m > mass_break
    out[lo] = -alpha_low * np.log(m[lo])
    out[hi] = -alpha_high * np.log(m[hi]) + (alpha_high - alpha_low) * np.log(mass_break)
    return out


def imf_lnprior(mgrid, alpha_low=1.3, alpha_high=2.3, mass_break=0.5, mgrid2=None):
    """Kroupa-like broken power-law IMF over initial mass, optionally times the same for a binary companion
    (brutus/pdf.py:38-108)."""
    lnprior = _imf_piece(mgrid, alpha_low, alpha_high, mass_break)
    n_lo = mass_break ** (1. - alpha_low) / (alpha_high - 1.)
    n_hi = (0.08 ** (1. - alpha_low) - mass_break ** (1. - alpha_low)) / (alpha_low - 1.)   # from the H-burning limit
    norm = n_lo + n_hi
    if mgrid2 is not None:
        lnprior = lnprior + _imf_piece(mgrid2, alpha_low, alpha_high, mass_break)
        norm = n_lo ** 2 + n_hi ** 2 + 2. * n_lo * n_hi
    return lnprior - np.log(norm)


_ps1_table = None


def ps1_MrLF_lnprior(Mr):
    """PanSTARRS r-band luminosity-function prior over absolute magnitude: linear interpolation of the tabulated
    ln(prior), linearly extrapolated beyond the table like ``interp1d(fill_value='extrapolate')``
    (brutus/pdf.py:111-141; table: brutus_b200/data/ps1_mr_lf.npz, built from the reference's
    PSMrLF_lnprior.dat by tools/make_ps1_table.py)."""
    global _ps1_table
    if _ps1_table is None:
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ps1_mr_lf.npz"))
        _ps1_table = (np.asarray(d["Mr"], dtype=float), np.asarray(d["lnprior"], dtype=float))
    x, y = _ps1_table
    Mr = np.asarray(Mr, dtype=float)
    k = np.clip(np.searchsorted(x, Mr, side="right") - 1, 0, len(x) - 2)
    return y[k] + (Mr - x[k]) * (y[k + 1] - y[k]) / (x[k + 1] - x[k])


def parallax_lnprior(parallaxes, p_meas, p_err):
    """Gaussian parallax likelihood, flat if there is no measurement (brutus/pdf.py:144-175)."""
    if np.isfinite(p_meas) and np.isfinite(p_err):
        return -0.5 * ((parallaxes - p_meas) ** 2 / p_err ** 2 + np.log(2. * np.pi * p_err ** 2))
    return np.zeros_like(parallaxes)


def parallax_to_scale(p_meas, p_err, snr_lim=4.):
    """Mean and standard deviation of scale = parallax**2 (brutus/pdf.py:225-260)."""
    if p_meas / p_err <= snr_lim:
        return np.nan, np.nan
    pm = max(0., p_meas)
    return pm ** 2 + p_err ** 2, np.sqrt(2 * p_err ** 4 + 4 * pm ** 2 * p_err ** 2)


def scale_parallax_lnprior(scales, scale_errs, p_meas, p_err, snr_lim=4.):
    """Parallax prior mapped to scale = parallax**2 (brutus/pdf.py:178-222)."""
    if np.isfinite(p_meas) and np.isfinite(p_err) and p_meas / p_err > snr_lim:
        s_mean, s_std = parallax_to_scale(p_meas, p_err, snr_lim=snr_lim)
        vtot = s_std ** 2 + scale_errs ** 2
        return -0.5 * ((scales - s_mean) ** 2 / vtot + np.log(2. * np.pi * vtot))
    return np.zeros_like(scales)


def _galactic_to_cyl(dists, coord):
    """Galactic (l, b) [deg] + distance [kpc] -> Galactocentric cylindrical (R, Z) [kpc]: shift the heliocentric
    vector by the Sun-centre distance, then tilt about y so that the Sun sits Z_SUN above the plane."""
    ell, b = np.deg2rad(coord[0]), np.deg2rad(coord[1])
    d = np.asarray(dists, dtype=np.float64)
    x = d * np.cos(b) * np.cos(ell) - GALCEN_DISTANCE
    y = d * np.cos(b) * np.sin(ell)
    z = d * np.sin(b)
    st = Z_SUN / GALCEN_DISTANCE
    ct = np.sqrt(1. - st * st)
    return np.hypot(x * ct + z * st, y), z * ct - x * st


def logn_disk(R, Z, R_solar=8.2, Z_solar=0.025, R_scale=2.6, Z_scale=0.3, R_smooth=2.):
    """ln number density of a double-exponential disk relative to the solar neighbourhood (brutus/pdf.py:263-307)."""
    Reff = np.sqrt(R ** 2 + R_smooth ** 2)
    return -((Reff - R_solar) / R_scale + (np.abs(Z) - np.abs(Z_solar)) / Z_scale)


def logn_halo(R, Z, R_solar=8.2, Z_solar=0.025, R_smooth=2., eta=4.2, q_ctr=0.2, q_inf=0.8, r_q=6.):
    """ln number density of a power-law halo with radius-dependent oblateness (brutus/pdf.py:310-377)."""
    def reff(R_, Z_):
        rp = np.sqrt(R_ ** 2 + Z_ ** 2 + r_q ** 2)
        q = q_inf - (q_inf - q_ctr) * np.exp(1. - rp / r_q)
        return np.sqrt(R_ ** 2 + (Z_ / q) ** 2 + R_smooth ** 2)
    return -eta * np.log(reff(R, Z) / reff(R_solar, Z_solar))


def logp_feh(feh, feh_mean=-0.2, feh_sigma=0.3):
    """Gaussian ln prior over [Fe/H] (brutus/pdf.py:380-408)."""
    return -0.5 * ((feh_mean - feh) ** 2 / feh_sigma ** 2 + np.log(2. * np.pi * feh_sigma ** 2))


def logp_age_from_feh(age, feh_mean=-0.2, max_age=13.8, min_age=0., feh_age_ctr=-0.5, feh_age_scale=0.5,
                      nsigma_from_max_age=2., max_sigma=4., min_sigma=1.):
    """Truncated-normal ln prior over age [Gyr] whose mean follows the component's metallicity
    (brutus/pdf.py:411-473, brutus/utils.py:232-284)."""
    mean = (max_age - min_age) / (1. + np.exp((feh_mean - feh_age_ctr) / feh_age_scale)) + min_age
    sig = min(max((max_age - mean) / nsigma_from_max_age, min_sigma), max_sigma)
    a, b = (min_age - mean) / sig, (max_age - mean) / sig
    xi = (age - mean) / sig
    lnden = np.log(sig / 2.) + np.log(erf(b / np.sqrt(2.)) - erf(a / np.sqrt(2.)))
    out = -0.5 * np.log(2. * np.pi) - 0.5 * xi ** 2 - lnden
    return np.where((age < min_age) | (age > max_age), -np.inf, out)


def gal_lnprior(dists, coord, labels=None, R_solar=8.2, Z_solar=0.025, R_thin=2.6, Z_thin=0.3, Rs_thin=2.,
                R_thick=2.0, Z_thick=0.9, f_thick=0.04, Rs_thick=2., Rs_halo=2., q_halo_ctr=0.2, q_halo_inf=0.8,
                r_q_halo=6.0, eta_halo=4.2, f_halo=0.005, feh_thin=-0.2, feh_thin_sigma=0.3, feh_thick=-0.7,
                feh_thick_sigma=0.4, feh_halo=-1.6, feh_halo_sigma=0.5, max_age=13.8, min_age=0., feh_age_ctr=-0.5,
                feh_age_scale=0.5, nsigma_from_max_age=2., max_sigma=4., min_sigma=1., return_components=False):
    """ln prior of a thin disk + thick disk + halo Galactic model in distance, and in metallicity / age when
    ``labels`` carries 'feh' / 'loga' (brutus/pdf.py:476-749)."""
    dists = np.asarray(dists, dtype=np.float64)
    age_kw = dict(max_age=max_age, min_age=min_age, feh_age_ctr=feh_age_ctr, feh_age_scale=feh_age_scale,
                  nsigma_from_max_age=nsigma_from_max_age, max_sigma=max_sigma, min_sigma=min_sigma)
    with np.errstate(all="ignore"):
        vol = 2. * np.log(dists + 1e-300)                                   # dV ~ d^2
        R, Z = _galactic_to_cyl(dists, coord)
        comp = [logn_disk(R, Z, R_solar, Z_solar, R_thin, Z_thin, Rs_thin) + vol,
                logn_disk(R, Z, R_solar, Z_solar, R_thick, Z_thick, Rs_thick) + vol + np.log(f_thick),
                logn_halo(R, Z, R_solar, Z_solar, Rs_halo, eta_halo, q_halo_ctr, q_halo_inf, r_q_halo) + vol
                + np.log(f_halo)]
        lnprior = logsumexp(comp, axis=0)
        out_comp = {"number_density": comp}
        names = () if labels is None else (labels.dtype.names or ())
        if labels is not None:
            member = [c - lnprior for c in comp]                            # ln membership probabilities
            feh_mu, feh_sig = (feh_thin, feh_thick, feh_halo), (feh_thin_sigma, feh_thick_sigma, feh_halo_sigma)
            if "feh" in names:
                terms = [logp_feh(labels["feh"], m, s) + w for m, s, w in zip(feh_mu, feh_sig, member)]
                lnprior = lnprior + logsumexp(terms, axis=0)
                out_comp["feh"] = terms
            if "loga" in names:
                age = 10. ** labels["loga"] / 1e9
                terms = [logp_age_from_feh(age, feh_mean=m, **age_kw) + w for m, w in zip(feh_mu, member)]
                lnprior = lnprior + logsumexp(terms, axis=0)
                out_comp["age"] = terms
    if return_components:
        return lnprior, out_comp
    return lnprior


def dust_lnprior(dists, coord, avs, dustfile="bayestar2019_v1.h5", **kwargs):
    """The reference's 3-D dust prior (brutus/pdf.py:752-840) interpolates the Bayestar19 map (2 GB HDF5, HEALPix
    look-ups through healpy): neither is available to this build.  Pass ``lndustprior=<callable>`` to ``fit`` or
    ``dustfile=None`` for the flat A(V) prior (brutus/fitting.py:1396-1398)."""
    raise NotImplementedError("the Bayestar dust map is not bundled; pass `lndustprior` or `dustfile=None`")
