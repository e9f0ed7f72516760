"""Host-side mirror of the reference's fit path (brutus/fitting.py) over the CUDA sweep.

Public names follow the reference: :func:`loglike` (brutus/fitting.py:579), :func:`lnpost` (:823)
and :class:`BruteForce` (:1110) with ``fit`` / ``_fit``.  All O(Nmodel) arithmetic runs in
``libbrutus_b200.so``; this module only validates arguments, raises the reference's
``ValueError``s, and post-processes the (small) selected subsets.  No CPU fallback exists.
"""
import weakref

import numpy as np

from . import _lib

__all__ = ["loglike", "get_handle", "release_handles"]

_handles = {}


def get_handle(mag_coeffs, precision="f32", device=0):
    """Return a sweep handle with ``mag_coeffs`` staged in HBM, uploading only when the array
    object changes (the reference re-reads the host array on every call, brutus/fitting.py:714)."""
    key = (precision, int(device))
    ent = _handles.get(key)
    if ent is None:
        ent = {"h": _lib.Handle(device, precision), "ref": None, "shape": None}
        _handles[key] = ent
    ref = ent["ref"]() if ent["ref"] is not None else None
    if ref is not mag_coeffs or ent["shape"] != mag_coeffs.shape:
        ent["h"].set_grid(mag_coeffs)
        try:
            ent["ref"] = weakref.ref(mag_coeffs)
        except TypeError:
            ent["ref"] = None
        ent["shape"] = mag_coeffs.shape
    return ent["h"]


def release_handles():
    for ent in _handles.values():
        ent["h"].close()
    _handles.clear()


def loglike(data, data_err, data_mask, mag_coeffs,
            avlim=(0., 20.), av_gauss=(0., 1e6),
            rvlim=(1., 8.), rv_gauss=(3.32, 0.18),
            av_init=None, rv_init=None,
            dim_prior=True, ltol=3e-2, ltol_subthresh=1e-2, init_thresh=5e-3,
            parallax=None, parallax_err=None,
            return_vals=False, precision="f32", device=0, return_diag=False, *args, **kwargs):
    """Drop-in for the reference's ``loglike`` (brutus/fitting.py:579-820): same arguments, same
    return tuple ``(lnl, Ndim, chi2[, scale, av, rv, icov_sar])`` (float64 arrays of length
    Nmodel), same in-place clean-up of ``data_mask`` (:709) and the same ``ValueError`` (:691-693).

    ``av_init``/``rv_init`` other than the defaults (the prior means, :700-703) are not supported
    by the kernels.  ``precision`` selects float32 (throughput) or float64 (verification) math.
    """
    if init_thresh is None:
        raise NotImplementedError("init_thresh=None (no cull) is not supported")
    if init_thresh > ltol_subthresh:
        raise ValueError("The initial threshold must be smaller than or equal "
                         "to the final threshold applied to be useful!")
    if av_init is not None or rv_init is not None:
        raise NotImplementedError("av_init/rv_init must be None (prior means are used)")
    h = get_handle(mag_coeffs, precision=precision, device=device)
    opts = _lib.make_options(avlim=avlim, av_gauss=av_gauss, rvlim=rvlim, rv_gauss=rv_gauss,
                             dim_prior=dim_prior, ltol=ltol, ltol_subthresh=ltol_subthresh,
                             init_thresh=init_thresh)
    par = np.nan if parallax is None or parallax_err is None else float(parallax)
    perr = np.nan if parallax is None or parallax_err is None else float(parallax_err)
    lnl, chi2, sc, av, rv, icov, mclean, diag = h.loglike_full(
        data, data_err, data_mask, par, perr, opts, want_icov=return_vals)
    data_mask[...] = mclean  # brutus/fitting.py:709 mutates the caller's mask
    out = (lnl, int(diag[0]), chi2)
    if return_vals:
        out = out + (sc, av, rv, icov)
    if return_diag:
        out = out + ({"n_iter_mag": int(diag[1]), "n_iter_flux": int(diag[2]),
                      "n_surv": int(diag[3]), "stats": h.stats()},)
    return out


# =====================================================================================================
# Host-side consumers of the sweep ("next" rows, kept on the host for now).
# They only ever touch the compacted selection the GPU returns, never O(Nmodel) arrays.
# =====================================================================================================
import sys
import time
import warnings

try:
    from scipy.special import logsumexp
except ImportError:  # pragma: no cover
    from scipy.misc import logsumexp

__all__ += ["lnpost_selected", "BruteForce", "imf_lnprior", "parallax_lnprior",
            "scale_parallax_lnprior"]


def imf_lnprior(mgrid, alpha_low=1.3, alpha_high=2.3, mass_break=0.5):
    """Kroupa-like broken power-law IMF prior over initial mass (brutus/pdf.py:38-108, single
    stars)."""
    mgrid = np.asarray(mgrid, dtype=float)
    out = np.full_like(mgrid, -np.inf)
    lo = (mgrid > 0.08) & (mgrid <= mass_break)
    hi = mgrid > mass_break
    out[lo] = -alpha_low * np.log(mgrid[lo])
    out[hi] = -alpha_high * np.log(mgrid[hi]) + (alpha_high - alpha_low) * np.log(mass_break)
    n_lo = mass_break ** (1. - alpha_low) / (alpha_high - 1.)
    n_hi = (0.08 ** (1. - alpha_low) - mass_break ** (1. - alpha_low)) / (alpha_low - 1.)
    return out - np.log(n_lo + n_hi)


def parallax_lnprior(parallaxes, p_meas, p_err):
    """Gaussian parallax likelihood, flat if there is no measurement (brutus/pdf.py:144-175)."""
    if np.isfinite(p_meas) and np.isfinite(p_err):
        return -0.5 * ((parallaxes - p_meas) ** 2 / p_err ** 2 + np.log(2. * np.pi * p_err ** 2))
    return np.zeros_like(parallaxes)


def scale_parallax_lnprior(scales, scale_errs, p_meas, p_err, snr_lim=4.):
    """Parallax prior mapped to scale = parallax**2 (brutus/pdf.py:178-260)."""
    if np.isfinite(p_meas) and np.isfinite(p_err) and p_meas / p_err > snr_lim:
        pm = max(0., p_meas)
        s_mean = pm ** 2 + p_err ** 2
        s_var = 2 * p_err ** 4 + 4 * pm ** 2 * p_err ** 2
        vtot = s_var + scale_errs ** 2
        return -0.5 * ((scales - s_mean) ** 2 / vtot + np.log(2. * np.pi * vtot))
    return np.zeros_like(scales)


def _inv3(mats):
    """Batched inverse of 3x3 matrices through the adjugate (brutus/utils.py:71-114)."""
    adj = np.empty_like(mats)
    for i in range(3):
        adj[..., i, :] = np.cross(mats[..., i - 2, :], mats[..., i - 1, :])
    det = np.einsum("...i,...i->...", adj, mats).mean(axis=-1)
    return np.swapaxes(adj / det[..., None, None], -1, -2)


def _mvn_draws(mean, cov, size, rstate, eps=1e-30):
    """Nmc draws from each of N trivariate normals; returns (3, size, N) like
    brutus/utils.py:845-905 and consumes the generator identically."""
    n, d = mean.shape
    chol = np.linalg.cholesky(cov + eps * np.identity(d)[None, :, :])
    z = rstate.normal(loc=0, scale=1, size=d * size * n).reshape(n, d, size)
    draws = mean[:, :, None] + np.matmul(chol, z)
    return np.transpose(draws, (1, 2, 0))


def _unpack_icov(icov6):
    """(6, n) rows (ss, sa, sr, aa, ar, rr) -> (n, 3, 3) symmetric matrices."""
    ss, sa, sr, aa, ar, rr = [np.asarray(x, dtype=np.float64) for x in icov6]
    out = np.empty((len(ss), 3, 3))
    out[:, 0, 0], out[:, 1, 1], out[:, 2, 2] = ss, aa, rr
    out[:, 0, 1] = out[:, 1, 0] = sa
    out[:, 0, 2] = out[:, 2, 0] = sr
    out[:, 1, 2] = out[:, 2, 1] = ar
    return out


def lnpost_selected(sel, lnlike, scales, avs, rvs, icovs_sar, parallax=None, parallax_err=None,
                    coord=None, Nmc_prior=100, lnprior=None, wt_thresh=1e-3, lngalprior=None,
                    lndustprior=None, dustfile=None, dlabels=None, avlim=(0., 20.), rvlim=(1., 8.),
                    mem_lim=8000., rstate=None, apply_av_prior=True):
    """The part of the reference's ``lnpost`` (brutus/fitting.py:999-1107) that follows the first
    selection, which ``bf_sweep_batch`` already performed on the GPU (:976-991).  Inputs are the
    compacted records of the first selection ``sel`` (model indices, ascending).  Returns the
    reference's tuple ``(sel, cov_sar, lnp, dist_mc, a_mc, r_mc, lnp_mc)``."""
    if rstate is None:
        rstate = np.random
    if lngalprior is None:
        raise NotImplementedError("the default Galactic prior (brutus/pdf.py:476-749, astropy) is outside "
                                  "this build's scope; pass `lngalprior`")
    if lndustprior is None and apply_av_prior:
        raise NotImplementedError("the default 3-D dust prior needs the Bayestar map; pass `lndustprior`")
    if coord is None:
        coord = np.zeros(2)
    if lnprior is None:
        lnprior = np.zeros(1)
    lnprior_sel = lnprior[sel] if np.ndim(lnprior) else np.full(len(sel), float(lnprior))
    nsel_max = int(mem_lim / Nmc_prior / 4.0e-4) if Nmc_prior > 0 else len(sel)  # :969-970
    have_par = parallax is not None and parallax_err is not None
    lnlike = np.asarray(lnlike, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    avs = np.asarray(avs, dtype=np.float64)
    rvs = np.asarray(rvs, dtype=np.float64)
    first = np.arange(len(sel))
    # MLE-based prior evaluation and second threshold (:1000-1016)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lnp = lnlike + lnprior_sel
        dist = 1. / np.sqrt(scales)
        lnp = lnp + lngalprior(dist, coord, labels=None if dlabels is None else dlabels[sel])
        if apply_av_prior:
            lnp = lnp + lndustprior(dist, coord, avs, dustfile=dustfile)
    keep = first[lnp > np.log(wt_thresh) + np.max(lnp)]
    lnp = lnlike[keep] + lnprior_sel[keep]
    if len(keep) > nsel_max:  # :1029-1036
        order = np.argsort(lnp)[::-1][:nsel_max]
        keep, lnp = keep[order], lnp[order]
    scale, av, rv = scales[keep], avs[keep], rvs[keep]
    icov = np.array(icovs_sar[keep], dtype=np.float64)
    sel_out = np.asarray(sel)[keep]
    nsel = len(keep)
    # covariances, regularised until positive definite (:1039-1065)
    cov = _inv3(icov)
    bad = np.where(~np.all(np.linalg.eigvals(cov) > 0, axis=1))[0]
    width, count = 0.02, 1
    while len(bad) > 0:
        sfr = scale[bad] * width
        n1, n2, n3 = (cov[bad][:, k, k] <= 0 for k in range(3))
        m1 = n1 + (~n2 * ~n3)
        m2 = n2 + (~n1 * ~n3)
        m3 = n3 + (~n1 * ~n2)
        add = np.zeros((len(bad), 3, 3))
        add[:, 0, 0] = count / sfr ** 2 * m1
        add[:, 1, 1] = count / width ** 2 * m2
        add[:, 2, 2] = count / width ** 2 * m3
        icov[bad] += add
        cov[bad] = _inv3(icov[bad])
        bad = bad[np.where(~np.all(np.linalg.eigvals(cov[bad]) > 0, axis=1))[0]]
        count *= 2
    # Monte-Carlo integration over the priors (:1068-1101)
    if Nmc_prior > 0:
        s_mc, a_mc, r_mc = _mvn_draws(np.transpose([scale, av, rv]), cov, Nmc_prior, rstate)
        lab_mc = None
        if dlabels is not None:
            lab_mc = np.tile(dlabels[sel_out], Nmc_prior).reshape(-1, nsel)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            par_mc = np.sqrt(s_mc)
            dist_mc = 1. / par_mc
            lnp_mc = lngalprior(dist_mc, coord, labels=lab_mc)
            if apply_av_prior:
                lnp_mc = lnp_mc + lndustprior(dist_mc, coord, a_mc, dustfile=dustfile)
        if have_par:
            lnp_mc = lnp_mc + parallax_lnprior(par_mc, parallax, parallax_err)
        inb = ((s_mc >= 1e-20) & (a_mc >= avlim[0]) & (a_mc <= avlim[1])
               & (r_mc >= rvlim[0]) & (r_mc <= rvlim[1]))
        lnp_mc = np.where(inb, lnp_mc, -1e300)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lnp = lnp + logsumexp(lnp_mc, axis=0) - np.log(np.sum(inb, axis=0))
    else:
        lnp = lnlike[keep]
        if have_par:
            lnp = lnp + scale_parallax_lnprior(scale, 1. / np.sqrt(np.abs(icov[:, 0, 0])), parallax, parallax_err)
        dist_mc = a_mc = r_mc = lnp_mc = np.zeros((0, nsel))
    lnp = np.where(np.isfinite(lnp), lnp, -1e300)
    return sel_out, keep, cov, lnp, dist_mc.T, a_mc.T, r_mc.T, lnp_mc.T


class BruteForce(object):
    """Same constructor, ``fit`` and ``_fit`` as the reference's ``BruteForce``
    (brutus/fitting.py:1110-2065), with the per-star full-grid sweep on the GPU.

    Two posterior paths.  With ``lngalprior=None`` (the reference's default, ``gal_lnprior``) the whole
    per-object body of ``_fit`` runs on the device (``bf_fit_batch``): sweep, ``lnpost`` with the built-in
    Galactic prior (brutus/pdf.py:476-749), evidence and resampling; only ``Ndraws`` samples per object
    cross PCIe and the random numbers come from a counter-based generator seeded from ``rstate``.  With
    a user ``lngalprior`` callable the selected models are shipped to the host and ``lnpost`` runs in NumPy
    with the caller's ``rstate`` exactly as in the reference.

    Differences, all outside the hot path: the 3-D dust prior is not bundled (no Bayestar map offline;
    pass ``lndustprior`` on the host path; with ``dustfile=None`` and no ``lndustprior`` the A(V) prior is
    flat, as in the reference :1396-1398); ``parallax=None`` is accepted (the reference raises TypeError);
    results go to ``<save_file>.h5`` when h5py is importable and
    to ``<save_file>.npz`` with the same dataset names otherwise."""

    def __init__(self, models, models_labels, labels_mask, precision="f32", device=0):
        self.NMODEL, self.NDIM, self.NCOEF = models.shape
        self.models = models
        self.models_labels = models_labels
        self.labels_mask = labels_mask
        self.NLABELS = len(models_labels.dtype.names) if models_labels.dtype.names else 0
        self.precision = precision
        self.device = device
        self._handle = None

    # -- device handle, created on first use so that constructing the object needs no GPU --
    def _get_handle(self):
        if self._handle is None:
            h = _lib.Handle(self.device, self.precision)
            h.set_grid(self.models)
            self._handle = h
        return self._handle

    # test hook: host-supplied normals / uniforms for the device posterior (tests/test_posterior_gpu.py)
    _z_override = None
    _u_override = None

    def _post_test_hooks(self, b0, b1):
        kw = {}
        if self._z_override is not None:
            kw["z_override"] = self._z_override
        if self._u_override is not None:
            kw["u_override"] = self._u_override[b0:b1]
        return kw

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def _setup(self, data, data_err, data_mask, data_labels=None, phot_offsets=None, parallax=None,
               parallax_err=None, av_gauss=None, lnprior=None, wt_thresh=1e-3, cdf_thresh=2e-3,
               apply_agewt=True, apply_grad=True, lngalprior=None, lndustprior=None, dustfile=None,
               data_coords=None, ltol_subthresh=1e-2, logl_initthresh=5e-3, mag_max=50.,
               merr_max=0.25, rstate=None):
        """Argument checking and data cleaning of brutus/fitting.py:1144-1424."""
        data = np.array(data, dtype=np.float64)
        data_err = np.array(data_err, dtype=np.float64)
        data_mask = np.array(data_mask, dtype=bool)
        ndata, nfilt = data.shape
        if logl_initthresh > ltol_subthresh:
            raise ValueError("The initial threshold must be smaller than or equal to the "
                             "convergence threshold in order to be useful!")
        if wt_thresh is None and cdf_thresh is None:
            wt_thresh = -np.inf
        if wt_thresh is None:
            raise NotImplementedError("CDF-based thresholding (wt_thresh=None) is not supported; "
                                      "the GPU selection uses wt_thresh (brutus/fitting.py:988-991)")
        if rstate is None:
            rstate = np.random
        if parallax is not None and parallax_err is None:
            raise ValueError("Must provide both `parallax` and `parallax_err`.")
        if phot_offsets is None:
            phot_offsets = np.ones(nfilt)
        if lnprior is None:
            names = self.models_labels.dtype.names or ()
            if "mini" in names:
                lnprior = imf_lnprior(self.models_labels["mini"])
            else:
                raise NotImplementedError("default prior without a 'mini' label needs the PS1 luminosity "
                                          "function table (brutus/pdf.py:111-141); pass `lnprior`")
        lnprior = np.array(lnprior, dtype=np.float64)
        names = self.models_labels.dtype.names or ()
        if apply_agewt and "agewt" in names:
            lnprior = lnprior + np.log(np.abs(self.models_labels["agewt"]))
        if apply_grad:
            for l in names:
                if self.labels_mask[l][0]:
                    ul = np.unique(self.models_labels[l])
                    if len(ul) > 1:
                        lnprior = lnprior + np.interp(self.models_labels[l], ul, np.log(np.gradient(ul)))
        if lngalprior is None and data_coords is None:   # brutus/fitting.py:1362-1365
            raise ValueError("`data_coords` must be provided if using the default Galactic model prior.")
        if lndustprior is None and dustfile is not None:
            raise NotImplementedError("pass `lndustprior`: the Bayestar dust prior is not bundled")
        if lndustprior is None and av_gauss is None:
            av_gauss = (0, 1e6)  # flat A(V) prior, brutus/fitting.py:1396-1398
        if data_coords is None:
            data_coords = np.zeros((ndata, 2))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mag = -2.5 * np.log10(data)
            merr = 2.5 / np.log(10.) * data_err / data
            bad = (mag > mag_max) | (merr > merr_max)
            clean = np.isfinite(data) & np.isfinite(data_err) & (data_err > 0.)
            data_mask = data_mask & clean & ~bad
        if np.any(np.sum(data_mask, axis=1) < 4):
            raise ValueError("Objects with fewer than 4 bands of acceptable photometry are currently "
                             "included in the dataset. These objects give degenerate fits and cannot be "
                             "properly modeled. Please remove these objects or modify `mag_max` or "
                             "`merr_max`.")
        return (data * phot_offsets, data_err * phot_offsets, data_mask, data_labels, data_coords,
                lnprior, lngalprior, lndustprior, av_gauss, wt_thresh, rstate)

    def _fit(self, data, data_err, data_mask, parallax=None, parallax_err=None, Nmc_prior=100,
             avlim=(0., 20.), av_gauss=None, rvlim=(1., 8.), rv_gauss=(3.32, 0.18), lnprior=None,
             lnprior_ext=None, wt_thresh=1e-3, cdf_thresh=2e-3, Ndraws=250, lngalprior=None,
             lndustprior=None, dustfile=None, apply_dlabels=True, data_coords=None,
             return_distreds=True, logl_dim_prior=True, ltol=3e-2, ltol_subthresh=1e-2,
             logl_initthresh=5e-3, mem_lim=8000., rstate=None, batch=1024, _device_arrays=False):
        """Generator with the reference's contract (brutus/fitting.py:1803-2061): yields, per object,
        ``(sidxs, scales, avs, rvs, cov_sar, Ndim, lnprob, levid, chi2min[, dists, reds, dreds,
        logwts])``.  Stars go to the GPU ``batch`` at a time; the prior integration and resampling
        run on the host in catalogue order with the caller's ``rstate``."""
        (data, data_err, data_mask, _, data_coords, lnprior, lngalprior, lndustprior, av_gauss,
         wt_thresh, rstate) = self._setup(
            data, data_err, data_mask, parallax=parallax, parallax_err=parallax_err,
            av_gauss=av_gauss, lnprior=lnprior, wt_thresh=wt_thresh, cdf_thresh=cdf_thresh,
            apply_agewt=False, apply_grad=False, lngalprior=lngalprior, lndustprior=lndustprior,
            dustfile=dustfile, data_coords=data_coords, ltol_subthresh=ltol_subthresh,
            logl_initthresh=logl_initthresh, mag_max=np.inf, merr_max=np.inf, rstate=rstate)
        apply_av_prior = av_gauss is None  # brutus/fitting.py:1965
        ndata = data.shape[0]
        if parallax is None:  # the reference indexes parallax[i] unconditionally (:1989)
            parallax = np.full(ndata, np.nan)
            parallax_err = np.full(ndata, np.nan)
        parallax = np.asarray(parallax, dtype=np.float64)
        parallax_err = np.asarray(parallax_err, dtype=np.float64)
        dlabels = self.models_labels if apply_dlabels else None
        device_posterior = lngalprior is None
        if device_posterior and lndustprior is not None:
            raise NotImplementedError("a user `lndustprior` needs a user `lngalprior` too (host posterior path)")
        if device_posterior and Nmc_prior < 1:
            raise NotImplementedError("Nmc_prior = 0 is only supported on the host posterior path")
        ext_keys = []
        if lnprior_ext is not None:   # validated before any device work, like every other argument error
            ext_keys = list(lnprior_ext.keys())
            for k in ext_keys:
                if k not in self.models_labels.dtype.names:
                    raise ValueError("Provided `lnprior_ext` has keys which do not match the "
                                     "underlying model labels.")
        h = self._get_handle()
        if ext_keys:
            h.set_labels(np.stack([np.asarray(self.models_labels[k], dtype=np.float64) for k in ext_keys]))
        elif h.nlabel:
            h.set_labels(np.zeros((0, self.NMODEL)))
        opts = _lib.make_options(avlim=avlim, av_gauss=av_gauss, rvlim=rvlim, rv_gauss=rv_gauss,
                                 dim_prior=logl_dim_prior, ltol=ltol, ltol_subthresh=ltol_subthresh,
                                 init_thresh=logl_initthresh, wt_thresh=wt_thresh)
        if device_posterior:
            names = (dlabels.dtype.names or ()) if dlabels is not None else ()
            h.set_model_priors(lnprior=lnprior if np.ndim(lnprior) else np.full(self.NMODEL, float(lnprior)),
                               feh=dlabels["feh"] if "feh" in names else None,
                               loga=dlabels["loga"] if "loga" in names else None)
            seed = int(rstate.randint(0, 2 ** 31 - 1)) if hasattr(rstate, "randint") else 0
        for b0 in range(0, ndata, batch):
            b1 = min(ndata, b0 + batch)
            em = es = None
            if ext_keys:
                ext = np.array([[lnprior_ext[k][i] for k in ext_keys] for i in range(b0, b1)], dtype=np.float64)
                em, es = ext[:, :, 0], ext[:, :, 1]
            if device_posterior:
                # star_base keeps the generator keyed by the catalogue index of every star
                r = h.fit_batch(data[b0:b1], data_err[b0:b1], data_mask[b0:b1], parallax[b0:b1],
                                parallax_err[b0:b1], coords=data_coords[b0:b1], ext_mean=em, ext_std=es,
                                opts=opts, nmc_prior=Nmc_prior, ndraws=Ndraws,
                                seed=seed, star_base=b0, mem_lim=mem_lim, copy=not _device_arrays,
                                **(self._post_test_hooks(b0, b1)))
                if _device_arrays:   # fit(): whole-batch arrays, no per-object Python loop
                    yield b0, b1, r
                    continue
                for k in range(b1 - b0):
                    out = (r["sidxs"][k].astype(np.int64), r["scales"][k], r["avs"][k], r["rvs"][k],
                           r["cov_sar"][k], int(r["ndim"][k]), r["lnprob"][k], float(r["levid"][k]),
                           float(r["chi2min"][k]))
                    if return_distreds:
                        out = out + (r["dists"][k], r["reds"][k], r["dreds"][k], r["logwts"][k])
                    yield out
                continue
            res = h.sweep_batch(data[b0:b1], data_err[b0:b1], data_mask[b0:b1], parallax[b0:b1],
                                parallax_err[b0:b1], ext_mean=em, ext_std=es, opts=opts,
                                rows=_lib.REC_FULL, copy=True)
            for i in range(b0, b1):
                lo, hi = res["offsets"][i - b0], res["offsets"][i - b0 + 1]
                sel = res["model_idx"][lo:hi]
                chi2 = np.array(res["chi2"][lo:hi], dtype=np.float64)
                scales = np.array(res["scale"][lo:hi], dtype=np.float64)
                avs = np.array(res["av"][lo:hi], dtype=np.float64)
                rvs = np.array(res["rv"][lo:hi], dtype=np.float64)
                ndim = int(res["ndim"][i - b0])
                sel2, keep, cov_sar, lnprob, dists, reds, dreds, logwts = lnpost_selected(
                    sel, res["lnl"][lo:hi], scales, avs, rvs, _unpack_icov(res["icov6"][:, lo:hi]),
                    parallax=parallax[i], parallax_err=parallax_err[i], coord=data_coords[i],
                    Nmc_prior=Nmc_prior, lnprior=lnprior, wt_thresh=wt_thresh, lngalprior=lngalprior,
                    lndustprior=lndustprior, dustfile=dustfile, dlabels=dlabels, avlim=avlim,
                    rvlim=rvlim, mem_lim=mem_lim, rstate=rstate, apply_av_prior=apply_av_prior)
                # parallax enters chi2 and Ndim (:2025-2030)
                chi2k = chi2[keep]
                if np.isfinite(parallax[i]) and np.isfinite(parallax_err[i]):
                    chi2k = chi2k + (np.sqrt(scales[keep]) - parallax[i]) ** 2 / parallax_err[i] ** 2
                    ndim += 1
                # evidence and resampling (:2032-2061)
                levid = logsumexp(lnprob)
                chi2min = np.min(chi2k)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    wt = np.exp(lnprob - levid)
                    wt /= wt.sum()
                    draws = rstate.choice(len(sel2), size=Ndraws, p=wt)
                sidxs = sel2[draws]
                out = (sidxs, scales[keep][draws], avs[keep][draws], rvs[keep][draws], cov_sar[draws],
                       ndim, lnprob[draws], levid, chi2min)
                if return_distreds:
                    imc = np.zeros(Ndraws, dtype="int")
                    for j, d in enumerate(draws):
                        w = np.exp(logwts[d] - logsumexp(logwts[d]))
                        w /= w.sum()
                        imc[j] = rstate.choice(Nmc_prior, p=w)
                    out = out + (dists[draws, imc], reds[draws, imc], dreds[draws, imc], logwts[draws, imc])
                yield out

    def fit(self, data, data_err, data_mask, data_labels, save_file, phot_offsets=None, parallax=None,
            parallax_err=None, Nmc_prior=50, avlim=(0., 20.), av_gauss=None, rvlim=(1., 8.),
            rv_gauss=(3.32, 0.18), lnprior=None, lnprior_ext=None, wt_thresh=1e-3, cdf_thresh=2e-3,
            Ndraws=250, apply_agewt=True, apply_grad=True, lngalprior=None, lndustprior=None,
            dustfile=None, apply_dlabels=True, data_coords=None, logl_dim_prior=True, ltol=3e-2,
            ltol_subthresh=1e-2, logl_initthresh=5e-3, mag_max=50., merr_max=0.25, rstate=None,
            save_dar_draws=True, running_io=True, mem_lim=8000., verbose=True):
        """Fit every object and write the reference's output schema (brutus/fitting.py:1426-1800):
        ``labels, model_idx, ml_scale, ml_av, ml_rv, ml_cov_sar, obj_log_post, obj_log_evid,
        obj_chi2min, obj_Nbands[, samps_dist, samps_red, samps_dred, samps_logp]``."""
        (data, data_err, data_mask, data_labels, data_coords, lnprior, lngalprior, lndustprior,
         av_gauss, wt_thresh, rstate) = self._setup(
            data, data_err, data_mask, data_labels, phot_offsets=phot_offsets, parallax=parallax,
            parallax_err=parallax_err, av_gauss=av_gauss, lnprior=lnprior, wt_thresh=wt_thresh,
            cdf_thresh=cdf_thresh, apply_agewt=apply_agewt, apply_grad=apply_grad,
            lngalprior=lngalprior, lndustprior=lndustprior, dustfile=dustfile,
            data_coords=data_coords, ltol_subthresh=ltol_subthresh, logl_initthresh=logl_initthresh,
            mag_max=mag_max, merr_max=merr_max, rstate=rstate)
        ndata = data.shape[0]
        out = {"model_idx": np.full((ndata, Ndraws), -99, dtype="int32"),
               "ml_scale": np.ones((ndata, Ndraws), dtype="float32"),
               "ml_av": np.zeros((ndata, Ndraws), dtype="float32"),
               "ml_rv": np.zeros((ndata, Ndraws), dtype="float32"),
               "ml_cov_sar": np.zeros((ndata, Ndraws, 3, 3), dtype="float32"),
               "obj_log_post": np.zeros((ndata, Ndraws), dtype="float32"),
               "obj_log_evid": np.zeros(ndata, dtype="float32"),
               "obj_chi2min": np.zeros(ndata, dtype="float32"),
               "obj_Nbands": np.zeros(ndata, dtype="int16")}
        if save_dar_draws:
            for k in ("samps_dist", "samps_red", "samps_dred", "samps_logp"):
                out[k] = np.ones((ndata, Ndraws), dtype="float32")
        t0 = time.time()
        device_posterior = lngalprior is None
        gen = self._fit(data, data_err, data_mask, parallax=parallax, parallax_err=parallax_err,
                        avlim=avlim, rvlim=rvlim, av_gauss=av_gauss,
                        rv_gauss=rv_gauss, Nmc_prior=Nmc_prior, lnprior=lnprior, lnprior_ext=lnprior_ext,
                        wt_thresh=wt_thresh, cdf_thresh=cdf_thresh, Ndraws=Ndraws, rstate=rstate,
                        lngalprior=lngalprior, lndustprior=lndustprior, dustfile=dustfile,
                        apply_dlabels=apply_dlabels, data_coords=data_coords,
                        return_distreds=save_dar_draws, ltol_subthresh=ltol_subthresh,
                        logl_dim_prior=logl_dim_prior, logl_initthresh=logl_initthresh, ltol=ltol,
                        mem_lim=mem_lim, _device_arrays=device_posterior)
        if device_posterior:   # the device returns (Nbatch, Ndraws) arrays: assign them wholesale
            for b0, b1, r in gen:
                out["model_idx"][b0:b1], out["ml_scale"][b0:b1] = r["sidxs"], r["scales"]
                out["ml_av"][b0:b1], out["ml_rv"][b0:b1], out["ml_cov_sar"][b0:b1] = r["avs"], r["rvs"], r["cov_sar"]
                out["obj_Nbands"][b0:b1], out["obj_log_post"][b0:b1] = r["ndim"], r["lnprob"]
                out["obj_log_evid"][b0:b1], out["obj_chi2min"][b0:b1] = r["levid"], r["chi2min"]
                if save_dar_draws:
                    out["samps_dist"][b0:b1], out["samps_red"][b0:b1] = r["dists"], r["reds"]
                    out["samps_dred"][b0:b1], out["samps_logp"][b0:b1] = r["dreds"], r["logwts"]
                if verbose:
                    sys.stderr.write("\rFitted objects {:d}/{:d} (mean time: {:2.6f} s/obj)    ".format(
                        b1, ndata, (time.time() - t0) / b1))
                    sys.stderr.flush()
            gen = ()
        for i, r in enumerate(gen):
            out["model_idx"][i], out["ml_scale"][i], out["ml_av"][i], out["ml_rv"][i] = r[0], r[1], r[2], r[3]
            out["ml_cov_sar"][i], out["obj_Nbands"][i], out["obj_log_post"][i] = r[4], r[5], r[6]
            out["obj_log_evid"][i], out["obj_chi2min"][i] = r[7], r[8]
            if save_dar_draws:
                out["samps_dist"][i], out["samps_red"][i] = r[9], r[10]
                out["samps_dred"][i], out["samps_logp"][i] = r[11], r[12]
            if verbose:
                t_avg = (time.time() - t0) / (i + 1)
                sys.stderr.write("\rFitting object {:d}/{:d} [chi2/n: {:2.1f}/{:d}] (mean time: {:2.3f} s/obj, "
                                 "est. remaining: {:10.3f} s)    ".format(i + 1, ndata, r[8], r[5], t_avg,
                                                                          t_avg * (ndata - i - 1)))
                sys.stderr.flush()
        if verbose:
            sys.stderr.write("\n")
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            with h5py.File("{0}.h5".format(save_file), "w-") as f:
                f.create_dataset("labels", data=data_labels)
                for k, v in out.items():
                    f.create_dataset(k, data=v)
        else:
            np.savez("{0}.npz".format(save_file), labels=np.asarray(data_labels), **out)
        return out
