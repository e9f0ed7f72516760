"""Host-side mirror of the consumers of the fit outputs that re-evaluate SEDs on the same grid
(SURVEY.md section 8f row 4): ``get_seds`` (brutus/utils.py:1089-1159), ``phot_loglike`` (:1162-1222) and
``photometric_offsets`` (:1225-1400).

Everything that scales with the number of posterior samples runs on the device: ``_get_seds`` for the (model, Av,
Rv) samples (:286-347, ``bf_get_seds``) and, for ``photometric_offsets``, the leave-one-band-out chi2 likelihoods of
the samples and their normalised weights (``bf_offsets_weights``).  The bootstrap over objects stays on the host: it
consumes the caller's ``rstate`` exactly as the reference does, so results are reproducible draw for draw.
``phot_loglike`` itself is kept as the reference's public helper for arbitrary ``models`` arrays.  No CPU fallback
for the device parts.
"""
import sys

import numpy as np
from scipy.special import gammaln, xlogy

from . import fitting

__all__ = ["get_seds", "phot_loglike", "photometric_offsets"]


def get_seds(mag_coeffs, av=None, rv=None, return_flux=False, return_rvec=False, return_drvec=False,
             idx=None, precision="f32", device=0):
    """Drop-in for the reference's ``get_seds`` (brutus/utils.py:1089-1159): reddened SEDs of every model of
    ``mag_coeffs`` at ``(av, rv)`` (scalars or per-model arrays; defaults 0 and 3.3).  ``idx`` (extension)
    evaluates the models ``mag_coeffs[idx]`` instead, without copying rows of the grid on the host."""
    h = fitting.get_handle(mag_coeffs, precision=precision, device=device)
    n = mag_coeffs.shape[0] if idx is None else len(idx)
    if av is None:
        av = np.zeros(n)
    elif isinstance(av, (int, float)):
        av = np.full(n, float(av))
    if rv is None:
        rv = np.full(n, 3.3)
    elif isinstance(rv, (int, float)):
        rv = np.full(n, float(rv))
    seds, rvecs, drvecs = h.get_seds(av, rv, idx=idx, return_flux=return_flux, want_rvec=return_rvec,
                                     want_drvec=return_drvec)
    if return_rvec and return_drvec:
        return seds, rvecs, drvecs
    elif return_rvec:
        return seds, rvecs
    elif return_drvec:
        return seds, drvecs
    return seds


def phot_loglike(data, data_err, data_mask, models, dim_prior=True):
    """Log-likelihood of model fluxes ``models`` (Nmodel, Nfilt) given one object's photometry
    (brutus/utils.py:1162-1222)."""
    data_mask = np.asarray(data_mask, dtype=bool)
    ndim = int(np.sum(data_mask))
    flux, fluxerr = data[data_mask], data_err[data_mask]
    mfluxes = models[:, data_mask]
    tot_var = np.square(fluxerr) + np.zeros_like(mfluxes)
    chi2 = np.sum(np.square(flux - mfluxes) / tot_var, axis=1)
    lnl = -0.5 * chi2 - 0.5 * (ndim * np.log(2. * np.pi) + np.sum(np.log(tot_var), axis=1))
    if dim_prior:
        a = 0.5 * (ndim - 3)
        lnl = xlogy(a - 1., chi2) - (chi2 / 2.) - gammaln(a) - (np.log(2.) * a)
    return lnl


def photometric_offsets(phot, err, mask, models, idxs, reds, dreds, dists, sel=None, weights=None,
                        mask_fit=None, Nmc=150, old_offsets=None, dim_prior=True, prior_mean=None,
                        prior_std=None, verbose=True, rstate=None, precision="f32", device=0):
    """Drop-in for the reference's ``photometric_offsets`` (brutus/utils.py:1225-1400): multiplicative
    photometric offsets per band from the posterior samples ``(idxs, reds, dreds, dists)`` of a fit.
    Returns ``(ratios, ratios_err, nratio)``."""
    phot, err = np.asarray(phot, dtype=np.float64), np.asarray(err, dtype=np.float64)
    mask = np.asarray(mask, dtype=bool)
    nobj, nfilt = phot.shape
    nsamps = idxs.shape[1]
    if sel is None:
        sel = np.ones(nobj, dtype=bool)
    if weights is None:
        weights = np.ones((nobj, nsamps), dtype=float)
    if mask_fit is None:
        mask_fit = np.ones(nfilt, dtype=bool)
    if old_offsets is None:
        old_offsets = np.ones(nfilt)
    if rstate is None:
        rstate = np.random
    # On the device, over the (Nobj, Nsamps, Nfilt) samples (bf_offsets_weights): their SEDs (:1268-1271) and,
    # per fitted band, the likelihood weights of the samples with that band left out (:1299-1309)
    h = fitting.get_handle(models, precision=precision, device=device)
    seds, wt_dev = h.offsets_weights(phot, err, mask, idxs, reds, dreds, dists, old_offsets=old_offsets,
                                     mask_fit=mask_fit, dim_prior=dim_prior)
    ratios, nratio = np.ones(nfilt), np.zeros(nfilt, dtype=int)
    ratios_err = np.zeros(nfilt)
    for i in range(nfilt):
        nband = np.sum(mask, axis=1)
        if mask_fit[i]:   # (:1282-1286) observed, selected, > 3 bands besides this one
            s = np.where(mask[:, i] & sel & (nband > 3 + 1) & (np.sum(weights, axis=1) > 0))[0]
        else:             # (:1290-1291)
            s = np.where(mask[:, i] & sel & (nband > 3) & (np.sum(weights, axis=1) > 0))[0]
        n = len(s)
        nratio[i] = n
        if n == 0:
            continue
        ratio = seds[s, :, i] / phot[s, None, i]
        if mask_fit[i]:   # weights from the likelihood ignoring the current band (:1299-1309), from the device
            wt = wt_dev[i][s]
        else:
            wt = np.ones((n, nsamps))
        wt = wt * weights[s]
        wt /= wt.sum(axis=1)[:, None]
        wt_obj = np.array(np.sum(weights[s], axis=1) > 0, dtype=float)
        wt_obj /= sum(wt_obj)
        offsets = []
        for j in range(Nmc):   # bootstrap (:1320-1333)
            if verbose:
                sys.stderr.write("\rBand {0} ({1}/{2})     ".format(i + 1, j + 1, Nmc))
                sys.stderr.flush()
            ridx = rstate.choice(n, size=n, p=wt_obj)
            midx = [rstate.choice(nsamps, p=w) for w in wt[ridx]]
            offsets.append(np.median(ratio[ridx, midx]))
        ratios[i], ratios_err[i] = np.median(offsets), np.std(offsets)
    if verbose:
        sys.stderr.write("\n")
    if prior_mean is not None and prior_std is not None:   # (:1340-1343)
        var_tot = ratios_err ** 2 + prior_std ** 2
        ratios = (ratios * prior_std ** 2 + prior_mean * ratios_err ** 2) / var_tot
        ratios_err = ratios_err * prior_std / np.sqrt(var_tot)
    return ratios, ratios_err, nratio
