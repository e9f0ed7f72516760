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
