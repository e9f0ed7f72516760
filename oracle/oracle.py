"""ctypes wrapper around oracle/loglike_ref.c (TEST INFRASTRUCTURE ONLY; parity pinned, see the
header of loglike_ref.c).  Mirrors the call signature of the reference's ``loglike``
(brutus/fitting.py:579-585)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libbrutus_oracle.so")
_lib = None


class RefOptions(C.Structure):
    _fields_ = [("avlim", C.c_double * 2), ("av_gauss", C.c_double * 2),
                ("rvlim", C.c_double * 2), ("rv_gauss", C.c_double * 2),
                ("ltol", C.c_double), ("ltol_subthresh", C.c_double),
                ("init_thresh", C.c_double), ("dim_prior", C.c_int32), ("max_iter", C.c_int32)]


def build(force=False):
    """Compile the C restatement into oracle/_ref/ (gcc; a few seconds)."""
    src = os.path.join(_HERE, "loglike_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_SO)
        dp, fp, u8p, i64p = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8),
                             C.POINTER(C.c_int64))
        lib.brutus_ref_loglike.restype = C.c_int
        lib.brutus_ref_loglike.argtypes = [dp, dp, u8p, C.c_int, fp, C.c_int64,
                                           C.POINTER(RefOptions), C.c_double, C.c_double,
                                           dp, dp, dp, dp, dp, dp, i64p, u8p, dp]
        lib.brutus_ref_loglike_init.restype = C.c_int
        lib.brutus_ref_loglike_init.argtypes = [dp, dp, u8p, C.c_int, fp, C.c_int64,
                                                C.POINTER(RefOptions), C.c_double, C.c_double, dp, dp,
                                                dp, dp, dp, dp, dp, dp, i64p, u8p, dp]
        lib.brutus_ref_select.restype = C.c_int64
        lib.brutus_ref_select.argtypes = [C.c_int64, dp, dp, dp, C.c_int, C.c_double, C.c_double,
                                          C.c_int, dp, dp, dp, C.c_double, dp, u8p]
        lib.brutus_ref_loglike_batch.restype = C.c_int
        lib.brutus_ref_loglike_batch.argtypes = [C.c_int, dp, dp, u8p, C.c_int, fp, C.c_int64,
                                                 C.POINTER(RefOptions), dp, dp, dp, i64p, C.c_int]
        lib.brutus_ref_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def make_options(avlim=(0., 20.), av_gauss=(0., 1e6), rvlim=(1., 8.), rv_gauss=(3.32, 0.18),
                 dim_prior=True, ltol=3e-2, ltol_subthresh=1e-2, init_thresh=5e-3, max_iter=0):
    if av_gauss is None:  # brutus/fitting.py:695-696
        av_gauss = (0., 1e6)
    o = RefOptions()
    o.avlim[:] = avlim
    o.av_gauss[:] = av_gauss
    o.rvlim[:] = rvlim
    o.rv_gauss[:] = rv_gauss
    o.ltol, o.ltol_subthresh, o.init_thresh = ltol, ltol_subthresh, init_thresh
    o.dim_prior, o.max_iter = int(bool(dim_prior)), max_iter
    return o


def loglike(data, data_err, data_mask, mag_coeffs, parallax=None, parallax_err=None,
            return_vals=False, return_diag=False, av_init=None, rv_init=None, **kwargs):
    """Same contract as the reference ``loglike`` (brutus/fitting.py:579): returns
    ``(lnl, Ndim, chi2[, scale, av, rv, icov_sar])`` and cleans ``data_mask`` in place."""
    lib = _load()
    o = make_options(**kwargs)
    if o.init_thresh > o.ltol_subthresh:
        raise ValueError("The initial threshold must be smaller than or equal "
                         "to the final threshold applied to be useful!")
    co = np.ascontiguousarray(mag_coeffs, dtype=np.float32)
    n, nf, _ = co.shape
    d = np.ascontiguousarray(data, dtype=np.float64)
    e = np.ascontiguousarray(data_err, dtype=np.float64)
    m = np.ascontiguousarray(data_mask).astype(np.uint8)
    par = np.nan if parallax is None or parallax_err is None else float(parallax)
    perr = np.nan if parallax is None or parallax_err is None else float(parallax_err)
    lnl, chi2, sc, av, rv = (np.empty(n) for _ in range(5))
    icov = np.empty((n, 3, 3))
    diag = np.zeros(4, dtype=np.int64)
    surv = np.zeros(n, dtype=np.uint8)
    lnlp = np.zeros(n)
    a0 = None if av_init is None else np.ascontiguousarray(av_init, dtype=np.float64)
    r0 = None if rv_init is None else np.ascontiguousarray(rv_init, dtype=np.float64)
    rc = lib.brutus_ref_loglike_init(_p(d, C.c_double), _p(e, C.c_double), _p(m, C.c_uint8), nf,
                                     _p(co, C.c_float), n, C.byref(o), par, perr,
                                     None if a0 is None else _p(a0, C.c_double),
                                     None if r0 is None else _p(r0, C.c_double),
                                     _p(lnl, C.c_double), _p(chi2, C.c_double), _p(sc, C.c_double),
                                _p(av, C.c_double), _p(rv, C.c_double), _p(icov, C.c_double),
                                _p(diag, C.c_int64), _p(surv, C.c_uint8), _p(lnlp, C.c_double))
    if rc:
        raise RuntimeError("oracle failed rc=%d" % rc)
    data_mask[...] = m.astype(bool)  # in-place clean-up, brutus/fitting.py:709
    out = (lnl, int(diag[0]), chi2)
    if return_vals:
        out = out + (sc, av, rv, icov)
    if return_diag:
        out = out + ({"n_iter_mag": int(diag[1]), "n_iter_flux": int(diag[2]),
                      "n_surv": int(diag[3]), "survivors": surv.astype(bool), "lnl_p": lnlp},)
    return out


def select(lnl, scale, icov, parallax=None, parallax_err=None, labels=None, ext_mean=None,
           ext_std=None, wt_thresh=1e-3):
    """lnprior_ext + rough parallax prior + first threshold.
    Returns (lnl_with_ext, lnprob, sel_indices)."""
    lib = _load()
    n = len(lnl)
    lnl = np.array(lnl, dtype=np.float64)
    sc = np.ascontiguousarray(scale, dtype=np.float64)
    ic = np.ascontiguousarray(icov, dtype=np.float64)
    have = parallax is not None and parallax_err is not None
    nl = 0 if labels is None else len(labels)
    lab = np.ascontiguousarray(labels if nl else np.zeros((1, 1)), dtype=np.float64)
    em = np.ascontiguousarray(ext_mean if nl else np.zeros(1), dtype=np.float64)
    es = np.ascontiguousarray(ext_std if nl else np.zeros(1), dtype=np.float64)
    lnprob = np.empty(n)
    flag = np.zeros(n, dtype=np.uint8)
    lib.brutus_ref_select(n, _p(lnl, C.c_double), _p(sc, C.c_double), _p(ic, C.c_double),
                          int(have), float(parallax) if have else np.nan,
                          float(parallax_err) if have else np.nan, nl, _p(lab, C.c_double),
                          _p(em, C.c_double), _p(es, C.c_double), wt_thresh,
                          _p(lnprob, C.c_double), _p(flag, C.c_uint8))
    return lnl, lnprob, np.where(flag)[0]


def loglike_batch(flux, err, mask, mag_coeffs, parallax=None, parallax_err=None, nthreads=0,
                  **kwargs):
    """OpenMP-over-stars batch (CPU baseline).  Returns (best (nstar,6), diag (nstar,4))."""
    lib = _load()
    o = make_options(**kwargs)
    co = np.ascontiguousarray(mag_coeffs, dtype=np.float32)
    n, nf, _ = co.shape
    f = np.ascontiguousarray(flux, dtype=np.float64)
    e = np.ascontiguousarray(err, dtype=np.float64)
    m = np.ascontiguousarray(mask).astype(np.uint8)
    ns = f.shape[0]
    par = None if parallax is None else np.ascontiguousarray(parallax, dtype=np.float64)
    perr = None if parallax_err is None else np.ascontiguousarray(parallax_err, dtype=np.float64)
    best = np.zeros((ns, 6))
    diag = np.zeros((ns, 4), dtype=np.int64)
    rc = lib.brutus_ref_loglike_batch(ns, _p(f, C.c_double), _p(e, C.c_double), _p(m, C.c_uint8),
                                      nf, _p(co, C.c_float), n, C.byref(o),
                                      _p(par, C.c_double) if par is not None else None,
                                      _p(perr, C.c_double) if perr is not None else None,
                                      _p(best, C.c_double), _p(diag, C.c_int64), int(nthreads))
    if rc:
        raise RuntimeError("oracle batch failed rc=%d" % rc)
    return best, diag


def num_threads():
    return int(_load().brutus_ref_num_threads())
