"""ctypes binding of libbrutus_b200.so (C ABI: include/brutus_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present, every
compute entry point raises :class:`BrutusCudaError`.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbrutus_b200.so")

BF_OK, BF_E_INVALID, BF_E_CUDA, BF_E_NOGRID, BF_E_CAPACITY, BF_E_THRESH, BF_E_NOMEM = 0, -1, -2, -3, -4, -5, -6
LAYOUT_C, LAYOUT_F = 0, 1
PRECISION_F32, PRECISION_F64 = 0, 1
MAX_FILT = 16

# every symbol include/brutus_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ("bf_default_options", "bf_create", "bf_destroy", "bf_last_error", "bf_set_grid",
           "bf_set_grid_device", "bf_set_labels", "bf_loglike_full", "bf_sweep_batch",
           "bf_get_stats", "bf_flush_l2", "bf_device_count", "bf_version",
           "bf_default_gal_params", "bf_default_post_options", "bf_set_model_priors", "bf_fit_batch",
           "bf_get_seds", "bf_offsets_weights", "bf_set_init", "bf_get_trace", "bf_create_multi", "bf_num_devices", "bf_nccl_unique_id",
           "bf_nccl_init", "bf_set_grid_bcast", "bf_bcast_host", "bf_allreduce_max")


class BrutusCudaError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [("avlim", C.c_double * 2), ("av_gauss", C.c_double * 2), ("rvlim", C.c_double * 2),
                ("rv_gauss", C.c_double * 2), ("ltol", C.c_double), ("ltol_subthresh", C.c_double),
                ("init_thresh", C.c_double), ("wt_thresh", C.c_double), ("select_slack", C.c_double),
                ("dim_prior", C.c_int32),
                ("max_iter", C.c_int32), ("apply_parallax_clip", C.c_int32), ("skip_d2h", C.c_int32)]


class Records(C.Structure):
    _fields_ = [("n", C.c_int64), ("stride", C.c_int64), ("elem_size", C.c_int32),
                ("nrows", C.c_int32), ("model_idx", C.c_void_p), ("rows", C.c_void_p)]


REC_BASIC, REC_FIT, REC_FULL = 3, 5, 11
ROW_NAMES = ("lnl", "scale", "av", "chi2", "rv")


class Stats(C.Structure):
    _fields_ = [("ms_device", C.c_double), ("ms_magfit", C.c_double), ("ms_flux", C.c_double),
                ("ms_select", C.c_double), ("kernel_launches", C.c_int64),
                ("magfit_launches", C.c_int64), ("magfit_star_passes", C.c_int64),
                ("resweeps", C.c_int64), ("candidates", C.c_int64), ("fallbacks", C.c_int64),
                ("survivors", C.c_int64),
                ("selected", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("ms_post", C.c_double), ("selected2", C.c_int64), ("clipped", C.c_int64),
                ("fixups", C.c_int64), ("flux_more_launches", C.c_int64), ("regroups", C.c_int64),
                ("unconverged", C.c_int64), ("host_syncs", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


GAL_FIELDS = ("R_solar", "Z_solar", "R_thin", "Z_thin", "Rs_thin", "R_thick", "Z_thick", "f_thick",
              "Rs_thick", "Rs_halo", "q_halo_ctr", "q_halo_inf", "r_q_halo", "eta_halo", "f_halo",
              "feh_thin", "feh_thin_sigma", "feh_thick", "feh_thick_sigma", "feh_halo", "feh_halo_sigma",
              "max_age", "min_age", "feh_age_ctr", "feh_age_scale", "nsigma_from_max_age", "max_sigma",
              "min_sigma", "galcen_distance", "z_sun")


class GalParams(C.Structure):
    """bf_gal_params: keyword arguments of gal_lnprior (brutus/pdf.py:476-486) + frame constants."""
    _fields_ = [(k, C.c_double) for k in GAL_FIELDS]


class PostOptions(C.Structure):
    _fields_ = [("nmc_prior", C.c_int32), ("ndraws", C.c_int32), ("seed", C.c_uint64),
                ("use_gal_prior", C.c_int32), ("reserved", C.c_int32), ("star_base", C.c_int64),
                ("nsel_max", C.c_int64), ("gal", GalParams),
                ("z_override", C.c_void_p), ("u_override", C.c_void_p)]


class Draws(C.Structure):
    _fields_ = [("model_idx", C.c_void_p)] + [(k, C.c_void_p) for k in
                                              ("scale", "av", "rv", "cov_sar", "lnprob", "dist", "red",
                                               "dred", "logwt")]


_lib = None


def load():
    """Load the shared library (building is __graft_entry__.build()/brutus_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BrutusCudaError("libbrutus_b200.so not built (run `python -m brutus_b200.build`); "
                              "brutus_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, dp, fp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)
    u8p, i32p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    op = C.POINTER(Options)
    lib.bf_default_options.argtypes = [op]
    lib.bf_default_options.restype = None
    lib.bf_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    lib.bf_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(vp)]
    lib.bf_num_devices.argtypes = [vp]
    lib.bf_nccl_unique_id.argtypes = [vp]
    lib.bf_nccl_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.bf_set_grid_bcast.argtypes = [vp, fp, C.c_int64, C.c_int32, C.c_int32, C.c_int32]
    lib.bf_bcast_host.argtypes = [vp, vp, C.c_int64, C.c_int]
    lib.bf_allreduce_max.argtypes = [vp, dp, C.c_int32]
    lib.bf_destroy.argtypes = [vp]
    lib.bf_last_error.argtypes = [vp]
    lib.bf_last_error.restype = C.c_char_p
    lib.bf_set_grid.argtypes = [vp, fp, C.c_int64, C.c_int32, C.c_int32]
    lib.bf_set_grid_device.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32]
    lib.bf_set_labels.argtypes = [vp, dp, C.c_int32]
    lib.bf_loglike_full.argtypes = [vp, dp, dp, u8p, C.c_double, C.c_double, op,
                                    dp, dp, dp, dp, dp, dp, u8p, i64p]
    lib.bf_sweep_batch.argtypes = [vp, C.c_int64, dp, dp, u8p, dp, dp, dp, dp, op, C.c_int32,
                                   i32p, i32p, i64p, dp, i64p, C.POINTER(Records)]
    lib.bf_default_gal_params.argtypes = [C.POINTER(GalParams)]
    lib.bf_default_gal_params.restype = None
    lib.bf_default_post_options.argtypes = [C.POINTER(PostOptions)]
    lib.bf_default_post_options.restype = None
    lib.bf_set_model_priors.argtypes = [vp, dp, dp, dp]
    lib.bf_fit_batch.argtypes = [vp, C.c_int64, dp, dp, u8p, dp, dp, dp, dp, dp, op,
                                 C.POINTER(PostOptions), i32p, i32p, i64p, dp, dp, C.POINTER(Draws)]
    lib.bf_get_seds.argtypes = [vp, C.c_int64, i32p, dp, dp, C.c_int32, dp, dp, dp]
    lib.bf_offsets_weights.argtypes = [vp, C.c_int64, C.c_int32, dp, dp, u8p, i32p, dp, dp, dp, dp, u8p, C.c_int32, dp, dp]
    lib.bf_set_init.argtypes = [vp, dp, dp]
    lib.bf_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.bf_get_trace.argtypes = [vp]
    lib.bf_get_trace.restype = C.c_char_p
    lib.bf_flush_l2.argtypes = [vp]
    lib.bf_device_count.restype = C.c_int
    lib.bf_version.restype = C.c_char_p
    _lib = lib
    return lib


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def make_options(avlim=(0., 20.), av_gauss=(0., 1e6), rvlim=(1., 8.), rv_gauss=(3.32, 0.18),
                 dim_prior=True, ltol=3e-2, ltol_subthresh=1e-2, init_thresh=5e-3, wt_thresh=1e-3,
                 max_iter=0, apply_parallax_clip=True, skip_d2h=False, select_slack=0.5):
    if av_gauss is None:  # brutus/fitting.py:695-696
        av_gauss = (0., 1e6)
    o = Options()
    o.avlim[:] = [float(x) for x in avlim]
    o.av_gauss[:] = [float(x) for x in av_gauss]
    o.rvlim[:] = [float(x) for x in rvlim]
    o.rv_gauss[:] = [float(x) for x in rv_gauss]
    o.ltol, o.ltol_subthresh, o.init_thresh = float(ltol), float(ltol_subthresh), float(init_thresh)
    o.wt_thresh = float(wt_thresh)
    o.select_slack = float(select_slack)
    o.dim_prior, o.max_iter = int(bool(dim_prior)), int(max_iter)
    o.apply_parallax_clip = int(bool(apply_parallax_clip))
    o.skip_d2h = int(bool(skip_d2h))
    return o


class Handle:
    """A sweep engine on one CUDA device, or -- ``device`` a sequence of ordinals -- on several devices of this
    process (``bf_create_multi``: the batch calls shard the stars over them).  Not re-entrant."""

    def __init__(self, device=0, precision="f32"):
        self._lib = load()
        self._h = C.c_void_p()
        prec = {"f32": PRECISION_F32, "f64": PRECISION_F64}[precision]
        devs = [int(d) for d in device] if isinstance(device, (list, tuple, np.ndarray)) else [int(device)]
        arr = (C.c_int * len(devs))(*devs)
        rc = self._lib.bf_create_multi(arr, len(devs), prec, C.byref(self._h))
        if rc != BF_OK:
            raise BrutusCudaError(self._lib.bf_last_error(None).decode())
        self.precision = precision
        self.devices = devs
        self.device = devs[0]
        self.nmodel = 0
        self.nfilt = 0
        self.nlabel = 0
        self.rank, self.world = 0, 1
        self._grid_token = None

    # ---- one process per GPU: NCCL process group inside the library (no PyTorch) ----
    @staticmethod
    def nccl_unique_id():
        """128-byte NCCL id, to be created on rank 0 and handed to every rank."""
        buf = C.create_string_buffer(128)
        rc = load().bf_nccl_unique_id(buf)
        if rc != BF_OK:
            raise BrutusCudaError(load().bf_last_error(None).decode())
        return buf.raw

    def nccl_init(self, unique_id, rank, world):
        self._check(self._lib.bf_nccl_init(self._h, C.c_char_p(bytes(unique_id)), int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def set_grid_bcast(self, mag_coeffs, shape, root=0):
        """Replicate the grid held by rank ``root`` (``mag_coeffs`` may be None elsewhere) with one NCCL broadcast."""
        nmodel, nfilt = int(shape[0]), int(shape[1])
        ptr, layout = None, LAYOUT_C
        if self.rank == root:
            a = np.asarray(mag_coeffs)
            if a.shape != (nmodel, nfilt, 3):
                raise ValueError("grid shape %r does not match %r" % (a.shape, tuple(shape)))
            if a.dtype != np.float32:
                a = a.astype(np.float32)
            if a.flags.f_contiguous and not a.flags.c_contiguous:
                layout = LAYOUT_F
            elif not a.flags.c_contiguous:
                a = np.ascontiguousarray(a)
            ptr = _ptr(a, C.c_float)
        layout = int(self.bcast_array(np.array([layout], dtype=np.int64), root)[0])
        self._check(self._lib.bf_set_grid_bcast(self._h, ptr, nmodel, nfilt, layout, int(root)))
        self.nmodel, self.nfilt, self.nlabel = nmodel, nfilt, 0

    def bcast_array(self, a, root=0):
        """In-place broadcast of a C-contiguous NumPy array of identical shape/dtype on every rank."""
        a = np.ascontiguousarray(a)
        self._check(self._lib.bf_bcast_host(self._h, a.ctypes.data_as(C.c_void_p), a.nbytes, int(root)))
        return a

    def allreduce_max(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.float64).copy()
        self._check(self._lib.bf_allreduce_max(self._h, _ptr(v, C.c_double), v.size))
        return v

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bf_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc):
        if rc != BF_OK:
            msg = self._lib.bf_last_error(self._h).decode()
            if rc in (BF_E_THRESH, BF_E_INVALID):
                raise ValueError(msg)
            raise BrutusCudaError("rc=%d: %s" % (rc, msg))

    def set_grid(self, mag_coeffs):
        """Stage a float32 (Nmodel, Nfilt, 3) grid (C or Fortran order, no copy for either)."""
        a = np.asarray(mag_coeffs)
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("mag_coeffs must have shape (Nmodel, Nfilt, 3)")
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        if a.flags.c_contiguous:
            layout = LAYOUT_C
        elif a.flags.f_contiguous:
            layout = LAYOUT_F
        else:
            a, layout = np.ascontiguousarray(a), LAYOUT_C
        self._check(self._lib.bf_set_grid(self._h, _ptr(a, C.c_float), a.shape[0], a.shape[1], layout))
        self.nmodel, self.nfilt = a.shape[0], a.shape[1]
        self.nlabel = 0

    def set_grid_device(self, dev_ptr, nmodel, nfilt, layout=LAYOUT_C):
        self._check(self._lib.bf_set_grid_device(self._h, C.c_void_p(int(dev_ptr)), int(nmodel),
                                                 int(nfilt), int(layout)))
        self.nmodel, self.nfilt = int(nmodel), int(nfilt)
        self.nlabel = 0

    def set_labels(self, cols):
        """cols: (nlabel, Nmodel) float64."""
        a = np.ascontiguousarray(cols, dtype=np.float64)
        if a.ndim != 2 or a.shape[1] != self.nmodel:
            raise ValueError("labels must have shape (nlabel, Nmodel)")
        self._check(self._lib.bf_set_labels(self._h, _ptr(a, C.c_double), a.shape[0]))
        self.nlabel = a.shape[0]

    def flush_l2(self):
        self._check(self._lib.bf_flush_l2(self._h))

    def stats(self):
        s = Stats()
        self._lib.bf_get_stats(self._h, C.byref(s))
        return s.as_dict()

    def _warn_unconverged(self):
        n = self.stats()["unconverged"]
        if n:
            import warnings
            warnings.warn("%d star(s) hit the iteration cap (bf_options.max_iter) before the reference's convergence "
                          "test passed; their results are those of the last iteration" % n, RuntimeWarning, stacklevel=3)

    def trace(self):
        """Per-kernel device times since the last call (needs BRUTUS_B200_TRACE=1 when the handle is created):
        ``{kernel: (launches, ms)}``."""
        out = {}
        for line in self._lib.bf_get_trace(self._h).decode().splitlines():
            name, n, _, ms, _ = line.split()
            out[name] = (int(n), float(ms))
        return out

    def loglike_full(self, flux, err, mask, parallax, parallax_err, opts, want_icov=True):
        n = self.nmodel
        f = np.ascontiguousarray(flux, dtype=np.float64)
        e = np.ascontiguousarray(err, dtype=np.float64)
        m = np.ascontiguousarray(mask).astype(np.uint8)
        if f.shape != (self.nfilt,) or e.shape != (self.nfilt,) or m.shape != (self.nfilt,):
            raise ValueError("data, data_err and data_mask must have shape (Nfilt,)")
        lnl, chi2, sc, av, rv = (np.empty(n) for _ in range(5))
        icov = np.empty((n, 3, 3)) if want_icov else None
        mclean = np.zeros(self.nfilt, dtype=np.uint8)
        diag = np.zeros(4, dtype=np.int64)
        self._check(self._lib.bf_loglike_full(
            self._h, _ptr(f, C.c_double), _ptr(e, C.c_double), _ptr(m, C.c_uint8), float(parallax),
            float(parallax_err), C.byref(opts), _ptr(lnl, C.c_double), _ptr(chi2, C.c_double),
            _ptr(sc, C.c_double), _ptr(av, C.c_double), _ptr(rv, C.c_double), _ptr(icov, C.c_double),
            _ptr(mclean, C.c_uint8), _ptr(diag, C.c_int64)))
        self._warn_unconverged()
        return lnl, chi2, sc, av, rv, icov, mclean.astype(bool), diag

    def sweep_batch(self, flux, err, mask, parallax=None, parallax_err=None, ext_mean=None,
                    ext_std=None, opts=None, rows=REC_FULL, copy=False):
        """B2 entry point.  Returns per-star arrays plus the CSR-compacted records of the selected
        models: ``model_idx, lnl, scale, av[, chi2, rv[, icov6 (6, n)]]``.  Unless ``copy`` is set
        the record arrays are zero-copy views of the library's pinned result arena and are only
        valid until the next ``sweep_batch`` call on this handle."""
        f = np.ascontiguousarray(flux, dtype=np.float64)
        e = np.ascontiguousarray(err, dtype=np.float64)
        m = np.ascontiguousarray(mask).astype(np.uint8)
        ns = f.shape[0]
        if not self.nfilt:
            raise BrutusCudaError("no grid staged (call set_grid first)")
        if f.ndim != 2 or f.shape[1] != self.nfilt or e.shape != f.shape or m.shape != f.shape:
            raise ValueError("data, data_err and data_mask must have shape (Ndata, Nfilt)")
        par = None if parallax is None else np.ascontiguousarray(parallax, dtype=np.float64)
        perr = None if parallax_err is None else np.ascontiguousarray(parallax_err, dtype=np.float64)
        em = es = None
        if ext_mean is not None and self.nlabel:
            em = np.ascontiguousarray(ext_mean, dtype=np.float64)
            es = np.ascontiguousarray(ext_std, dtype=np.float64)
        if opts is None:
            opts = make_options()
        ndim = np.zeros(ns, dtype=np.int32)
        nit = np.zeros((ns, 2), dtype=np.int32)
        nsurv = np.zeros(ns, dtype=np.int64)
        mx = np.zeros(ns)
        offsets = np.zeros(ns + 1, dtype=np.int64)
        rec = Records()
        self._check(self._lib.bf_sweep_batch(
            self._h, ns, _ptr(f, C.c_double), _ptr(e, C.c_double), _ptr(m, C.c_uint8),
            _ptr(par, C.c_double), _ptr(perr, C.c_double), _ptr(em, C.c_double),
            _ptr(es, C.c_double), C.byref(opts), int(rows), _ptr(ndim, C.c_int32),
            _ptr(nit, C.c_int32), _ptr(nsurv, C.c_int64), _ptr(mx, C.c_double),
            _ptr(offsets, C.c_int64), C.byref(rec)))
        self._warn_unconverged()
        n = int(rec.n)
        out = dict(ndim=ndim, n_iter=nit, n_surv=nsurv, max_lnprob=mx, offsets=offsets)
        dt = np.float32 if rec.elem_size == 4 else np.float64
        if n > 0:
            idx = np.ctypeslib.as_array(C.cast(rec.model_idx, C.POINTER(C.c_int32)), shape=(n,))
            mat = np.ctypeslib.as_array(C.cast(rec.rows, C.POINTER(C.c_float if rec.elem_size == 4 else C.c_double)),
                                        shape=(rec.nrows, int(rec.stride)))[:, :n]
        else:
            idx = np.zeros(0, dtype=np.int32)
            mat = np.zeros((rec.nrows, 0), dtype=dt)
        if copy:
            idx, mat = idx.copy(), mat.copy()
        out["model_idx"] = idx
        for k, name in enumerate(ROW_NAMES[:min(5, rec.nrows)]):
            out[name] = mat[k]
        out["icov6"] = mat[5:11] if rec.nrows >= 11 else None
        return out

    def get_seds(self, av, rv, idx=None, return_flux=False, want_rvec=True, want_drvec=True):
        """``_get_seds`` (brutus/utils.py:286-347) on the staged grid for the models ``idx`` (default: every
        model, in order).  Returns ``(seds, rvecs, drvecs)``, float64 ``(n, Nfilt)`` (None where not wanted)."""
        a = np.ascontiguousarray(av, dtype=np.float64)
        r = np.ascontiguousarray(rv, dtype=np.float64)
        ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
        n = a.shape[0]
        if a.ndim != 1 or r.shape != a.shape or (ix is not None and ix.shape != a.shape):
            raise ValueError("av, rv (and idx) must be 1-D arrays of the same length")
        seds = np.empty((n, self.nfilt))
        rvecs = np.empty((n, self.nfilt)) if want_rvec else None
        drvecs = np.empty((n, self.nfilt)) if want_drvec else None
        self._check(self._lib.bf_get_seds(self._h, n, _ptr(ix, C.c_int32), _ptr(a, C.c_double), _ptr(r, C.c_double),
                                          int(bool(return_flux)), _ptr(seds, C.c_double),
                                          _ptr(rvecs, C.c_double), _ptr(drvecs, C.c_double)))
        return seds, rvecs, drvecs

    def offsets_weights(self, phot, err, mask, idxs, reds, dreds, dists, old_offsets=None, mask_fit=None,
                        dim_prior=True):
        """The device part of ``photometric_offsets`` (brutus/utils.py:1268-1271, :1299-1309): returns
        ``(seds, wt)``, the flux SEDs of the posterior samples ``(Nobj, Nsamps, Nfilt)`` and, for every band with
        ``mask_fit``, the likelihood weights of the samples with that band left out, ``(Nfilt, Nobj, Nsamps)``."""
        phot = np.ascontiguousarray(phot, dtype=np.float64)
        err = np.ascontiguousarray(err, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        ix = np.ascontiguousarray(idxs, dtype=np.int32)
        nobj, nfilt = phot.shape
        if nfilt != self.nfilt or err.shape != phot.shape or mask.shape != phot.shape or ix.ndim != 2 or ix.shape[0] != nobj:
            raise ValueError("phot, err, mask must be (Nobj, Nfilt) and idxs (Nobj, Nsamps)")
        nsamps = ix.shape[1]
        a = np.ascontiguousarray(reds, dtype=np.float64)
        r = np.ascontiguousarray(dreds, dtype=np.float64)
        d = np.ascontiguousarray(dists, dtype=np.float64)
        if a.shape != ix.shape or r.shape != ix.shape or d.shape != ix.shape:
            raise ValueError("reds, dreds, dists must have the shape of idxs")
        oo = None if old_offsets is None else np.ascontiguousarray(old_offsets, dtype=np.float64)
        mf = np.ones(nfilt, dtype=np.uint8) if mask_fit is None else np.ascontiguousarray(mask_fit, dtype=np.uint8)
        if (oo is not None and oo.shape != (nfilt,)) or mf.shape != (nfilt,):
            raise ValueError("old_offsets and mask_fit must have Nfilt entries")
        seds = np.empty((nobj, nsamps, nfilt))
        wt = np.empty((nfilt, nobj, nsamps))
        self._check(self._lib.bf_offsets_weights(
            self._h, nobj, nsamps, _ptr(phot, C.c_double), _ptr(err, C.c_double), _ptr(mask, C.c_uint8),
            _ptr(ix, C.c_int32), _ptr(a, C.c_double), _ptr(r, C.c_double), _ptr(d, C.c_double),
            _ptr(oo, C.c_double), _ptr(mf, C.c_uint8), int(bool(dim_prior)), _ptr(seds, C.c_double),
            _ptr(wt, C.c_double)))
        return seds, wt

    def set_init(self, av_init=None, rv_init=None):
        """Per-model start of the magnitude fit for ``loglike_full`` (``av_init`` / ``rv_init`` of the reference's
        ``loglike``, brutus/fitting.py:700-703); ``set_init()`` restores the default (the prior means)."""
        if av_init is None and rv_init is None:
            self._check(self._lib.bf_set_init(self._h, None, None))
            return
        a = np.ascontiguousarray(av_init, dtype=np.float64)
        r = np.ascontiguousarray(rv_init, dtype=np.float64)
        if a.shape != (self.nmodel,) or r.shape != (self.nmodel,):
            raise ValueError("av_init and rv_init must have one entry per model")
        self._check(self._lib.bf_set_init(self._h, _ptr(a, C.c_double), _ptr(r, C.c_double)))

    def set_model_priors(self, lnprior=None, feh=None, loga=None):
        """Stage the static inputs of lnpost: the `lnprior` grid (brutus/fitting.py:1004) and the label
        columns 'feh' / 'loga' of the Galactic prior (brutus/pdf.py:669, :694).  None = absent."""
        arrs = []
        for a in (lnprior, feh, loga):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                if a.shape != (self.nmodel,):
                    raise ValueError("per-model priors / labels must have shape (Nmodel,)")
            arrs.append(a)
        self._check(self._lib.bf_set_model_priors(self._h, *[_ptr(a, C.c_double) for a in arrs]))

    def fit_batch(self, flux, err, mask, parallax=None, parallax_err=None, coords=None, ext_mean=None,
                  ext_std=None, opts=None, nmc_prior=50, ndraws=250, seed=0, use_gal_prior=True,
                  gal=None, star_base=0, mem_lim=None, z_override=None, u_override=None, copy=True):
        """The per-object body of ``BruteForce._fit`` on the device (``bf_fit_batch``): returns a dict
        with the reference's 13-tuple members as (Ndata, Ndraws) arrays (``sidxs, scales, avs, rvs,
        cov_sar, lnprob, dists, reds, dreds, logwts``) and per-object ``ndim, levid, chi2min, nsel,
        n_iter``.  With ``copy=False`` the draw arrays are zero-copy views of the library's pinned result
        arena, valid until the next ``fit_batch`` call on this handle."""
        f = np.ascontiguousarray(flux, dtype=np.float64)
        e = np.ascontiguousarray(err, dtype=np.float64)
        m = np.ascontiguousarray(mask).astype(np.uint8)
        ns = f.shape[0]
        if not self.nfilt:
            raise BrutusCudaError("no grid staged (call set_grid first)")
        if f.ndim != 2 or f.shape[1] != self.nfilt or e.shape != f.shape or m.shape != f.shape:
            raise ValueError("data, data_err and data_mask must have shape (Ndata, Nfilt)")
        par = None if parallax is None else np.ascontiguousarray(parallax, dtype=np.float64)
        perr = None if parallax_err is None else np.ascontiguousarray(parallax_err, dtype=np.float64)
        co = None if coords is None else np.ascontiguousarray(coords, dtype=np.float64)
        if co is not None and co.shape != (ns, 2):
            raise ValueError("data_coords must have shape (Ndata, 2)")
        em = es = None
        if ext_mean is not None and self.nlabel:
            em = np.ascontiguousarray(ext_mean, dtype=np.float64)
            es = np.ascontiguousarray(ext_std, dtype=np.float64)
        if opts is None:
            opts = make_options()
        po = PostOptions()
        self._lib.bf_default_post_options(C.byref(po))
        po.nmc_prior, po.ndraws, po.seed = int(nmc_prior), int(ndraws), int(seed) & (2 ** 64 - 1)
        po.use_gal_prior = int(bool(use_gal_prior))
        po.star_base = int(star_base)
        # lnpost's memory clip: Nsel_max = int(mem_lim / Nmc_prior / 4e-4) (brutus/fitting.py:969-970); None = off
        po.nsel_max = 0 if mem_lim is None else max(1, int(float(mem_lim) / po.nmc_prior / 4.0e-4))
        for k, v in (gal or {}).items():
            if k not in GAL_FIELDS:
                raise ValueError("unknown Galactic prior parameter %r" % k)
            setattr(po.gal, k, float(v))
        keep = []
        if z_override is not None:
            z = np.ascontiguousarray(z_override, dtype=np.float64)
            if z.shape != (self.nmodel, 3, po.nmc_prior):
                raise ValueError("z_override must have shape (Nmodel, 3, Nmc_prior)")
            po.z_override = z.ctypes.data
            keep.append(z)
        if u_override is not None:
            u = np.ascontiguousarray(u_override, dtype=np.float64)
            if u.shape != (ns, 2, po.ndraws):
                raise ValueError("u_override must have shape (Ndata, 2, Ndraws)")
            po.u_override = u.ctypes.data
            keep.append(u)
        nd = po.ndraws
        dr = Draws()
        ndim = np.zeros(ns, dtype=np.int32)
        nit = np.zeros((ns, 2), dtype=np.int32)
        nsel = np.zeros(ns, dtype=np.int64)
        levid = np.zeros(ns)
        chi2min = np.zeros(ns)
        self._check(self._lib.bf_fit_batch(
            self._h, ns, _ptr(f, C.c_double), _ptr(e, C.c_double), _ptr(m, C.c_uint8),
            _ptr(par, C.c_double), _ptr(perr, C.c_double), _ptr(co, C.c_double), _ptr(em, C.c_double),
            _ptr(es, C.c_double), C.byref(opts), C.byref(po), _ptr(ndim, C.c_int32),
            _ptr(nit, C.c_int32), _ptr(nsel, C.c_int64), _ptr(levid, C.c_double),
            _ptr(chi2min, C.c_double), C.byref(dr)))
        self._warn_unconverged()
        del keep
        out = {}
        if ns > 0:
            out["sidxs"] = np.ctypeslib.as_array(C.cast(dr.model_idx, C.POINTER(C.c_int32)), shape=(ns, nd))
            for cname, key in (("scale", "scales"), ("av", "avs"), ("rv", "rvs"), ("lnprob", "lnprob"),
                               ("dist", "dists"), ("red", "reds"), ("dred", "dreds"), ("logwt", "logwts")):
                out[key] = np.ctypeslib.as_array(C.cast(getattr(dr, cname), C.POINTER(C.c_double)), shape=(ns, nd))
            out["cov_sar"] = np.ctypeslib.as_array(C.cast(dr.cov_sar, C.POINTER(C.c_double)), shape=(ns, nd, 3, 3))
            if copy:
                out = {k: v.copy() for k, v in out.items()}
        else:
            out = dict(sidxs=np.zeros((0, nd), dtype=np.int32), cov_sar=np.zeros((0, nd, 3, 3)))
            for key in ("scales", "avs", "rvs", "lnprob", "dists", "reds", "dreds", "logwts"):
                out[key] = np.zeros((0, nd))
        out.update(ndim=ndim, n_iter=nit, nsel=nsel, levid=levid, chi2min=chi2min)
        return out
