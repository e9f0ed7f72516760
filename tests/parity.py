"""Shared comparison of the CUDA path's compacted records (bf_sweep_batch) with the oracle, star by star.

Stated tolerances of the float32 kernels against the float64 oracle, for a model whose own
chi2 is <= 100; for a worse-fitting model every tolerance is multiplied by f = chi2 / 100 (the float32 residuals
t_j = d_j/sigma_j - s M_j/sigma_j carry a rounding error relative to the S/N, not to t_j: the fit of a model that
misses the data by hundreds of sigma is conditioned accordingly worse -- and matters accordingly less):
  chi2, lnl   |d| <= 2e-3 + 2e-5 |x|
  av          |d| <= 3e-4 f          rv  |d| <= 3e-3 f
  scale       rel 2e-5 f + 0.92 (1.3 |d av| + 0.06 Av |d rv|)   (conditional MLE at the fitted reddening)
  icov        2e-3 f of sqrt(|ii| |jj|)
  max_lnprob  2e-3 + 2e-5 |x|
Up to 1 % of a star's models may miss these by up to 2x, and a handful (<= max(2, n / 5000)) by up to 20x: the reference divides a model's
stepsize by 1.2 whenever lnl_new < lnl_old (brutus/fitting.py:802), and a float32 run can take that discrete
decision differently from float64 when the two are equal to rounding, after which the model follows a slightly
different path to the same tolerance.  Scales far below the star's typical scale (|s| < 1e-2 median: the MLE
numerator cancels, the 1e-20 floor is near) are compared in absolute terms.
Membership: the selection (brutus/fitting.py:988-991) may differ from the oracle's only for models within
2e-3 + 2e-5 |thr| of the selection threshold.  The cull (:758-759) decides whether a model is flux-refined, so
a model within 2e-3 + 2e-5 |thr| of the CULL threshold may legitimately carry either its magnitude-fit or its
refined values: those (few) models are exempt from the value comparisons and counted.  When such a model is the
star's best (the refinement can lift a barely-surviving model to the top; 1 of the 1 517 NGC 2682 objects on the
C4 lattice), the maximum of lnprob and the selection threshold are accepted under either reading.
Iteration counts: the magnitude-loop count must equal the oracle's; so must the flux-loop count, except that a float32
run may stop one iteration earlier or later than float64 when the reference's convergence test
(max |lnl_new - lnl_old| <= ltol over the near-best survivors, brutus/fitting.py:798-799) lands within float32 rounding
of ltol.  Such a knife-edge star (1 of the 1 517 NGC 2682 objects) is reported to the caller, which bounds how many there
may be; its records are then compared loosely (one refinement step more or less moves lnl by less than ltol).
The float64 kernels must match to 1e-8 relative (icov 1e-7) with identical sets and counts."""
import numpy as np


def unpack6(ic):
    return np.stack([ic[:, 0, 0], ic[:, 0, 1], ic[:, 0, 2], ic[:, 1, 1], ic[:, 1, 2], ic[:, 2, 2]], axis=1)


def oracle_star(oracle_mod, grid, st, i, labels=None, ext=None, wt_thresh=1e-3, **kw):
    pk = dict(parallax=st["parallax"][i], parallax_err=st["parallax_err"][i])
    ref = oracle_mod.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), grid, return_vals=True,
                             return_diag=True, **pk, **kw)
    ek = {}
    if labels is not None:
        ek = dict(labels=labels, ext_mean=ext[0][i], ext_std=ext[1][i])
    lnl, lnprob, sel = oracle_mod.select(ref[0], ref[3], ref[6], wt_thresh=wt_thresh, **pk, **ek)
    return ref, lnl, lnprob, sel


def star_records(res, i):
    lo, hi = res["offsets"][i], res["offsets"][i + 1]
    rec = {k: np.asarray(res[k][lo:hi], dtype=np.float64) for k in ("lnl", "chi2", "scale", "av", "rv")}
    rec["model_idx"] = res["model_idx"][lo:hi]
    rec["icov6"] = np.asarray(res["icov6"][:, lo:hi], dtype=np.float64).T if res.get("icov6") is not None else None
    return rec


class KnifeEdge(int):
    """Return value of check_star for a float32 star whose flux loop stopped one iteration off (see the header)."""


def check_star(res, i, ref, lnl, lnprob, sel, precision, init_thresh=5e-3, wt_thresh=1e-3, tag=None):
    """Compare star i of `res` with the oracle outputs; returns the number of cull-borderline models exempted (a
    KnifeEdge instance if the star's flux-loop count is off by one in float32)."""
    rec = star_records(res, i)
    idx = rec["model_idx"]
    diag = ref[7]
    assert np.all(np.diff(idx) > 0), tag
    assert res["ndim"][i] == ref[1], tag
    assert res["n_iter"][i][0] == diag["n_iter_mag"], (tag, res["n_iter"][i], diag["n_iter_mag"])
    if precision == "f32" and abs(int(res["n_iter"][i][1]) - diag["n_iter_flux"]) == 1:
        common, ia, ib = np.intersect1d(idx, sel, return_indices=True)
        assert len(common) >= 0.97 * max(len(sel), len(idx)) - 2, (tag, "knife-edge star: selection", len(common), len(sel))
        assert np.max(np.abs(rec["chi2"][ia] - ref[2][common])) < 0.2, (tag, "knife-edge star: chi2")
        assert abs(res["max_lnprob"][i] - lnprob.max()) < 0.1, (tag, "knife-edge star: max_lnprob")
        return KnifeEdge(0)
    assert tuple(res["n_iter"][i]) == (diag["n_iter_mag"], diag["n_iter_flux"]), (tag, res["n_iter"][i])
    mx = lnprob.max()
    if precision == "f64":
        assert res["n_surv"][i] == diag["n_surv"], tag
        assert np.array_equal(idx, sel), tag
        assert abs(res["max_lnprob"][i] - mx) < 1e-8 * max(1, abs(mx)), tag
        for k, r in (("lnl", lnl), ("chi2", ref[2]), ("scale", ref[3])):
            assert np.max(np.abs(rec[k] - r[sel]) / np.maximum(np.abs(r[sel]), 1e-300)) < 1e-8, (tag, k)
        for k, r in (("av", ref[4]), ("rv", ref[5])):     # may sit at a bound such as 0: absolute
            assert np.max(np.abs(rec[k] - r[sel])) < 1e-8, (tag, k)
        if rec["icov6"] is not None:
            r6 = unpack6(ref[6][sel])
            assert np.max(np.abs(rec["icov6"] - r6) / np.maximum(np.abs(r6), 1e-300)) < 1e-7, tag
        return 0
    # ---- float32 ----
    thr = mx + np.log(wt_thresh)
    lp = diag["lnl_p"]
    cthr = lp.max() + np.log(init_thresh)
    near_cull = np.abs(lp - cthr) < 2e-3 + 2e-5 * abs(cthr)      # may or may not have been flux-refined
    sym = np.setxor1d(idx, sel)
    sym = sym[~near_cull[sym]]
    # A cull-borderline model may be the best model of the star (the flux refinement can lift a model that barely
    # survived the cull to the top): whether it counts moves the maximum of lnprob, and with it the selection
    # threshold.  Both readings are accepted: the threshold lies in [thr_lo, thr].
    mx_sure = lnprob[~near_cull].max() if (~near_cull).any() else mx
    thr_lo = min(mx, mx_sure) + np.log(wt_thresh)
    # the threshold (a maximum of lnprob) and the model's own lnprob each carry the lnl tolerance
    mtol = 6e-3 + 6e-5 * abs(thr)
    dm = np.maximum(np.maximum(lnprob[sym] - thr, thr_lo - lnprob[sym]), 0.)
    assert (dm > mtol).sum() <= max(2, len(sel) // 5000) and not np.any(dm > 20 * mtol), (tag, "selection membership", len(sym))
    tolm = 2e-3 + 2e-5 * abs(mx)
    assert min(mx, mx_sure) - tolm < res["max_lnprob"][i] < mx + tolm, tag
    common, ia, ib = np.intersect1d(idx, sel, return_indices=True)
    border = near_cull[common]
    nb = int(border.sum())
    assert nb <= max(32, len(common) // 20), (tag, "too many cull-borderline models", nb)
    assert abs(int(res["n_surv"][i]) - diag["n_surv"]) <= int(near_cull.sum()), tag
    ok = ~border
    c, a = common[ok], ia[ok]
    f = np.maximum(1., ref[2][c] / 100.)     # per-model factor: see the header

    nout = max(2, len(c) // 5000)

    def close(x, y, atol, rtol, name):
        d = np.abs(x - y)
        lim = atol + rtol * np.abs(y)
        bad = d > lim
        # soft edge: up to 1 % of the models may sit within 2x of the tolerance, `nout` models within 20x
        assert bad.mean() <= 0.01 and (d > 2 * lim).sum() <= nout and not np.any(d > 20 * lim), \
            (tag, name, int(bad.sum()), float((d / lim).max()), int(c[np.argmax(d / lim)]))
    close(rec["chi2"][a], ref[2][c], 2e-3, 2e-5, "chi2")
    close(rec["lnl"][a], lnl[c], 2e-3, 2e-5, "lnl")
    close(rec["av"][a], ref[4][c], 3e-4 * f, 0, "av")
    close(rec["rv"][a], ref[5][c], 3e-3 * f, 0, "rv")
    prop = 0.92 * (1.3 * np.abs(rec["av"][a] - ref[4][c]) + 0.06 * np.abs(ref[4][c]) * np.abs(rec["rv"][a] - ref[5][c]))
    sref = np.maximum(np.abs(ref[3][c]), 1e-2 * np.median(np.abs(ref[3][c])))
    dsc = np.abs(rec["scale"][a] - ref[3][c]) / sref
    lim = 2e-5 * f + prop
    bad = dsc > lim
    assert bad.mean() <= 0.01 and (dsc > 2 * lim).sum() <= nout and not np.any(dsc > 20 * lim), \
        (tag, "scale", int(bad.sum()), float((dsc / lim).max()))
    if rec["icov6"] is not None:
        r6 = unpack6(ref[6][c])
        d = np.sqrt(np.abs(r6[:, [0, 3, 5]]))
        sc = np.stack([d[:, 0] * d[:, 0], d[:, 0] * d[:, 1], d[:, 0] * d[:, 2], d[:, 1] * d[:, 1], d[:, 1] * d[:, 2],
                       d[:, 2] * d[:, 2]], axis=1)
        e6 = np.abs(rec["icov6"][a] - r6) / sc / f[:, None]
        assert (e6.max(axis=1) > 2e-3).sum() <= nout and e6.max() < 4e-2, (tag, "icov", float(e6.max()))
    return nb
