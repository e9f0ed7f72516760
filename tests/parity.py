"""Shared comparison of the CUDA path's compacted records (bf_sweep_batch) with the oracle, star by star.

Stated tolerances of the float32 kernels against the float64 oracle:
  chi2, lnl   |d| <= 2e-3 + 2e-5 |x|
  av          |d| <= 2e-4            rv  |d| <= 2e-3
  scale       rel 2e-5 + 0.92 (1.3 |d av| + 0.06 Av |d rv|)   (conditional MLE at the fitted reddening)
  icov        2e-3 of sqrt(|ii| |jj|)
  max_lnprob  2e-3 + 2e-5 |x|
Membership: the selection (brutus/fitting.py:988-991) may differ from the oracle's only for models within
2e-3 + 2e-5 |thr| of the selection threshold.  The cull (:758-759) decides whether a model is flux-refined, so
a model within 2e-3 + 2e-5 |thr| of the CULL threshold may legitimately carry either its magnitude-fit or its
refined values: those (few) models are exempt from the value comparisons and counted.
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


def check_star(res, i, ref, lnl, lnprob, sel, precision, init_thresh=5e-3, wt_thresh=1e-3, tag=None):
    """Compare star i of `res` with the oracle outputs; returns the number of cull-borderline models exempted."""
    rec = star_records(res, i)
    idx = rec["model_idx"]
    diag = ref[7]
    assert np.all(np.diff(idx) > 0), tag
    assert res["ndim"][i] == ref[1], tag
    assert tuple(res["n_iter"][i]) == (diag["n_iter_mag"], diag["n_iter_flux"]), (tag, res["n_iter"][i])
    mx = lnprob.max()
    if precision == "f64":
        assert res["n_surv"][i] == diag["n_surv"], tag
        assert np.array_equal(idx, sel), tag
        assert abs(res["max_lnprob"][i] - mx) < 1e-8 * max(1, abs(mx)), tag
        for k, r in (("lnl", lnl), ("chi2", ref[2]), ("scale", ref[3]), ("av", ref[4]), ("rv", ref[5])):
            assert np.max(np.abs(rec[k] - r[sel]) / np.maximum(np.abs(r[sel]), 1e-300)) < 1e-8, (tag, k)
        if rec["icov6"] is not None:
            r6 = unpack6(ref[6][sel])
            assert np.max(np.abs(rec["icov6"] - r6) / np.maximum(np.abs(r6), 1e-300)) < 1e-7, tag
        return 0
    # ---- float32 ----
    thr = mx + np.log(wt_thresh)
    sym = np.setxor1d(idx, sel)
    assert np.all(np.abs(lnprob[sym] - thr) < 2e-3 + 2e-5 * abs(thr)), (tag, "selection membership", len(sym))
    assert abs(res["max_lnprob"][i] - mx) < 2e-3 + 2e-5 * abs(mx), tag
    lp = diag["lnl_p"]
    cthr = lp.max() + np.log(init_thresh)
    common, ia, ib = np.intersect1d(idx, sel, return_indices=True)
    border = np.abs(lp[common] - cthr) < 2e-3 + 2e-5 * abs(cthr)
    nb = int(border.sum())
    assert nb <= max(16, len(common) // 1000), (tag, "too many cull-borderline models", nb)
    assert abs(int(res["n_surv"][i]) - diag["n_surv"]) <= int((np.abs(lp - cthr) < 2e-3 + 2e-5 * abs(cthr)).sum()), tag
    ok = ~border
    c, a = common[ok], ia[ok]

    def close(x, y, atol, rtol, name):
        d = np.abs(x - y)
        bad = d > atol + rtol * np.abs(y)
        assert not bad.any(), (tag, name, int(bad.sum()), float(d[bad].max()), int(c[np.argmax(d)]))
    close(rec["chi2"][a], ref[2][c], 2e-3, 2e-5, "chi2")
    close(rec["lnl"][a], lnl[c], 2e-3, 2e-5, "lnl")
    close(rec["av"][a], ref[4][c], 2e-4, 0, "av")
    close(rec["rv"][a], ref[5][c], 2e-3, 0, "rv")
    prop = 0.92 * (1.3 * np.abs(rec["av"][a] - ref[4][c]) + 0.06 * np.abs(ref[4][c]) * np.abs(rec["rv"][a] - ref[5][c]))
    dsc = np.abs(rec["scale"][a] / ref[3][c] - 1)
    bad = dsc > 2e-5 + prop
    assert not bad.any(), (tag, "scale", int(bad.sum()), float(dsc[bad].max()))
    if rec["icov6"] is not None:
        r6 = unpack6(ref[6][c])
        d = np.sqrt(np.abs(r6[:, [0, 3, 5]]))
        sc = np.stack([d[:, 0] * d[:, 0], d[:, 0] * d[:, 1], d[:, 0] * d[:, 2], d[:, 1] * d[:, 1], d[:, 1] * d[:, 2],
                       d[:, 2] * d[:, 2]], axis=1)
        e6 = np.abs(rec["icov6"][a] - r6) / sc
        assert e6.max() < 2e-3, (tag, "icov", float(e6.max()))
    return nb
