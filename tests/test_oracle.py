"""CPU-only: the C oracle (oracle/loglike_ref.c) against golden vectors from the unmodified
reference, and -- when /root/reference is present -- against the live reference."""
import numpy as np
import pytest

import golden_cases as gc
from oracle import ref_import

TOL = 1e-9  # float64 restatement vs float64 reference (observed: 1e-15 .. 4e-10)


@pytest.mark.parametrize("name", sorted(gc.LOGLIKE_CASES))
def test_oracle_matches_golden(oracle_mod, name):
    grid, labels, st, kw = gc.build_case(name)
    gold = gc.load_loglike(name)
    for i in range(len(st["flux"])):
        m = st["mask"][i].copy()
        out = oracle_mod.loglike(st["flux"][i], st["err"][i], m, grid, return_vals=True,
                                 parallax=st["parallax"][i], parallax_err=st["parallax_err"][i], **kw)
        assert np.array_equal(m, gold["mask_%d" % i])
        assert out[1] == int(gold["ndim_%d" % i])
        for key, val in zip(gc.KEYS, out):
            if key == "ndim":
                continue
            assert gc.rel_err(val, gold["%s_%d" % (key, i)]) < TOL, (name, i, key)


def test_oracle_threshold_error(oracle_mod):
    grid, labels, st, kw = gc.build_case("nodimprior")
    with pytest.raises(ValueError):  # brutus/fitting.py:691-693
        oracle_mod.loglike(st["flux"][0], st["err"][0], st["mask"][0].copy(), grid,
                           init_thresh=0.5, ltol_subthresh=1e-2)


def test_oracle_batch_matches_single(oracle_mod):
    grid, labels, st, kw = gc.build_case("mixed_9band")
    best, diag = oracle_mod.loglike_batch(st["flux"], st["err"], st["mask"], grid,
                                          parallax=st["parallax"], parallax_err=st["parallax_err"])
    for i in range(len(st["flux"])):
        out = oracle_mod.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), grid,
                                 return_vals=True, parallax=st["parallax"][i],
                                 parallax_err=st["parallax_err"][i], return_diag=True)
        k = int(np.argmax(out[0]))
        assert int(best[i, 0]) == k
        assert best[i, 2] == out[2][k] and best[i, 4] == out[4][k]
        assert diag[i, 1] == out[7]["n_iter_mag"] and diag[i, 2] == out[7]["n_iter_flux"]


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_live_reference(oracle_mod):
    from brutus_b200 import mock
    fit = ref_import.import_reference()
    grid, labels = mock.make_grid(4000, 7, seed=31)
    st = mock.make_stars(grid, 4, seed=32, dropout=0.1)
    gF = np.array(grid, order="F")
    for i in range(4):
        ref = fit.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), gF, return_vals=True,
                          parallax=st["parallax"][i], parallax_err=st["parallax_err"][i])
        out = oracle_mod.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), grid,
                                 return_vals=True, parallax=st["parallax"][i],
                                 parallax_err=st["parallax_err"][i])
        for key, a, b in zip(gc.KEYS, out, ref):
            if key != "ndim":
                assert gc.rel_err(a, b) < TOL, (i, key)
        # a-7 ops: reference lnpost's first stage (brutus/fitting.py:976-991)
        from brutus.pdf import scale_parallax_lnprior
        lnl, _, _, sc, _, _, ic = ref
        if np.isfinite(st["parallax"][i]):
            lp = lnl + scale_parallax_lnprior(sc, 1. / np.sqrt(np.abs(ic[:, 0, 0])),
                                              st["parallax"][i], st["parallax_err"][i])
        else:
            lp = lnl.copy()
        lp[~np.isfinite(lp)] = -1e300
        sel = np.where(lp > np.log(1e-3) + lp.max())[0]
        _, lnprob, sel_o = oracle_mod.select(out[0], out[3], out[6], st["parallax"][i],
                                             st["parallax_err"][i])
        assert np.array_equal(sel, sel_o)
        assert gc.rel_err(lnprob, lp) < TOL


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_live_reference_av_rv_init(oracle_mod):
    """Caller-supplied per-model start of the magnitude fit (brutus/fitting.py:583, :700-703), including the
    reference's in-place update of the caller's arrays (:202, :232, :809)."""
    from brutus_b200 import mock
    fit = ref_import.import_reference()
    grid, labels = mock.make_grid(3000, 6, seed=41, kind="locus")
    st = mock.make_stars(grid, 3, seed=42, dropout=0.1)
    gF = np.array(grid, order="F")
    rs = np.random.RandomState(43)
    a0, r0 = rs.uniform(0., 2.5, grid.shape[0]), rs.uniform(2.7, 4.1, grid.shape[0])
    for i in range(3):
        pk = dict(parallax=st["parallax"][i], parallax_err=st["parallax_err"][i])
        for kw in (dict(av_init=a0, rv_init=r0), dict(rv_init=r0), dict(av_init=a0, av_gauss=(0.5, 1.5))):
            kr = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            ko = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            ref = fit.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), gF, return_vals=True, **pk, **kr)
            ref = tuple(np.array(x) for x in ref)
            out = oracle_mod.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), grid, return_vals=True, **pk, **ko)
            for key, a, b in zip(gc.KEYS, out, ref):
                if key != "ndim":
                    assert gc.rel_err(a, b) < TOL, (i, sorted(kw), key)
            for k, pos in (("av_init", 4), ("rv_init", 5)):   # the reference returns the caller's arrays as av / rv
                if k in kr:
                    assert np.array_equal(kr[k], ref[pos])
