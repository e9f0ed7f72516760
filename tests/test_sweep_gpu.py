"""GPU parity tests for the B2 seam (bf_sweep_batch): the batched sweep with label priors and
lnpost's first selection, against the C oracle (loglike + select) star by star.

float64 kernels: identical selections and iteration counts; records within 1e-8 relative.
float32 kernels: selections may differ only for models within 2e-3 of the threshold; records of
common models within the float32 tolerances stated in tests/test_loglike_gpu.py.
"""
import numpy as np
import pytest

from brutus_b200 import mock

pytestmark = pytest.mark.gpu


def _oracle_star(oracle_mod, grid, st, i, labels=None, ext=None, **kw):
    pk = dict(parallax=st["parallax"][i], parallax_err=st["parallax_err"][i])
    ref = oracle_mod.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), grid, return_vals=True,
                             return_diag=True, **pk, **kw)
    ek = {}
    if labels is not None:
        ek = dict(labels=labels, ext_mean=ext[0][i], ext_std=ext[1][i])
    lnl, lnprob, sel = oracle_mod.select(ref[0], ref[3], ref[6], **pk, **ek)
    return ref, lnl, lnprob, sel


def _unpack6(ic):
    return np.stack([ic[:, 0, 0], ic[:, 0, 1], ic[:, 0, 2], ic[:, 1, 1], ic[:, 1, 2], ic[:, 2, 2]], axis=1)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_sweep_batch_vs_oracle(oracle_mod, precision):
    from brutus_b200 import _lib
    grid, labels = mock.make_grid(50_000, 8, seed=1200)
    st = mock.make_stars(grid, 24, seed=2200, dropout=0.05)
    st["flux"][3, 2] = -0.2 * abs(st["flux"][3, 2])
    lab = np.stack([labels["Mr"], labels["feh"]])
    rs = np.random.RandomState(5)
    ext_mean = np.stack([st["truth"]["idx"] * 0 + labels["Mr"][st["truth"]["idx"]] + rs.normal(0, 0.3, 24),
                         np.full(24, np.nan)], axis=1)
    ext_std = np.stack([np.full(24, 0.5), np.full(24, 0.2)], axis=1)
    ext_std[::3, 0] = -1.0  # inactive constraints (brutus/fitting.py:1999)
    h = _lib.Handle(0, precision)
    h.set_grid(grid)
    h.set_labels(lab)
    res = h.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"],
                        ext_mean=ext_mean, ext_std=ext_std, copy=True)
    stats = h.stats()
    h.close()
    assert stats["kernel_launches"] > 0
    assert res["offsets"][0] == 0 and res["offsets"][-1] == len(res["model_idx"])
    import parity
    for i in range(24):
        ref, lnl, lnprob, sel = parity.oracle_star(oracle_mod, grid, st, i, labels=lab, ext=(ext_mean, ext_std))
        parity.check_star(res, i, ref, lnl, lnprob, sel, precision, tag=(precision, i))


def test_batch_equals_single_star():
    """Shard invariance: a star's records do not depend on what else is in the batch (bitwise)."""
    from brutus_b200 import _lib
    grid, labels = mock.make_grid(20_000, 6, seed=1300)
    st = mock.make_stars(grid, 40, seed=2300)
    h = _lib.Handle(0, "f32")
    h.set_grid(grid)
    full = h.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], copy=True)
    for i in (0, 17, 39):
        one = h.sweep_batch(st["flux"][i:i + 1], st["err"][i:i + 1], st["mask"][i:i + 1],
                            st["parallax"][i:i + 1], st["parallax_err"][i:i + 1])
        lo, hi = full["offsets"][i], full["offsets"][i + 1]
        assert np.array_equal(one["model_idx"], full["model_idx"][lo:hi])
        for k in ("lnl", "chi2", "scale", "av", "rv"):
            assert np.array_equal(one[k], full[k][lo:hi]), k
        assert np.array_equal(one["icov6"], full["icov6"][:, lo:hi])
    h.close()


def test_empty_and_errors():
    from brutus_b200 import _lib
    grid, labels = mock.make_grid(5_000, 5, seed=1400)
    h = _lib.Handle(0, "f32")
    with pytest.raises(_lib.BrutusCudaError):  # no grid yet
        h.sweep_batch(np.ones((1, 5)), np.ones((1, 5)), np.ones((1, 5), bool))
    h.set_grid(grid)
    res = h.sweep_batch(np.zeros((0, 5)), np.zeros((0, 5)), np.zeros((0, 5), bool))
    assert len(res["model_idx"]) == 0 and res["offsets"].tolist() == [0]
    with pytest.raises(ValueError):
        h.sweep_batch(np.ones((1, 5)), np.ones((1, 5)), np.ones((1, 5), bool),
                      opts=_lib.make_options(init_thresh=0.5))
    with pytest.raises(ValueError):
        h.set_grid(np.zeros((10, 17, 3), np.float32))
    h.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_candidate_fallback_is_exact(precision):
    """The sweep only keeps a candidate superset; a negative slack forces the 'redo with every model
    as a candidate' path for every star and must reproduce the default path bit for bit, as must a
    tiny candidate pool (many groups) and a tiny star batch."""
    import os
    from brutus_b200 import _lib
    grid, labels = mock.make_grid(30_000, 7, seed=1500)
    st = mock.make_stars(grid, 12, seed=2500, dropout=0.1)
    args = (st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"])
    h = _lib.Handle(0, precision)
    h.set_grid(grid)
    a = h.sweep_batch(*args, copy=True)
    sa = h.stats()
    b = h.sweep_batch(*args, opts=_lib.make_options(select_slack=-100.), copy=True)
    sb = h.stats()
    h.close()
    assert sa["fallbacks"] == 0 and sb["fallbacks"] == 12
    assert 0 < sa["candidates"] < 12 * 30_000 < sb["candidates"]
    os.environ["BRUTUS_B200_BATCH"] = "5"
    os.environ["BRUTUS_B200_POOL"] = "1"
    try:
        h = _lib.Handle(0, precision)
        h.set_grid(grid)
        c = h.sweep_batch(*args, copy=True)
        h.close()
    finally:
        del os.environ["BRUTUS_B200_BATCH"], os.environ["BRUTUS_B200_POOL"]
    for other in (b, c):
        for k in ("offsets", "model_idx", "lnl", "chi2", "scale", "av", "rv", "icov6", "n_iter", "n_surv",
                  "max_lnprob", "ndim"):
            assert np.array_equal(a[k], other[k]), k
