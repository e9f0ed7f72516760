"""CPU: the host-side mirror of the reference's argument handling (no GPU is touched: every error below is raised
before the first device call, like in the reference) and the small prior functions it carries."""
import numpy as np
import pytest

import golden_cases as gc
from brutus_b200 import fitting, mock
from oracle import ref_import


@pytest.fixture(scope="module")
def bf_case():
    grid, labels = mock.make_grid(500, 6, seed=3)
    st = mock.make_stars(grid, 4, seed=4)
    lmask = np.ones(1, dtype=[("Mr", bool), ("feh", bool)])
    return fitting.BruteForce(grid, labels, lmask), st


def _kw(st, **over):
    kw = dict(parallax=st["parallax"], parallax_err=st["parallax_err"], lnprior=np.zeros(500),
              lngalprior=gc.toy_galprior, data_coords=np.zeros((4, 2)), verbose=False)
    kw.update(over)
    return kw


def test_threshold_order_is_checked_first(bf_case, tmp_path):      # brutus/fitting.py:1305-1308
    bf, st = bf_case
    with pytest.raises(ValueError, match="threshold"):
        bf.fit(st["flux"], st["err"], st["mask"], np.arange(4), str(tmp_path / "a"),
               **_kw(st, logl_initthresh=0.5, ltol_subthresh=1e-2))


def test_parallax_needs_its_error(bf_case, tmp_path):              # brutus/fitting.py:1325-1327
    bf, st = bf_case
    with pytest.raises(ValueError, match="parallax"):
        bf.fit(st["flux"], st["err"], st["mask"], np.arange(4), str(tmp_path / "b"), **_kw(st, parallax_err=None))


def test_fewer_than_four_bands(bf_case, tmp_path):                 # brutus/fitting.py:1413-1420
    bf, st = bf_case
    bad = st["mask"].copy()
    bad[1, :3] = False
    with pytest.raises(ValueError, match="fewer than 4 bands"):
        bf.fit(st["flux"], st["err"], bad, np.arange(4), str(tmp_path / "c"), **_kw(st))
    # mag_max / merr_max cuts count too (:1404-1411): a faint-magnitude cut that removes every band
    with pytest.raises(ValueError):
        bf.fit(st["flux"], st["err"], st["mask"], np.arange(4), str(tmp_path / "d"), **_kw(st, mag_max=-100.))


def test_default_prior_needs_coordinates(bf_case, tmp_path):       # brutus/fitting.py:1362-1365
    bf, st = bf_case
    with pytest.raises(ValueError, match="data_coords"):
        bf.fit(st["flux"], st["err"], st["mask"], np.arange(4), str(tmp_path / "e"),
               **_kw(st, lngalprior=None, data_coords=None, dustfile=None))


def test_lnprior_ext_keys_are_validated(bf_case):                  # brutus/fitting.py:1972-1976
    bf, st = bf_case
    with pytest.raises(ValueError, match="lnprior_ext"):
        next(bf._fit(st["flux"], st["err"], st["mask"], lnprior_ext={"nope": np.zeros((4, 2))},
                     **{k: v for k, v in _kw(st).items() if k != "verbose"}))


def test_loglike_argument_errors():
    grid, _ = mock.make_grid(200, 5, seed=5)
    st = mock.make_stars(grid, 1, seed=6)
    with pytest.raises(ValueError, match="initial threshold"):    # brutus/fitting.py:691-693
        fitting.loglike(st["flux"][0], st["err"][0], st["mask"][0].copy(), grid, init_thresh=0.5)
    with pytest.raises(ValueError, match="av_init"):              # one start value per model (:700-703)
        fitting.loglike(st["flux"][0], st["err"][0], st["mask"][0].copy(), grid, av_init=np.zeros(199))


def test_imf_prior_is_normalised_and_broken_at_half_a_solar_mass():
    m = np.linspace(0.0801, 100., 2_000_001)
    p = np.exp(fitting.imf_lnprior(m))
    assert abs(np.trapezoid(p, m) - 1.) < 2e-3          # brutus/pdf.py:38-108 normalises over [0.08, inf)
    assert np.all(np.isneginf(fitting.imf_lnprior(np.array([0.05, 0.08]))))
    lo, hi = fitting.imf_lnprior(np.array([0.4999999, 0.5000001]))
    assert abs(lo - hi) < 1e-5                          # continuous at the break


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_small_priors_match_the_live_reference():
    ref_import.import_reference()
    from brutus import pdf as rpdf
    rs = np.random.RandomState(8)
    m = 10. ** rs.uniform(-1.2, 1.5, 500)
    assert np.allclose(fitting.imf_lnprior(m), rpdf.imf_lnprior(m), rtol=1e-12, atol=0, equal_nan=True)
    p = rs.uniform(0.05, 3., 200)
    for pm, pe in ((1.0, 0.1), (0.2, 0.1), (np.nan, 0.1), (-0.3, 0.05)):
        assert np.allclose(fitting.parallax_lnprior(p, pm, pe), rpdf.parallax_lnprior(p, pm, pe), rtol=1e-12)
        s, se = p ** 2, rs.uniform(0.01, 0.5, 200)
        assert np.allclose(fitting.scale_parallax_lnprior(s, se, pm, pe), rpdf.scale_parallax_lnprior(s, se, pm, pe),
                           rtol=1e-12)


def test_result_store_is_exclusive_and_incremental(tmp_path):
    """fit()'s writer (brutus/fitting.py:1632-1662, :1734-1748): refuses an existing output BEFORE fitting ("w-"), and
    with running_io the rows fitted so far are on disk after every batch."""
    target = str(tmp_path / "out")
    st = fitting._ResultStore(target, np.arange(5), 5, 3, True, True)
    assert set(st.arrays) >= {"model_idx", "ml_scale", "ml_av", "ml_rv", "ml_cov_sar", "obj_log_post", "obj_log_evid",
                              "obj_chi2min", "obj_Nbands", "samps_dist", "samps_red", "samps_dred", "samps_logp"}
    assert st.arrays["model_idx"].dtype == np.int32 and st.arrays["obj_Nbands"].dtype == np.int16
    assert st.arrays["ml_cov_sar"].shape == (5, 3, 3, 3) and np.all(st.arrays["model_idx"] == -99)
    with pytest.raises((FileExistsError, OSError)):
        fitting._ResultStore(target, np.arange(5), 5, 3, True, True)
    st.write(0, 2, {"model_idx": np.full((2, 3), 7), "obj_log_evid": np.array([1.5, 2.5])})
    if st.h5 is None:    # no h5py in this image: the rows are already in the archive, written in place
        from brutus_b200 import npzstore
        part = npzstore.read_partial(target + ".npz")
        assert np.all(part["model_idx"][:2] == 7) and np.all(part["model_idx"][2:] == -99)
        assert list(part["row_done"]) == [1, 1, 0, 0, 0]
    st.close()
    if st.h5 is None and not (tmp_path / "out.h5").exists():
        d = np.load(target + ".npz")
        assert np.all(d["model_idx"][:2] == 7) and d["obj_log_evid"][1] == 2.5 and np.array_equal(d["labels"], np.arange(5))
        import zipfile
        assert zipfile.ZipFile(target + ".npz").testzip() is None      # checksums patched on completion
    with pytest.raises((FileExistsError, OSError)):        # still exclusive after completion
        fitting._ResultStore(target, np.arange(5), 5, 3, True, False)


def test_grid_fingerprint_sees_in_place_edits():
    """The handle cache re-stages a grid that was modified in place (the reference re-reads the array every call)."""
    grid, _ = mock.make_grid(20_000, 6, seed=9)
    fp = fitting._fingerprint(grid)
    assert fitting._fingerprint(grid) == fp
    grid[:, :, 0] += 0.25
    assert fitting._fingerprint(grid) != fp
    gF = np.asfortranarray(grid)
    assert fitting._fingerprint(gF) != fitting._fingerprint(grid)


def test_cdf_select_is_the_reference_rule():
    rs = np.random.RandomState(1)
    lnp = rs.normal(0, 3, 500)
    sel = fitting._cdf_select(lnp, 2e-3)
    order = np.argsort(lnp)
    p = np.exp(lnp - np.logaddexp.reduce(lnp))
    assert np.array_equal(sel, order[np.cumsum(p[order]) <= 1 - 2e-3])


def test_default_prior_is_ps1_lf_without_mini(bf_case):            # brutus/fitting.py:1335-1341
    bf, st = bf_case
    out = bf._setup(st["flux"], st["err"], st["mask"], parallax=st["parallax"], parallax_err=st["parallax_err"],
                    lngalprior=gc.toy_galprior, data_coords=np.zeros((4, 2)), apply_agewt=False, apply_grad=False)
    from brutus_b200 import pdf
    assert np.allclose(out[5], pdf.ps1_MrLF_lnprior(bf.models_labels["Mr"]))


def test_two_launch_sweep_tile_partition():
    """The sweep is issued in two launches: the model tiles that are multiples of S, then the
    others, enumerated as j -> j + j // (S - 1) + 1 (csrc/kernels_nb.cuh, k_sweep; csrc/api.cu, process_group).
    Restated here: together they visit every tile exactly once, for the S the host derives from the tile count."""
    for ntile in (1024, 1025, 1031, 3907, 4096, 11719, 40000):
        S = min(32, max(2, ntile // 128))
        nsub = (ntile + S - 1) // S
        first = [j * S for j in range(nsub)]
        rest = [j + j // (S - 1) + 1 for j in range(ntile - nsub)]
        assert sorted(first + rest) == list(range(ntile)), (ntile, S)
        assert len(first) >= 128 and all(t % S for t in rest)
