"""CPU: the host-side priors of the product (brutus_b200/pdf.py) against golden values of the unmodified reference
(tests/golden/galprior.npz, tests/golden/priors.npz -- tests/gen_golden.py) and, in the build container, against
the live reference functions."""
import os

import numpy as np
import pytest

import golden_cases as gc
from brutus_b200 import pdf
from oracle import ref_import


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(gc.GOLD, "galprior.npz"))


@pytest.mark.parametrize("kind,names", [("full", ("feh", "loga")), ("feh_only", ("feh",)), ("nolabels", ())])
def test_gal_lnprior_matches_reference(gold, kind, names):
    lab = None
    if names:
        lab = np.zeros(len(gold["dists"]), dtype=[(n, "f8") for n in names])
        for n in names:
            lab[n] = gold[n]
    for k, c in enumerate(gold["coords"]):
        ref = gold["%s_%d" % (kind, k)]
        out = pdf.gal_lnprior(gold["dists"], tuple(c), labels=lab)
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(out)) and fin.sum() > 300
        assert np.max(np.abs(out[fin] - ref[fin])) < 1e-10


def test_static_priors_match_golden():
    g = np.load(os.path.join(gc.GOLD, "priors.npz"))
    assert gc.rel_err(pdf.imf_lnprior(g["mini"]), g["imf"]) < 1e-12
    assert gc.rel_err(pdf.imf_lnprior(g["mini"], mgrid2=g["mini2"]), g["imf_binary"]) < 1e-12
    assert np.max(np.abs(pdf.ps1_MrLF_lnprior(g["Mr"]) - g["ps1"])) < 1e-10
    for k in range(len(g["par_cases"])):
        pm, pe = g["par_cases"][k]
        assert gc.rel_err(pdf.parallax_lnprior(g["parallaxes"], pm, pe), g["parallax_lnprior_%d" % k]) < 1e-12
        assert gc.rel_err(pdf.scale_parallax_lnprior(g["scales"], g["scale_errs"], pm, pe),
                          g["scale_parallax_lnprior_%d" % k]) < 1e-12


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="needs /root/reference")
def test_live_reference_priors():
    ref_import.import_reference()
    from brutus import pdf as rpdf   # the reference
    rs = np.random.RandomState(3)
    m = 10. ** rs.uniform(-1.3, 1., 500)
    assert gc.rel_err(pdf.imf_lnprior(m), rpdf.imf_lnprior(m)) < 1e-12
    mr = rs.uniform(-3., 24., 500)   # beyond both ends of the table: extrapolated
    assert np.max(np.abs(pdf.ps1_MrLF_lnprior(mr) - rpdf.ps1_MrLF_lnprior(mr))) < 1e-10
    s, se = 10. ** rs.uniform(-3, 1, 300), 10. ** rs.uniform(-4, 0, 300)
    for pm, pe in ((1.3, 0.1), (0.2, 0.1), (np.nan, 0.1), (-0.1, 0.3)):
        assert gc.rel_err(pdf.scale_parallax_lnprior(s, se, pm, pe), rpdf.scale_parallax_lnprior(s, se, pm, pe)) < 1e-12
        assert gc.rel_err(pdf.parallax_lnprior(np.sqrt(s), pm, pe), rpdf.parallax_lnprior(np.sqrt(s), pm, pe)) < 1e-12
