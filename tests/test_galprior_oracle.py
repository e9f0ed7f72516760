"""The NumPy restatement of the default Galactic prior (oracle/galprior.py) against golden values of
the unmodified reference `brutus.pdf.gal_lnprior` (tests/golden/galprior.npz).  The coordinate
transform is shared with the generator (astropy is absent): see the header of oracle/galprior.py."""
import os

import numpy as np
import pytest

import golden_cases as gc
from oracle import galprior as gp


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(gc.GOLD, "galprior.npz"))


def _labels(gold, names):
    lab = np.zeros(len(gold["dists"]), dtype=[(n, "f8") for n in names])
    for n in names:
        lab[n] = gold[n]
    return lab


@pytest.mark.parametrize("kind,names", [("full", ("feh", "loga")), ("feh_only", ("feh",)), ("nolabels", ())])
def test_gal_lnprior_matches_reference(gold, kind, names):
    lab = _labels(gold, names) if names else None
    for k, c in enumerate(gold["coords"]):
        ref = gold["%s_%d" % (kind, k)]
        out = gp.gal_lnprior(gold["dists"], tuple(c), labels=lab)
        assert np.array_equal(np.isfinite(ref), np.isfinite(out))
        fin = np.isfinite(ref)
        assert fin.sum() > 300
        assert np.max(np.abs(out[fin] - ref[fin])) < 1e-10


def test_sun_position():
    R, Z = gp.galactic_to_cyl(np.array([0.]), (0., 0.))
    assert abs(Z[0] - gp.Z_SUN) < 1e-12 and abs(np.hypot(R[0], Z[0]) - gp.GALCEN_DISTANCE) < 1e-12
    # 8.122 kpc towards the Galactic centre lands (almost) on the centre
    R, Z = gp.galactic_to_cyl(np.array([gp.GALCEN_DISTANCE]), (0., 0.))
    assert R[0] < 0.03 and abs(Z[0]) < 0.03


def test_replay_rstate_matches_numpy_choice():
    """ReplayRState.choice reproduces numpy's RandomState.choice given the same uniforms."""
    rs = np.random.RandomState(5)
    p = rs.uniform(size=37)
    p /= p.sum()
    st = np.random.RandomState(9)
    want = st.choice(37, size=50, p=p)
    u = np.random.RandomState(9).random_sample(50)
    rr = gp.ReplayRState(np.zeros((1, 3, 1)), [np.array([0])], u[None, :], u[None, :])
    rr.normal(size=3)
    assert np.array_equal(rr.choice(37, size=50, p=p), want)
