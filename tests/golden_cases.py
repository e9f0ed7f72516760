"""Shared access to the golden fixtures (tests/golden/*.npz, produced by tests/gen_golden.py)."""
import os

import numpy as np

import gen_golden

GOLD = gen_golden.GOLD
LOGLIKE_CASES = gen_golden.LOGLIKE_CASES
FIT_CASE = gen_golden.FIT_CASE
toy_galprior = gen_golden.toy_galprior
build_case = gen_golden.build_case
KEYS = ("lnl", "ndim", "chi2", "scale", "av", "rv", "icov")


def fresh(kw):
    """Copies of the array-valued keywords (av_init, rv_init): loglike fits IN those arrays (brutus/fitting.py:202,
    :232, :809), so every call gets its own."""
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}


def load_loglike(name):
    return np.load(os.path.join(GOLD, "loglike_%s.npz" % name))


def load_fit():
    return np.load(os.path.join(GOLD, "fit_generator.npz"))


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over finite entries; also checks the non-finite pattern."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(fin, np.isfinite(a)), "non-finite pattern differs"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)))
