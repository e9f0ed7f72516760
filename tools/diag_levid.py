"""GPU box: spread of the evidence over generator seeds, device posterior vs host posterior, on the five NGC 2682
objects tests/test_ngc2682_gpu.py compares (is a device/host difference Monte Carlo noise or a bias?)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from brutus_b200 import fitting  # noqa: E402
from oracle import galprior as gp  # noqa: E402
import gen_golden  # noqa: E402
import golden_cases as gc  # noqa: E402

d = np.load(os.path.join(gc.GOLD, "ngc2682.npz"))
grid, labels = gen_golden.ngc2682_grid()
ok = d["mask"].sum(axis=1) >= 4
idx = np.where(ok)[0]
rs = np.random.RandomState(4)
rs.choice(len(idx), 12, replace=False)
pick = idx[rs.choice(len(idx), 5, replace=False)]
lmask = np.ones(1, dtype=[(n, bool) for n in labels.dtype.names])
bf = fitting.BruteForce(grid, labels, lmask)
nmc = int(os.environ.get("NMC", "200"))
kw = dict(parallax=d["parallax"][pick], parallax_err=d["parallax_err"][pick], Nmc_prior=nmc, Ndraws=300, dustfile=None,
          lnprior=fitting.imf_lnprior(labels["mini"]), data_coords=d["coords"][pick])
args = (d["phot"][pick], d["err"][pick], d["mask"][pick])
dev = np.array([[r[7] for r in bf._fit(*args, rstate=np.random.RandomState(s), **kw)] for s in range(8)])
host = np.array([[r[7] for r in bf._fit(*args, rstate=np.random.RandomState(100 + s),
                                        lngalprior=lambda dd, c, labels=None: gp.gal_lnprior(dd, c, labels=labels), **kw)]
                 for s in range(4)])
bf.close()
np.set_printoptions(precision=4, suppress=True, linewidth=200)
print("nmc", nmc)
print("dev  mean", dev.mean(axis=0), "std", dev.std(axis=0))
print("host mean", host.mean(axis=0), "std", host.std(axis=0))
print("diff of means", dev.mean(axis=0) - host.mean(axis=0))
