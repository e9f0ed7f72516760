"""GPU box: sweep a grid of fuzz parameters, print the parity failures compactly (debugging aid)."""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from brutus_b200 import _lib, mock  # noqa: E402
from oracle import oracle  # noqa: E402
import parity  # noqa: E402
import test_edge_cases_gpu as te  # noqa: E402

grid, labels = mock.make_grid(20_000, 8, seed=1700, kind="locus")
hs = {p: _lib.Handle(0, p) for p in ("f64", "f32")}
for h in hs.values():
    h.set_grid(grid)
nfail = {"f64": 0, "f32": 0}
ntot = 0
for seed, nmask, nneg, pm, avhi in itertools.product(range(int(sys.argv[1]) if len(sys.argv) > 1 else 6), (0, 2, 4), (0, 1, 3, 5, 8),
                                                      ("none", "negative", "lowsnr", "good"), (20., 0.3)):
    st = te._star(grid, seed, nneg, nmask, pm, avhi)
    if st["mask"].sum() < 4 or int(((st["flux"][0] > 0) & st["mask"][0]).sum()) == 1:
        continue
    kw = dict(avlim=(0., avhi))
    ref, lnl, lnprob, sel = parity.oracle_star(oracle, grid, st, 0, **kw)
    ntot += 1
    for prec in ("f64", "f32"):
        res = hs[prec].sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"],
                                   opts=_lib.make_options(**kw), copy=True)
        try:
            parity.check_star(res, 0, ref, lnl, lnprob, sel, prec, tag=(seed, nmask, nneg, pm, avhi, prec))
        except AssertionError as e:
            nfail[prec] += 1
            print("FAIL", str(e)[:260].replace("\n", " "), "| ref n_iter", ref[7]["n_iter_mag"], ref[7]["n_iter_flux"], "dev", res["n_iter"][0],
                  "nsel", len(sel), len(res["model_idx"]), "maxlnprob %.2f" % lnprob.max())
print("cases", ntot, "failures", nfail)
