"""GPU box: one fuzz case in float64, worst av/rv/chi2 mismatches with the oracle's survivor flags (debugging aid)."""
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

seed, nmask, nneg, pm, avhi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], float(sys.argv[5])
grid, labels = mock.make_grid(20_000, 8, seed=1700, kind="locus")
st = te._star(grid, seed, nneg, nmask, pm, avhi)
kw = dict(avlim=(0., avhi))
ref, lnl, lnprob, sel = parity.oracle_star(oracle, grid, st, 0, **kw)
h = _lib.Handle(0, "f64")
h.set_grid(grid)
res = h.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], opts=_lib.make_options(**kw), copy=True)
print("n_iter", res["n_iter"][0], ref[7]["n_iter_mag"], ref[7]["n_iter_flux"], "nsurv", res["n_surv"][0], ref[7]["n_surv"], "sets equal", np.array_equal(res["model_idx"], sel))
idx = res["model_idx"]
d = np.abs(res["av"] - ref[4][idx])
o = np.argsort(d)[::-1][:10]
lp = ref[7]["lnl_p"]; thr = lp.max() + np.log(5e-3)
for k in o:
    m = idx[k]
    print(m, "dav", d[k], "av", res["av"][k], ref[4][m], "rv", res["rv"][k], ref[5][m], "chi2", res["chi2"][k], ref[2][m], "surv", bool(ref[7]["survivors"][m]), "lp-thr", lp[m] - thr)
print("n bad av > 1e-8:", int((d > 1e-8).sum()))
