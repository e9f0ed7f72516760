"""Diagnostic (GPU box): per-model comparison of bf_sweep_batch records with the oracle for one star of a
BASELINE config; prints the worst outliers with the oracle's survivor flag and distance to the thresholds.

    python tools/diag_sweep.py <cfg> <star> [nmodel]
"""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brutus_b200 import _lib, mock  # noqa: E402
from oracle import oracle  # noqa: E402

cfg_id, star = int(sys.argv[1]), int(sys.argv[2])
cfg = dict(mock.CONFIGS[cfg_id])
if len(sys.argv) > 3:
    cfg["nmodel"] = int(sys.argv[3])
grid, labels = mock.make_grid(cfg["nmodel"], cfg["nfilt"], seed=1000 + cfg_id, kind="locus")
st = mock.make_stars(grid, max(star + 1, 10), seed=3000 + cfg_id, av_max=cfg["av_max"], dropout=cfg["dropout"])
h = _lib.Handle(0, "f32")
h.set_grid(grid)
opts = _lib.make_options(avlim=cfg["avlim"])
sl = slice(star, star + 1)
res = h.sweep_batch(st["flux"][sl], st["err"][sl], st["mask"][sl], st["parallax"][sl], st["parallax_err"][sl],
                    opts=opts, copy=True)
print("stats", {k: v for k, v in h.stats().items() if isinstance(v, int) and v})
pk = dict(parallax=st["parallax"][star], parallax_err=st["parallax_err"][star])
ref = oracle.loglike(st["flux"][star], st["err"][star], st["mask"][star].copy(), grid, return_vals=True,
                     return_diag=True, avlim=cfg["avlim"], **pk)
_, lnprob, sel = oracle.select(ref[0], ref[3], ref[6], **pk)
surv = ref[7]["survivors"]
print("mask", st["mask"][star].astype(int), "par", pk, "n_iter dev", res["n_iter"][0], "ref",
      ref[7]["n_iter_mag"], ref[7]["n_iter_flux"], "nsurv dev", res["n_surv"][0], "ref", ref[7]["n_surv"],
      "nsel dev", len(res["model_idx"]), "ref", len(sel))
common, ia, ib = np.intersect1d(res["model_idx"], sel, return_indices=True)
d = np.abs(res["chi2"][ia] - ref[2][common])
order = np.argsort(d)[::-1][:15]
print("outliers: model dchi2 chi2_dev chi2_ref av_dev av_ref rv_dev rv_ref surv_ref")
for k in order:
    m = common[k]
    print(m, d[k], res["chi2"][ia][k], ref[2][m], res["av"][ia][k], ref[4][m], res["rv"][ia][k], ref[5][m], bool(surv[m]))
print("n outliers > 3e-3:", int((d > 3e-3).sum()), "of", len(d), "| of which oracle survivors:",
      int(((d > 3e-3) & surv[common]).sum()))
h.close()
