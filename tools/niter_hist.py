"""GPU box: histogram of mag / flux iteration counts and per-star candidate statistics for a bench config."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brutus_b200 import _lib, mock  # noqa: E402

cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nstar = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfg = mock.CONFIGS[cfg_id]
grid, labels = mock.make_grid(cfg["nmodel"], cfg["nfilt"], seed=1000 + cfg_id, kind="locus")
st = mock.make_stars(grid, nstar, seed=2000 + cfg_id, av_max=cfg["av_max"], dropout=cfg["dropout"])
h = _lib.Handle(0, "f32")
h.set_grid(grid)
res = h.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"],
                    opts=_lib.make_options(avlim=cfg["avlim"], skip_d2h=True))
print("mag iterations :", np.bincount(res["n_iter"][:, 0]))
print("flux iterations:", np.bincount(res["n_iter"][:, 1]))
ns = res["n_surv"]
print("survivors per star: mean %.0f median %.0f max %d" % (ns.mean(), np.median(ns), ns.max()))
nsel = np.diff(res["offsets"])
print("selected per star: mean %.0f median %.0f max %d" % (nsel.mean(), np.median(nsel), nsel.max()))
more = res["n_iter"][:, 1] > 2
print("stars with > 2 flux iterations: %d, their survivors: %d of %d" % (more.sum(), ns[more].sum(), ns.sum()))
print({k: v for k, v in h.stats().items() if v})
