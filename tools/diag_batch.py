"""GPU box: does a star's (mag, flux) iteration count depend on the batch it is swept in?  (debugging aid)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brutus_b200 import _lib, mock  # noqa: E402

grid, labels = mock.make_grid_lattice()
st = mock.load_ngc2682()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
h = _lib.Handle(0, "f32")
h.set_grid(grid)
args = lambda sl: (st["flux"][sl], st["err"][sl], st["mask"][sl], st["parallax"][sl], st["parallax_err"][sl])
full = h.sweep_batch(*args(slice(0, n)), copy=True)
print("batch stats", {k: v for k, v in h.stats().items() if k in ("resweeps", "fixups", "flux_more_launches", "fallbacks", "regroups", "candidates")})
single = np.zeros((n, 2), dtype=int)
nsurv1 = np.zeros(n, dtype=int)
for i in range(n):
    r = h.sweep_batch(*args(slice(i, i + 1)))
    single[i] = r["n_iter"][0]
    nsurv1[i] = r["n_surv"][0]
bad = np.where((single != full["n_iter"]).any(axis=1))[0]
print("stars whose n_iter differs between batch and single:", len(bad), bad[:40])
for i in bad[:12]:
    print(i, "batch", full["n_iter"][i], full["n_surv"][i], "single", single[i], nsurv1[i])
print("mag K histogram (single):", np.bincount(single[:, 0]), "batch:", np.bincount(full["n_iter"][:, 0]))
for sb in ("32", "64"):
    os.environ["BRUTUS_B200_SHIP_BATCH"] = sb
    r = h.sweep_batch(*args(slice(0, n)), copy=True)
    bad2 = np.where((single != r["n_iter"]).any(axis=1))[0]
    print("ship batch", sb, "mismatches", len(bad2), bad2[:20])
h.close()
