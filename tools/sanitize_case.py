"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck): both seams, the device
posterior and get_seds on a 6 000-model grid."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from brutus_b200 import _lib, mock, fitting
grid, labels = mock.make_grid(6_000, 8, seed=31, kind="locus")
st = mock.make_stars(grid, 40, seed=32, dropout=0.1)
for prec in ("f32", "f64"):
    h = _lib.Handle(0, prec)
    h.set_grid(grid)
    r = h.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], copy=True)
    h.set_model_priors(lnprior=fitting.imf_lnprior(labels["mini"]), feh=labels["feh"], loga=labels["loga"])
    f = h.fit_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], coords=st["coords"],
                    nmc_prior=8, ndraws=16, seed=1)
    o = h.loglike_full(st["flux"][0], st["err"][0], st["mask"][0], st["parallax"][0], st["parallax_err"][0],
                       _lib.make_options())
    s = h.get_seds(np.full(50, 0.3), np.full(50, 3.3), idx=np.arange(50), return_flux=True)
    print(prec, len(r["model_idx"]), float(f["levid"][0]), float(o[1][0]), float(s[0][0, 0]))
    h.close()
