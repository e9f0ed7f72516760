"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck): both seams, the device
posterior, get_seds, the per-model start of the magnitude fit (bf_set_init) and photometric_offsets' device part
on a 6 000-model grid."""
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
    rs = np.random.RandomState(5)
    h.set_init(rs.uniform(0., 2., grid.shape[0]), rs.uniform(2.8, 4., grid.shape[0]))
    oi = h.loglike_full(st["flux"][1], st["err"][1], st["mask"][1], st["parallax"][1], st["parallax_err"][1],
                        _lib.make_options())
    h.set_init()
    nobj, nsamp = 12, 9
    ix = rs.randint(0, grid.shape[0], size=(nobj, nsamp))
    sd, wt = h.offsets_weights(st["flux"][:nobj], st["err"][:nobj], st["mask"][:nobj], ix, rs.uniform(0, 1, ix.shape),
                               rs.uniform(3., 3.6, ix.shape), rs.uniform(0.5, 2., ix.shape))
    print(prec, len(r["model_idx"]), float(f["levid"][0]), float(o[1][0]), float(s[0][0, 0]), float(oi[1][0]),
          float(sd[0, 0, 0]), float(np.nansum(wt)))
    h.close()
