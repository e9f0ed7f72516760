"""Times bf_fit_batch (device posterior) on the C2 workload; prints stats per configuration."""
import sys, os, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from brutus_b200 import _lib, mock, fitting
nstar = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cfg = mock.CONFIGS[2]
grid, labels = mock.make_grid(cfg["nmodel"], cfg["nfilt"], seed=1002, kind="locus")
st = mock.make_stars(grid, nstar, seed=2002)
h = _lib.Handle(0, "f32")
h.set_grid(grid)
h.set_model_priors(lnprior=fitting.imf_lnprior(labels["mini"]), feh=labels["feh"], loga=labels["loga"])
for nmc in (50, 10):
    for it in range(3):
        h.flush_l2()
        t = time.perf_counter()
        r = h.fit_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], coords=st["coords"],
                        nmc_prior=nmc, ndraws=250, seed=3)
        w = time.perf_counter() - t
        s = h.stats()
    print(json.dumps(dict(nmc=nmc, wall_ms=1e3 * w, stars_per_s=nstar / w, ms_device=s["ms_device"], ms_post=s["ms_post"],
                          ms_magfit=s["ms_magfit"], selected=s["selected"], selected2=s["selected2"],
                          d2h=s["d2h_bytes"], launches=s["kernel_launches"])))
print("levid", r["levid"][:5], "nsel", r["nsel"][:5], "dist med", np.median(r["dists"], axis=1)[:5], "truth", st["truth"]["dist"][:5])
