"""Diagnostic: f32 scale error vs av/rv error on the locus golden case (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import golden_cases as gc
from brutus_b200 import fitting
name = "locus_8band"
grid, labels, st, kw = gc.build_case(name)
gold = gc.load_loglike(name)
for i in range(len(st["flux"])):
    pk = dict(parallax=st["parallax"][i], parallax_err=st["parallax_err"][i])
    out = fitting.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), grid, return_vals=True, precision="f32", **pk, **kw)
    sc, av, rv = out[3], out[4], out[5]
    rsc, rav, rrv = gold["scale_%d" % i], gold["av_%d" % i], gold["rv_%d" % i]
    rel = np.abs(sc - rsc) / np.abs(rsc)
    bad = np.where(rel > 2e-5)[0]
    print("star", i, "nbad", len(bad), "max rel", rel.max(), "max dav", np.abs(av - rav).max(), "max drv", np.abs(rv - rrv).max())
    for b in bad[:12]:
        print("   m=%d rel=%.3g dav=%.3g drv=%.3g av=%.4f rv=%.4f chi2=%.4g scale=%.4g" % (b, rel[b], av[b] - rav[b], rv[b] - rrv[b], rav[b], rrv[b], gold["chi2_%d" % i][b], rsc[b]))
