"""GPU box: detailed device-vs-oracle comparison for individual stars (debugging aid)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from brutus_b200 import _lib, mock  # noqa: E402
from oracle import oracle  # noqa: E402
import parity  # noqa: E402


def report(tag, grid, st, i, **kw):
    ref, lnl, lnprob, sel = parity.oracle_star(oracle, grid, st, i, **kw)
    d = ref[7]
    print("==", tag, "oracle n_iter", d["n_iter_mag"], d["n_iter_flux"], "nsurv", d["n_surv"], "nsel", len(sel),
          "ndim", ref[1], "flux", st["flux"][i], "snr", st["flux"][i] / st["err"][i], "par", st["parallax"][i], st["parallax_err"][i])
    for prec in ("f64", "f32"):
        h = _lib.Handle(0, prec)
        h.set_grid(grid)
        sl = slice(i, i + 1)
        res = h.sweep_batch(st["flux"][sl], st["err"][sl], st["mask"][sl], st["parallax"][sl], st["parallax_err"][sl],
                            opts=_lib.make_options(**kw), copy=True)
        print("  ", prec, "n_iter", res["n_iter"][0], "nsurv", res["n_surv"][0], "nsel", len(res["model_idx"]),
              "max_lnprob", res["max_lnprob"][0], "ref", lnprob.max(), {k: v for k, v in h.stats().items() if k in ("fixups", "flux_more_launches", "resweeps")})
        try:
            parity.check_star(res, 0, ref, lnl, lnprob, sel, prec, tag=tag)
            print("     check OK")
        except AssertionError as e:
            print("     check FAILED:", str(e)[:300])
        if prec == "f32" and os.environ.get("DIAG_DETAIL"):
            rec = parity.star_records(res, 0)
            common, ia, ib = np.intersect1d(rec["model_idx"], sel, return_indices=True)
            dl = rec["lnl"][ia] - lnl[common]
            print("     lnl diff: min %.4g max %.4g median %.4g; chi2 range %.4g..%.4g" %
                  (dl.min(), dl.max(), np.median(dl), ref[2][common].min(), ref[2][common].max()))
            top = common[np.argsort(-lnprob[common])[:6]]
            for m in top:
                k = int(np.where(rec["model_idx"] == m)[0][0])
                print("     model %d oracle lnprob %.5f lnl %.5f chi2 %.5f av %.5f rv %.5f s %.6g | dev lnl %.5f chi2 %.5f av %.5f rv %.5f s %.6g" %
                      (m, lnprob[m], lnl[m], ref[2][m], ref[4][m], ref[5][m], ref[3][m], rec["lnl"][k], rec["chi2"][k],
                       rec["av"][k], rec["rv"][k], rec["scale"][k]))
            w = np.argsort(-np.abs(dl))[:4]
            for j in w:
                m = common[j]; k = ia[j]
                print("     worst %d oracle lnl %.5f chi2 %.5f av %.5f rv %.5f | dev lnl %.5f chi2 %.5f av %.5f rv %.5f" %
                      (m, lnl[m], ref[2][m], ref[4][m], ref[5][m], rec["lnl"][k], rec["chi2"][k], rec["av"][k], rec["rv"][k]))
        h.close()


which = sys.argv[1:] or ["fuzz", "c4"]
if "fuzz" in which:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import test_edge_cases_gpu as te
    grid, labels = mock.make_grid(20_000, 8, seed=1700, kind="locus")
    for (seed, nmask, nneg, pm, avhi) in [(0, 0, 3, "none", 20.), (0, 0, 0, "none", 20.), (1, 1, 2, "good", 6.)]:
        st = te._star(grid, seed, nneg, nmask, pm, avhi)
        report(("fuzz", seed, nmask, nneg, pm, avhi), grid, st, 0, avlim=(0., avhi))
if "c4" in which:
    grid, labels = mock.make_grid_lattice()
    st = mock.load_ngc2682()
    want = [int(x) for x in os.environ.get("C4_INDEX", "65").split(",")]
    for i in [int(np.where(st["index"] == w)[0][0]) for w in want]:
        report(("c4", i, int(st["index"][i])), grid, st, i)
