"""CPU-only, world size 2: the multi-GPU host logic (brutus_b200/shard.py) -- contiguous star shards, no
collective in the hot loop, catalogue-ordered gather -- once over the product's own TCP communicator
(SocketComm, no PyTorch) and once over torch.distributed/gloo through a small adapter (test helper).  The
per-shard compute is stood in for by the CPU oracle (test infrastructure), packaged exactly like
``Handle.sweep_batch`` packages its records; the GPU tests check the kernels and the in-library NCCL broadcast."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shard_bounds_cover_catalogue():
    from brutus_b200.shard import shard_bounds
    for ndata in (0, 1, 7, 8, 1000, 100_003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(ndata, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == ndata
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def oracle_sweep(grid, st, lo, hi):
    """What Handle.sweep_batch returns for stars [lo, hi), computed by the CPU oracle."""
    from oracle import oracle
    out = dict(ndim=[], n_iter=[], n_surv=[], max_lnprob=[], offsets=[0], model_idx=[], lnl=[], scale=[],
               av=[], chi2=[], rv=[], icov6=[])
    for i in range(lo, hi):
        pk = dict(parallax=st["parallax"][i], parallax_err=st["parallax_err"][i])
        ref = oracle.loglike(st["flux"][i], st["err"][i], st["mask"][i].copy(), grid, return_vals=True,
                             return_diag=True, **pk)
        lnl, lnprob, sel = oracle.select(ref[0], ref[3], ref[6], **pk)
        out["ndim"].append(ref[1])
        out["n_iter"].append((ref[7]["n_iter_mag"], ref[7]["n_iter_flux"]))
        out["n_surv"].append(ref[7]["n_surv"])
        out["max_lnprob"].append(lnprob.max())
        out["offsets"].append(out["offsets"][-1] + len(sel))
        out["model_idx"].append(sel.astype(np.int32))
        for k, a in (("lnl", lnl), ("scale", ref[3]), ("av", ref[4]), ("chi2", ref[2]), ("rv", ref[5])):
            out[k].append(a[sel])
        ic = ref[6][sel]
        out["icov6"].append(np.stack([ic[:, 0, 0], ic[:, 0, 1], ic[:, 0, 2], ic[:, 1, 1], ic[:, 1, 2], ic[:, 2, 2]]))
    res = {k: np.asarray(out[k]) for k in ("ndim", "n_iter", "n_surv", "max_lnprob", "offsets")}
    res["n_iter"] = res["n_iter"].reshape(-1, 2)
    for k in ("model_idx", "lnl", "scale", "av", "chi2", "rv"):
        res[k] = np.concatenate(out[k]) if out[k] else np.zeros(0)
    res["icov6"] = np.concatenate(out["icov6"], axis=1) if out["icov6"] else np.zeros((6, 0))
    return res


class _FakeFitHandle(object):
    """Stands in for brutus_b200._lib.Handle.fit_batch on the CPU: per-star outputs that depend only on the star's
    data and on its catalogue index (as the real call's counter-based generator does through star_base)."""

    def fit_batch(self, flux, err, mask, parallax=None, parallax_err=None, coords=None, star_base=0, ndraws=4, **kw):
        n = len(flux)
        idx = star_base + np.arange(n)
        rs = [np.random.RandomState(1000 + int(i)) for i in idx]
        draws = np.array([r.randint(0, 3000, ndraws) for r in rs]).reshape(n, ndraws)
        return dict(sidxs=draws.astype(np.int32), levid=np.asarray(flux).sum(axis=1) + idx,
                    dists=np.array([r.uniform(size=ndraws) for r in rs]).reshape(n, ndraws),
                    ndim=np.asarray(mask).sum(axis=1).astype(np.int32))


class GlooComm(object):
    """Adapter: torch.distributed (gloo) behind the interface shard.py expects of a communicator."""

    def __init__(self, dist):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def gather_object(self, obj, dst=0):
        out = [None] * self.world if self.rank == dst else None
        self.dist.gather_object(obj, out, dst=dst)
        return out

    def bcast_object(self, obj, src=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=src)
        return box[0]


def _worker(rank, world, port, q, backend):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    from brutus_b200 import mock
    from brutus_b200.shard import SocketComm, gather_catalogue, shard_bounds
    dist = None
    if backend == "gloo":
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        comm = GlooComm(dist)
    else:
        assert "torch" not in sys.modules, "the product's communicator must not need PyTorch"
        comm = SocketComm(rank, world, "127.0.0.1", port)
    try:
        nmodel, nfilt, ndata = 3000, 6, 9
        # rank 0 owns the grid; the others receive it (on GPUs this is the library's ncclBroadcast, here the
        # communicator's object broadcast stands in for it)
        grid = comm.bcast_object(mock.make_grid(nmodel, nfilt, seed=1600)[0] if rank == 0 else None)
        st = mock.make_stars(mock.make_grid(nmodel, nfilt, seed=1600)[0], ndata, seed=2600, dropout=0.1)
        lo, hi = shard_bounds(ndata, world, rank)
        local = oracle_sweep(grid, st, lo, hi)                                 # hot loop: no communication
        merged = gather_catalogue(local, ndata, comm=comm)
        # the fit-level call: shards carry their first catalogue index (star_base), results gather in order
        from brutus_b200.shard import fit_shard, gather_draws
        flo, fhi, fres = fit_shard(_FakeFitHandle(), st["flux"], st["err"], st["mask"], st["parallax"],
                                   st["parallax_err"], coords=st["coords"], world=world, rank=rank, ndraws=4)
        assert (flo, fhi) == (lo, hi)
        draws = gather_draws(fres, ndata, comm=comm)
        if rank == 0:
            whole = oracle_sweep(grid, st, 0, ndata)
            ok = all(np.array_equal(merged[k], whole[k]) for k in whole)
            one = _FakeFitHandle().fit_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"],
                                             coords=st["coords"], star_base=0, ndraws=4)
            ok = ok and all(np.array_equal(draws[k], one[k]) for k in one)
            q.put(("ok", ok, int(merged["offsets"][-1])))
        else:
            assert merged is None and draws is None
    except Exception as e:  # pragma: no cover
        q.put(("error", repr(e), rank))
        raise
    finally:
        if dist is not None:
            dist.destroy_process_group()
        else:
            comm.close()


@pytest.mark.parametrize("backend", ["socket", "gloo"])
def test_two_rank_shards_equal_single_process(backend):
    import multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, backend)) for r in range(2)]
    for p in procs:
        p.start()
    tag, ok, n = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
    assert tag == "ok" and ok and n > 0, (tag, ok, n)
    assert all(p.exitcode == 0 for p in procs)


def test_product_shard_module_is_torch_free():
    src = open(os.path.join(ROOT, "brutus_b200", "shard.py")).read()
    assert "import torch" not in src
