"""GPU: several GPUs inside the library.  A star's results must not depend on
how many devices share the catalogue -- bitwise, since no arithmetic crosses stars and the posterior's random numbers
are keyed by the catalogue index.

* one process, n devices (bf_create_multi): the library broadcasts the grid itself (ncclCommInitAll + one
  ncclBroadcast) and shards the batch calls over host threads;
* one process per GPU: two spawned processes join an NCCL group inside the library (bf_nccl_init; the id travels
  over brutus_b200.shard.SocketComm), rank 0 alone holds the grid.
Both need >= 2 visible GPUs (`gpurun --gpus 2`); with one GPU only the degenerate cases run."""
import os
import socket
import sys

import numpy as np
import pytest

from brutus_b200 import mock

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _ngpu():
    from brutus_b200 import _lib
    return _lib.load().bf_device_count()


def _case():
    grid, labels = mock.make_grid(60_000, 8, seed=1800, kind="locus")
    st = mock.make_stars(grid, 37, seed=2800, dropout=0.05)
    return grid, labels, st


def _fit(h, st, labels, sl=slice(None), star_base=0):
    from brutus_b200 import pdf
    h.set_model_priors(lnprior=pdf.imf_lnprior(labels["mini"]), feh=labels["feh"], loga=labels["loga"])
    return h.fit_batch(st["flux"][sl], st["err"][sl], st["mask"][sl], st["parallax"][sl], st["parallax_err"][sl],
                       coords=st["coords"][sl], nmc_prior=20, ndraws=40, seed=99, star_base=star_base, mem_lim=8000.)


def test_single_device_multi_handle_and_world_of_one():
    from brutus_b200 import _lib, shard
    grid, labels, st = _case()
    a = _lib.Handle(0, "f32")
    b = _lib.Handle([0], "f32")
    try:
        a.set_grid(grid)
        shard.init_process_group(b, shard.SocketComm(0, 1))     # world of one: no NCCL needed
        shard.broadcast_grid(b, grid, grid.shape)
        ra, rb = _fit(a, st, labels), _fit(b, st, labels)
        for k in ra:
            assert np.array_equal(ra[k], rb[k]), k
        assert np.array_equal(b.allreduce_max([1.5, -2.0]), [1.5, -2.0])
    finally:
        a.close()
        b.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_two_devices_one_process_equal_one_device(precision):
    from brutus_b200 import _lib
    grid, labels, st = _case()
    one = _lib.Handle(0, precision)
    two = _lib.Handle([0, 1], precision)
    try:
        one.set_grid(grid)
        two.set_grid(grid)                      # H2D once, ncclBroadcast to the second device
        ra, rb = _fit(one, st, labels), _fit(two, st, labels)
        for k in ra:
            assert np.array_equal(ra[k], rb[k]), k
        sa = one.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], copy=True)
        sb = two.sweep_batch(st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"], copy=True)
        for k in ("offsets", "model_idx", "lnl", "chi2", "scale", "av", "rv", "icov6", "ndim", "n_iter", "n_surv", "max_lnprob"):
            assert np.array_equal(sa[k], sb[k]), k
        assert two.stats()["kernel_launches"] > one.stats()["kernel_launches"]
    finally:
        one.close()
        two.close()


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    from brutus_b200 import _lib, shard
    assert "torch" not in sys.modules
    grid, labels, st = _case()
    comm = shard.SocketComm(rank, world, "127.0.0.1", port)
    h = _lib.Handle(rank, "f32")
    try:
        shard.init_process_group(h, comm)
        shard.broadcast_grid(h, grid if rank == 0 else None, grid.shape)       # only rank 0 uploads
        from brutus_b200 import pdf
        shard.broadcast_model_priors(h, grid.shape[0], **(dict(lnprior=pdf.imf_lnprior(labels["mini"]), feh=labels["feh"],
                                                               loga=labels["loga"]) if rank == 0 else {}))
        lo, hi, res = shard.fit_shard(h, st["flux"], st["err"], st["mask"], st["parallax"], st["parallax_err"],
                                      coords=st["coords"], world=world, rank=rank, nmc_prior=20, ndraws=40, seed=99,
                                      mem_lim=8000.)
        t = h.allreduce_max([float(rank), -float(rank)])
        merged = shard.gather_draws(res, len(st["flux"]), comm=comm)
        if rank == 0:
            q.put(("ok", merged, t.tolist()))
    except Exception as e:  # pragma: no cover
        q.put(("error", repr(e), rank))
        raise
    finally:
        h.close()
        comm.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_processes_nccl_group_equal_one_device():
    import multiprocessing as mp
    from brutus_b200 import _lib
    grid, labels, st = _case()
    one = _lib.Handle(0, "f32")
    try:
        one.set_grid(grid)
        ref = _fit(one, st, labels)
    finally:
        one.close()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tag, merged, t = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert tag == "ok", merged
    assert t == [1.0, 0.0]
    for k in ref:
        assert np.array_equal(ref[k], merged[k]), k
