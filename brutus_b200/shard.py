"""Star sharding across GPUs: one process per GPU, contiguous star ranges, the
grid replicated by ONE broadcast at start-up, no collective in the hot loop.

The reference processes stars strictly serially (``for i in range(Ndata)``, brutus/fitting.py:1980)
and every star is independent of every other, so a shard's results do not depend on the number or
placement of shards (tests/test_shard.py checks that with world size 2 on CPU/gloo; the GPU tests
check the same property for batches, tests/test_sweep_gpu.py::test_batch_equals_single_star).

``torch.distributed`` is plumbing only: NCCL moves the grid between GPUs, gloo is used by the CPU
tests.  Nothing here computes.
"""
import numpy as np

__all__ = ["shard_bounds", "broadcast_grid", "gather_catalogue", "fit_shard", "gather_draws"]


def shard_bounds(ndata, world, rank):
    """Contiguous shard ``[lo, hi)`` of a catalogue of ``ndata`` stars for ``rank`` of ``world``;
    shard sizes differ by at most one star and concatenate to the catalogue in rank order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %r outside world of size %r" % (rank, world))
    base, extra = divmod(int(ndata), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_grid(grid, shape, dist=None, src=0, handle=None, device=None):
    """Replicate the (Nmodel, Nfilt, 3) float32 grid held by rank ``src`` on every rank.

    With ``handle`` (a :class:`brutus_b200._lib.Handle`) and a CUDA ``device`` the broadcast runs
    GPU to GPU over NCCL and the received buffer is re-tiled in place on the device
    (``bf_set_grid_device``): the grid crosses PCIe once, on rank ``src``.  Without a handle (CPU,
    gloo) the host array is broadcast and returned.  ``dist`` is ``torch.distributed`` (already
    initialised) or None for a single process."""
    import torch
    shape = tuple(int(x) for x in shape)
    rank = dist.get_rank() if dist is not None else 0
    if rank == src:
        g = np.ascontiguousarray(grid, dtype=np.float32)
        if g.shape != shape:
            raise ValueError("grid shape %r does not match %r" % (g.shape, shape))
    if handle is not None and device is not None:
        if rank == src:
            t = torch.from_numpy(g).to(device)
        else:
            t = torch.empty(shape, dtype=torch.float32, device=device)
        if dist is not None:
            dist.broadcast(t, src=src)
        torch.cuda.synchronize(device)
        from . import _lib
        handle.set_grid_device(t.data_ptr(), shape[0], shape[1], _lib.LAYOUT_C)
        return t
    t = torch.from_numpy(g) if rank == src else torch.empty(shape, dtype=torch.float32)
    if dist is not None:
        dist.broadcast(t, src=src)
    return t.numpy()


def gather_catalogue(local, ndata, dist=None, dst=0):
    """Concatenate per-shard results in catalogue order on rank ``dst``.

    ``local`` is the dict a shard's ``Handle.sweep_batch`` returned (per-star arrays ``ndim, n_iter,
    n_surv, max_lnprob``, CSR ``offsets`` and the record arrays).  Returns the merged dict on
    ``dst`` and None elsewhere.  The hot loop never calls this; it is the final host-side
    gather (object collective over the process group's CPU path)."""
    if dist is None:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = {k: (np.asarray(v) if v is not None else None) for k, v in local.items()}
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for k in ("ndim", "n_iter", "n_surv", "max_lnprob"):
        merged[k] = np.concatenate([o[k] for o in out])
    offs = [np.asarray(o["offsets"], dtype=np.int64) for o in out]
    base = np.cumsum([0] + [int(o[-1]) for o in offs])
    merged["offsets"] = np.concatenate([offs[0][:1]] + [o[1:] + b for o, b in zip(offs, base)])
    for k in ("model_idx", "lnl", "scale", "av", "chi2", "rv"):
        if out[0].get(k) is not None:
            merged[k] = np.concatenate([o[k] for o in out])
    merged["icov6"] = (np.concatenate([o["icov6"] for o in out], axis=1)
                       if out[0].get("icov6") is not None else None)
    if len(merged["ndim"]) != ndata:
        raise RuntimeError("shards do not add up to the catalogue")
    return merged


def fit_shard(handle, data, data_err, data_mask, parallax=None, parallax_err=None, coords=None, world=1, rank=0,
              **fit_kwargs):
    """This rank's shard of the catalogue through ``Handle.fit_batch`` (the per-object body of ``BruteForce.fit``
    on the device).  ``star_base`` is set to the shard's first catalogue index, so every star gets the random
    numbers it would get in a single-process run: the gathered result does not depend on the number of GPUs.
    Returns ``(lo, hi, result)``."""
    lo, hi = shard_bounds(len(data), world, rank)
    cut = lambda a: None if a is None else np.asarray(a)[lo:hi]
    res = handle.fit_batch(cut(data), cut(data_err), cut(data_mask), cut(parallax), cut(parallax_err),
                           coords=cut(coords), star_base=lo, **fit_kwargs)
    return lo, hi, res


def gather_draws(local, ndata, dist=None, dst=0):
    """Concatenate per-shard ``fit_batch`` results (every member is a per-star array) in catalogue order on
    rank ``dst``; None elsewhere.  Final host-side gather, outside the hot loop."""
    if dist is None:
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = {k: np.array(v) for k, v in local.items()}      # copies: the draws may be views of a pinned arena
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst)
    if rank != dst:
        return None
    merged = {k: np.concatenate([o[k] for o in out]) for k in out[0]}
    if len(merged["levid"]) != ndata:
        raise RuntimeError("shards do not add up to the catalogue")
    return merged
