"""Star sharding across GPUs: contiguous star ranges, the grid replicated by ONE NCCL
broadcast at start-up, no collective in the hot loop.

The reference processes stars strictly serially (``for i in range(Ndata)``, brutus/fitting.py:1980) and every
star is independent of every other, so a shard's results do not depend on the number or placement of shards.

Two ways to use several GPUs, neither needs PyTorch:

* one process, ``Handle([0, 1, ...])`` / ``BruteForce(..., device=[0, 1, ...])``: the library shards the stars of
  every batch call over its devices (``bf_create_multi``, host threads) and broadcasts the grid itself;
* one process per GPU (``torchrun`` or any launcher that sets RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT):
  :func:`init_process_group` joins the ranks' handles into an NCCL group *inside the library*
  (``bf_nccl_init``); the 128-byte NCCL id travels over :class:`SocketComm`, a few lines of TCP that also carry
  the final host-side gather of per-rank results (pickled NumPy dictionaries: bookkeeping, not compute).

Nothing here computes.  ``comm`` arguments accept any object with ``rank``, ``world`` and
``gather_object(obj, dst)`` (the tests also pass an adapter over ``torch.distributed``/gloo).
"""
import os
import pickle
import socket
import struct
import time

import numpy as np

__all__ = ["shard_bounds", "SocketComm", "init_process_group", "broadcast_grid", "broadcast_model_priors",
           "gather_catalogue", "fit_shard", "gather_draws"]


def shard_bounds(ndata, world, rank):
    """Contiguous shard ``[lo, hi)`` of a catalogue of ``ndata`` stars for ``rank`` of ``world``;
    shard sizes differ by at most one star and concatenate to the catalogue in rank order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %r outside world of size %r" % (rank, world))
    base, extra = divmod(int(ndata), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _send(sock, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock):
    def read(n):
        buf = bytearray()
        while len(buf) < n:
            chunk = sock.recv(min(n - len(buf), 1 << 20))
            if not chunk:
                raise ConnectionError("peer closed the connection")
            buf += chunk
        return bytes(buf)
    (n,) = struct.unpack("<Q", read(8))
    return pickle.loads(read(n))


class SocketComm(object):
    """Star-shaped TCP rendezvous between the ranks of one job: rank 0 listens on ``(addr, port)``, every other
    rank connects.  Carries control-plane traffic only: the NCCL id, and the gather of per-rank results."""

    def __init__(self, rank, world, addr="127.0.0.1", port=29517, timeout=300.):
        self.rank, self.world = int(rank), int(world)
        self.peers = {}
        self.sock = None
        if self.world == 1:
            return
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(self.world)
            srv.settimeout(timeout)
            while len(self.peers) < self.world - 1:
                c, _ = srv.accept()
                c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                c.settimeout(timeout)
                self.peers[_recv(c)] = c
            srv.close()
        else:
            t0 = time.time()
            while True:
                try:
                    s = socket.create_connection((addr, port), timeout=timeout)
                    break
                except OSError:
                    if time.time() - t0 > timeout:
                        raise
                    time.sleep(0.05)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            s.settimeout(timeout)
            _send(s, self.rank)
            self.sock = s

    @classmethod
    def from_env(cls, port_offset=17, **kw):
        """Ranks from RANK / WORLD_SIZE, address from MASTER_ADDR, port MASTER_PORT + ``port_offset`` (the
        launcher's own store owns MASTER_PORT)."""
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                   os.environ.get("MASTER_ADDR", "127.0.0.1"),
                   int(os.environ.get("MASTER_PORT", "29500")) + port_offset, **kw)

    def bcast_object(self, obj, src=0):
        if self.world == 1:
            return obj
        if src != 0:
            obj = self.gather_object(obj if self.rank == src else None, dst=0)
            obj = obj[src] if self.rank == 0 else None
        if self.rank == 0:
            for r in sorted(self.peers):
                _send(self.peers[r], obj)
            return obj
        return _recv(self.sock)

    def gather_object(self, obj, dst=0):
        """List of every rank's object on rank ``dst`` (rank order), None elsewhere."""
        if self.world == 1:
            return [obj]
        if dst != 0:
            raise ValueError("SocketComm gathers on rank 0")
        if self.rank == 0:
            out = [obj] + [None] * (self.world - 1)
            for r in sorted(self.peers):
                out[r] = _recv(self.peers[r])
            return out
        _send(self.sock, obj)
        return None

    def barrier(self):
        self.bcast_object(self.gather_object(None) is not None)

    def close(self):
        for c in self.peers.values():
            c.close()
        if self.sock is not None:
            self.sock.close()
        self.peers, self.sock = {}, None


def init_process_group(handle, comm):
    """Join this rank's one-device ``handle`` into the job's NCCL group (``ncclCommInitRank`` inside the library);
    the id is created on rank 0 and distributed over ``comm``."""
    if comm.world == 1:
        return
    uid = comm.bcast_object(handle.nccl_unique_id() if comm.rank == 0 else None)
    handle.nccl_init(uid, comm.rank, comm.world)


def broadcast_grid(handle, grid, shape, root=0):
    """Replicate the (Nmodel, Nfilt, 3) float32 grid held by rank ``root`` on every rank's GPU: one H2D copy on
    the root, ONE ``ncclBroadcast`` GPU to GPU, device-side re-tiling everywhere (``bf_set_grid_bcast``).  With a
    single rank this is ``set_grid``."""
    if handle.world == 1:
        handle.set_grid(grid)
    else:
        handle.set_grid_bcast(grid, shape, root=root)


def broadcast_model_priors(handle, nmodel, lnprior=None, feh=None, loga=None, root=0):
    """Stage the static per-model inputs of ``lnpost`` (brutus/fitting.py:1004, brutus/pdf.py:669, :694) on every
    rank from the arrays rank ``root`` holds: (3, Nmodel) float64 over NCCL."""
    have = np.zeros(3, dtype=np.int64)
    buf = np.zeros((3, int(nmodel)))
    if handle.rank == root:
        for k, a in enumerate((lnprior, feh, loga)):
            if a is not None:
                have[k] = 1
                buf[k] = np.asarray(a, dtype=np.float64)
    if handle.world > 1:
        have = handle.bcast_array(have, root)
        buf = handle.bcast_array(buf, root)
    pri = dict(zip(("lnprior", "feh", "loga"), (buf[k] if have[k] else None for k in range(3))))
    handle.set_model_priors(**pri)
    return pri


def gather_catalogue(local, ndata, comm=None, dst=0):
    """Concatenate per-shard results in catalogue order on rank ``dst``.

    ``local`` is the dict a shard's ``Handle.sweep_batch`` returned (per-star arrays ``ndim, n_iter,
    n_surv, max_lnprob``, CSR ``offsets`` and the record arrays).  Returns the merged dict on
    ``dst`` and None elsewhere.  The hot loop never calls this; it is the final host-side gather."""
    if comm is None or comm.world == 1:
        return local
    payload = {k: (np.asarray(v) if v is not None else None) for k, v in local.items()}
    out = comm.gather_object(payload, dst=dst)
    if comm.rank != dst:
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


def gather_draws(local, ndata, comm=None, dst=0):
    """Concatenate per-shard ``fit_batch`` results (every member is a per-star array) in catalogue order on
    rank ``dst``; None elsewhere.  Final host-side gather, outside the hot loop."""
    if comm is None or comm.world == 1:
        return local
    payload = {k: np.array(v) for k, v in local.items()}      # copies: the draws may be views of a pinned arena
    out = comm.gather_object(payload, dst=dst)
    if comm.rank != dst:
        return None
    merged = {k: np.concatenate([o[k] for o in out]) for k in out[0]}
    if len(merged["levid"]) != ndata:
        raise RuntimeError("shards do not add up to the catalogue")
    return merged
