"""A ``.npz`` archive whose member arrays are written IN PLACE, batch by batch, through a memory map.

``BruteForce.fit`` writes its results while it runs (the reference's ``running_io``, brutus/fitting.py:1594-1601,
which re-opens its HDF5 file after every object).  Without h5py the container here is NumPy's own ``.npz``: a ZIP
archive with stored (uncompressed) members.  The shapes and dtypes of all members are known before the first
object is fitted, so the archive is laid out once, filled with the reference's initial values, and mapped; rows
then land directly in the file and nothing is copied or re-packed at the end.  A ZIP
member carries a CRC-32 of its contents, in its local header and in the central directory: both are recomputed
and patched when the fit completes.  Until then the stored checksums are those of the initial fill, and a file
left behind by an interrupted fit is read with :func:`read_partial`, which skips the check (``row_done`` says
which rows had been written)."""
import io
import mmap
import struct
import time
import zlib

import numpy as np

__all__ = ["NpzInPlace", "read_partial"]

_ALIGN = 64
_PAD_ID = 0x706e          # private extra-field id of the alignment block


def _npy_header(shape, dtype):
    buf = io.BytesIO()
    np.lib.format.write_array_header_1_0(buf, {"descr": np.lib.format.dtype_to_descr(np.dtype(dtype)),
                                               "fortran_order": False, "shape": tuple(shape)})
    return buf.getvalue()


def _members(mm):
    """name -> (offset of the local header, offset of the member's data, stored size, offset of its central record)."""
    eocd = mm.rfind(b"PK\x05\x06")
    if eocd < 0:
        raise ValueError("not a ZIP archive")
    nrec, cd_size, cd_off = struct.unpack_from("<HII", mm, eocd + 10)
    loc = mm.rfind(b"PK\x06\x07", 0, eocd)                    # zip64 locator -> zip64 end record
    if loc >= 0:
        (e64,) = struct.unpack_from("<Q", mm, loc + 8)
        nrec, cd_size, cd_off = struct.unpack_from("<QQQ", mm, e64 + 32)
    out = {}
    p = cd_off
    for _ in range(nrec):
        if mm[p:p + 4] != b"PK\x01\x02":
            raise ValueError("bad central directory")
        csize, usize, nlen, elen, clen = struct.unpack_from("<IIHHH", mm, p + 20)
        (hoff,) = struct.unpack_from("<I", mm, p + 42)
        name = bytes(mm[p + 46:p + 46 + nlen]).decode()
        q, end = p + 46 + nlen, p + 46 + nlen + elen
        while q + 4 <= end:                                   # zip64 extra: sizes / offset that did not fit 32 bits
            eid, esz = struct.unpack_from("<HH", mm, q)
            if eid == 1:
                vals = list(struct.unpack_from("<%dQ" % (esz // 8), mm, q + 4))
                if usize == 0xFFFFFFFF:
                    usize = vals.pop(0)
                if csize == 0xFFFFFFFF:
                    csize = vals.pop(0)
                if hoff == 0xFFFFFFFF:
                    hoff = vals.pop(0)
            q += 4 + esz
        lnlen, lelen = struct.unpack_from("<HH", mm, hoff + 26)
        out[name] = (hoff, hoff + 30 + lnlen + lelen, usize, p)
        p += 46 + nlen + elen + clen
    return out


def _dos_now():
    t = time.localtime()
    return (t.tm_hour << 11) | (t.tm_min << 5) | (t.tm_sec // 2), ((max(t.tm_year, 1980) - 1980) << 9) | (t.tm_mon << 5) | t.tm_mday


def _local_header(name, size, crc, pad, when):
    # sizes always in a zip64 extra field (any member may exceed 4 GiB on a large catalogue), then the alignment block
    extra = struct.pack("<HHQQ", 1, 16, size, size) + struct.pack("<HH", _PAD_ID, pad) + b"\0" * pad
    return struct.pack("<4sHHHHHIIIHH", b"PK\x03\x04", 45, 0, 0, when[0], when[1], crc, 0xFFFFFFFF, 0xFFFFFFFF,
                       len(name), len(extra)) + name + extra


def _central_record(name, size, crc, hoff, when):
    extra = struct.pack("<HHQQQ", 1, 24, size, size, hoff)
    return struct.pack("<4sHHHHHHIIIHHHHHII", b"PK\x01\x02", 45, 45, 0, 0, when[0], when[1], crc, 0xFFFFFFFF,
                       0xFFFFFFFF, len(name), len(extra), 0, 0, 0, 0, 0xFFFFFFFF) + name + extra


class NpzInPlace(object):
    """``spec``: name -> (shape, dtype, fill value) of the members written in place; ``fixed``: name -> array,
    stored once.  ``arrays[name]`` are views of the file.  The file is created exclusively (an existing one raises
    ``FileExistsError``).  The archive is laid out directly (stored members, zip64 records throughout): the data
    regions start as holes of a sparse file, so that only non-zero fill values cost anything."""

    def __init__(self, path, spec, fixed=None):
        self.path = path
        when = _dos_now()
        blobs = []     # (name bytes, header bytes of the .npy, payload bytes or None, shape, dtype, fill)
        for name, a in (fixed or {}).items():
            buf = io.BytesIO()
            np.lib.format.write_array(buf, np.asanyarray(a), allow_pickle=False)
            blobs.append(((name + ".npy").encode(), buf.getvalue(), None, None, None))
        for name, (shape, dtype, fill) in spec.items():
            blobs.append(((name + ".npy").encode(), _npy_header(shape, dtype), tuple(shape), np.dtype(dtype), fill))
        pos, layout = 0, []
        for nm, head, shape, dtype, fill in blobs:
            size = len(head) + (int(np.prod(shape, dtype=np.int64)) * dtype.itemsize if shape is not None else 0)
            pad = (-(pos + 30 + len(nm) + 20 + 4)) % _ALIGN            # the member's data start on a 64-byte boundary
            lh = _local_header(nm, size, zlib.crc32(head) & 0xFFFFFFFF if shape is None else 0, pad, when)
            layout.append((nm, pos, pos + len(lh), size))
            pos += len(lh) + size
        cd_off = pos
        self._f = open(path, "x+b")
        try:
            cd = b""
            for (nm, hoff, data, size), (_, head, shape, dtype, fill) in zip(layout, blobs):
                pad = (-(hoff + 30 + len(nm) + 20 + 4)) % _ALIGN
                crc = zlib.crc32(head) & 0xFFFFFFFF if shape is None else 0
                self._f.seek(hoff)
                self._f.write(_local_header(nm, size, crc, pad, when))
                self._f.write(head)
                cd += _central_record(nm, size, crc, hoff, when)
            n = len(layout)
            end = struct.pack("<4sQHHIIQQQQ", b"PK\x06\x06", 44, 45, 45, 0, 0, n, n, len(cd), cd_off)
            end += struct.pack("<4sIQI", b"PK\x06\x07", 0, cd_off + len(cd), 1)
            end += struct.pack("<4sHHHHIIH", b"PK\x05\x06", 0, 0, min(n, 0xFFFF), min(n, 0xFFFF), 0xFFFFFFFF, 0xFFFFFFFF, 0)
            self._f.seek(cd_off)
            self._f.write(cd + end)
            self._f.flush()
            self._mm = mmap.mmap(self._f.fileno(), 0)
        except Exception:
            self._f.close()
            raise
        self._members = _members(self._mm)
        self.arrays = {}
        for (nm, hoff, data, size), (_, head, shape, dtype, fill) in zip(layout, blobs):
            if shape is None:
                continue
            a = np.ndarray(shape, dtype=dtype, buffer=self._mm, offset=data + len(head))
            if fill != 0:
                a[...] = fill
            self.arrays[nm[:-4].decode()] = a
        self._open = True

    def flush(self):
        """Rows written through ``arrays`` are in the file as far as any reader is concerned (a shared mapping
        writes to the page cache, which is where HDF5's flush leaves its data too): nothing to do.  A synchronous
        ``msync`` here costs as much as the fit of the batch itself and guards only against a crash of the host."""

    def close(self):
        """Patch the members' checksums; the arrays stay valid (read-only views of the file)."""
        if not self._open:
            return
        view = memoryview(self._mm)

        def crc_of(name):
            hoff, data, size, cdrec = self._members[name + ".npy"]
            return name, zlib.crc32(view[data:data + size]) & 0xFFFFFFFF      # releases the GIL

        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=4) as ex:
            crcs = list(ex.map(crc_of, list(self.arrays)))
        for name, crc in crcs:
            hoff, data, size, cdrec = self._members[name + ".npy"]
            struct.pack_into("<I", self._mm, hoff + 14, crc)
            struct.pack_into("<I", self._mm, cdrec + 16, crc)
            self.arrays[name].setflags(write=False)
        view.release()
        self._f.close()
        self._open = False


def read_partial(path):
    """The members of an archive written by :class:`NpzInPlace`, WITHOUT the checksum test ``numpy.load`` applies:
    for the file an interrupted ``fit`` leaves behind."""
    out = {}
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            for name, (_, data, size, _) in _members(mm).items():
                out[name[:-4] if name.endswith(".npy") else name] = np.lib.format.read_array(
                    io.BytesIO(mm[data:data + size]), allow_pickle=False)
        finally:
            mm.close()
    return out
