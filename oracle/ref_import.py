"""TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Imports the *unmodified* reference package from /root/reference so that golden vectors can be
generated and the C restatement (oracle/loglike_ref.c) can be pinned against it.  The reference
needs h5py/astropy/healpy/pooch at import time, none of which exist in this image; they are
replaced by inert placeholder modules.  /root/reference does not exist on
the GPU box, so nothing that runs there (``-m gpu`` tests, smoke(), bench.py) may call this.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BRUTUS_REFERENCE_ROOT", "/root/reference")


class _Placeholder(types.ModuleType):
    """A module whose every attribute is another callable placeholder."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Placeholder(self.__name__ + "." + name)
        object.__setattr__(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return _Placeholder(self.__name__ + "()")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "brutus"))


def import_reference():
    """Return the reference's ``brutus.fitting`` module (numba-jitted, CPU)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # keep numba caches / bytecode out of the read-only reference tree
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/brutus_b200_numba_cache")
    sys.dont_write_bytecode = True
    for name in ("h5py", "astropy", "astropy.units", "astropy.coordinates", "healpy", "pooch"):
        if name not in sys.modules:
            sys.modules[name] = _Placeholder(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from brutus import fitting  # noqa: E402  (the reference, not this repo)
    return fitting
