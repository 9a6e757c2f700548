"""Loader of the native library.  There is NO fallback: if libmmb200.so is missing the import of
any compute entry point raises — the product never computes on the CPU."""
import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libmmb200.so")
_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise NativeLibraryMissing(
                "libmmb200.so not found at %s — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % SO_PATH
            )
        _lib = _abi.bind(ctypes.CDLL(SO_PATH))
    return _lib


def last_error():
    return load().mm_last_error().decode()


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, last_error()))
