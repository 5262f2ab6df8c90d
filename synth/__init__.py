"""ctypes binding of the synthetic G-buffer generator (synth/libasvgf_synth.so): the stand-in for RTGL1's path tracer
as producer of the denoiser's inputs.  Used by tests and bench.py only."""
import ctypes
import os
import subprocess

import numpy as np

from raytracedgl1_b200 import layout

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "libasvgf_synth.so")
_lib = None
DEFAULT_SEED = 0x5EED0001


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_DIR, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("asvgf_synth.cpp", "Makefile")):
        subprocess.run(["make", "-C", _DIR, "-B" if force else "-s", "libasvgf_synth.so"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.asvgfSynthFillUniform.argtypes = [ctypes.POINTER(layout.GlobalUniform), ctypes.c_uint32, ctypes.c_uint32,
                                            ctypes.c_uint32]
        L.asvgfSynthGenerate.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                         ctypes.POINTER(ctypes.c_void_p)]
        L.asvgfSynthGenerateRows.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32, ctypes.c_uint32]
        _lib = L
    return _lib


def uniform(width, height, frame):
    u = layout.GlobalUniform()
    lib().asvgfSynthFillUniform(ctypes.byref(u), width, height, frame)
    return u


def generate(width, height, frame, seed=DEFAULT_SEED, out=None):
    """Returns {name: numpy array} of every producer image of `frame` (tightly packed, storage formats).
    `out` may supply preallocated arrays (e.g. views of pinned memory) keyed by name."""
    arrays = {}
    ptrs = (ctypes.c_void_p * layout.IMAGE_COUNT)()
    for name in layout.PRODUCER_IMAGES:
        dtype, shape = layout.image_numpy_shape(name, width, height)
        a = out[name] if out is not None and name in out else np.zeros(shape, dtype=dtype)
        assert a.dtype == dtype and a.shape == shape and a.flags["C_CONTIGUOUS"]
        arrays[name] = a
        ptrs[layout.INDEX[name]] = a.ctypes.data
    lib().asvgfSynthGenerate(width, height, frame, seed, ptrs)
    return arrays


def generate_rows(width, height, frame, y0, y1, seed=DEFAULT_SEED):
    """Rows [y0, y1) of every producer image of `frame` (what one rank of the band mode stores: band + halo)."""
    arrays = {}
    ptrs = (ctypes.c_void_p * layout.IMAGE_COUNT)()
    for name in layout.PRODUCER_IMAGES:
        dtype, shape = layout.image_numpy_shape(name, width, height)
        a = np.zeros((y1 - y0,) + tuple(shape[1:]), dtype=dtype)
        arrays[name] = a
        ptrs[layout.INDEX[name]] = a.ctypes.data
    lib().asvgfSynthGenerateRows(width, height, frame, seed, ptrs, y0, y1)
    return arrays
