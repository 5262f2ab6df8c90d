"""ctypes binding of the CPU oracle (oracle/libasvgf_oracle.so).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never from raytracedgl1_b200 (the product path has no CPU fallback).
"""
import ctypes
import os
import subprocess

import numpy as np

from raytracedgl1_b200 import layout

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "libasvgf_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_DIR, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("asvgf_oracle.cpp", "asvgf_oracle.h", "Makefile")):
        subprocess.run(["make", "-C", _DIR, "-B" if force else "-s", "libasvgf_oracle.so"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, u32, i32, f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_float
        L.asvgfOracleCreate.restype = vp
        L.asvgfOracleCreate.argtypes = [u32, u32, u32]
        L.asvgfOracleDestroy.argtypes = [vp]
        L.asvgfOracleSetThreads.argtypes = [i32]
        L.asvgfOracleGetThreads.restype = i32
        L.asvgfOracleGetImage.restype = i32
        L.asvgfOracleGetImage.argtypes = [vp, u32, u32, ctypes.POINTER(vp), ctypes.POINTER(u32), ctypes.POINTER(u32),
                                          ctypes.POINTER(u32)]
        L.asvgfOracleResetHistory.argtypes = [vp]
        L.asvgfOracleRunPass.argtypes = [vp, u32, u32, ctypes.POINTER(layout.GlobalUniform)]
        L.asvgfOracleGradientSamples.argtypes = [vp, u32, ctypes.POINTER(layout.GlobalUniform)]
        L.asvgfOracleDenoise.argtypes = [vp, u32, ctypes.POINTER(layout.GlobalUniform)]
        puni = ctypes.POINTER(layout.GlobalUniform)
        L.asvgfOracleResolveCheckerboard.argtypes = [vp, u32, puni]
        L.asvgfOraclePrepareFinal.argtypes = [vp, u32, puni, ctypes.c_void_p]
        L.asvgfOracleLuminanceHistogram.argtypes = [vp, u32, puni]
        L.asvgfOracleLuminanceAverage.argtypes = [vp, puni]
        L.asvgfOracleCalculateExposure.argtypes = [vp, u32, puni]
        L.asvgfOracleScatterAccumulate.argtypes = [vp, u32, puni, f32]
        L.asvgfOracleGetTonemapping.argtypes = [vp]
        L.asvgfOracleGetTonemapping.restype = ctypes.POINTER(layout.Tonemapping)
        L.asvgfOracleEncodeE5B9G9R9.restype = u32
        L.asvgfOracleEncodeE5B9G9R9.argtypes = [f32, f32, f32]
        L.asvgfOracleDecodeE5B9G9R9.argtypes = [u32, ctypes.POINTER(f32)]
        L.asvgfOracleEncodeNormal.restype = u32
        L.asvgfOracleEncodeNormal.argtypes = [f32, f32, f32]
        L.asvgfOracleDecodeNormal.argtypes = [u32, ctypes.POINTER(f32)]
        L.asvgfOracleFloatToHalf.restype = ctypes.c_uint16
        L.asvgfOracleFloatToHalf.argtypes = [f32]
        L.asvgfOracleHalfToFloat.restype = f32
        L.asvgfOracleHalfToFloat.argtypes = [ctypes.c_uint16]
        L.asvgfOraclePackUfloat111110.restype = u32
        L.asvgfOraclePackUfloat111110.argtypes = [f32, f32, f32]
        L.asvgfOracleUnpackUfloat111110.argtypes = [u32, ctypes.POINTER(f32)]
        L.asvgfOracleFloatToUnorm8.restype = ctypes.c_uint8
        L.asvgfOracleFloatToUnorm8.argtypes = [f32]
        L.asvgfOracleCheckerboardPix.argtypes = [i32, i32, i32, ctypes.POINTER(i32)]
        L.asvgfOracleRegularPix.argtypes = [i32, i32, i32, ctypes.POINTER(i32)]
        L.asvgfOraclePrevScreenPos.argtypes = [f32, f32, i32, i32, f32, f32, ctypes.POINTER(f32)]
        _lib = L
    return _lib


class Oracle:
    """Mirror of Framebuffers + Denoiser on the CPU: images are numpy views of the oracle's own storage."""

    def __init__(self, width, height, atrous_iterations=4):
        self.width, self.height, self.iterations = width, height, atrous_iterations
        self._h = lib().asvgfOracleCreate(width, height, atrous_iterations)
        if not self._h:
            raise ValueError("asvgfOracleCreate rejected %dx%d N=%d" % (width, height, atrous_iterations))

    def close(self):
        if self._h:
            lib().asvgfOracleDestroy(self._h)
            self._h = None

    __del__ = close

    def image(self, name, frame_index):
        """Writable numpy view (cur/prev swap applied for frame_index)."""
        idx = layout.INDEX[name] if isinstance(name, str) else name
        p, w, h, bpp = ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        if lib().asvgfOracleGetImage(self._h, idx, frame_index, ctypes.byref(p), ctypes.byref(w), ctypes.byref(h),
                                     ctypes.byref(bpp)):
            raise KeyError(name)
        dtype, shape = layout.image_numpy_shape(idx, self.width, self.height)
        assert shape[0] == h.value and shape[1] == w.value
        n = int(np.prod(shape))
        buf = (ctypes.c_uint8 * (n * dtype.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def reset_history(self):
        lib().asvgfOracleResetHistory(self._h)

    def run_pass(self, p, frame_index, uniform):
        lib().asvgfOracleRunPass(self._h, p, frame_index, ctypes.byref(uniform))

    def gradient_samples(self, frame_index, uniform):
        lib().asvgfOracleGradientSamples(self._h, frame_index, ctypes.byref(uniform))

    def denoise(self, frame_index, uniform):
        lib().asvgfOracleDenoise(self._h, frame_index, ctypes.byref(uniform))

    def scatter_accumulate(self, frame_index, uniform, max_history_length):
        lib().asvgfOracleScatterAccumulate(self._h, frame_index, ctypes.byref(uniform), max_history_length)

    def resolve_checkerboard(self, frame_index, uniform):
        lib().asvgfOracleResolveCheckerboard(self._h, frame_index, ctypes.byref(uniform))

    def prepare_final(self, frame_index, uniform, lpm_control):
        import numpy as np
        lpm = np.ascontiguousarray(lpm_control, dtype=np.uint32)
        assert lpm.size == 96
        lib().asvgfOraclePrepareFinal(self._h, frame_index, ctypes.byref(uniform), lpm.ctypes.data)

    def luminance_histogram(self, frame_index, uniform):
        lib().asvgfOracleLuminanceHistogram(self._h, frame_index, ctypes.byref(uniform))

    def luminance_average(self, uniform):
        lib().asvgfOracleLuminanceAverage(self._h, ctypes.byref(uniform))

    def calculate_exposure(self, frame_index, uniform):
        lib().asvgfOracleCalculateExposure(self._h, frame_index, ctypes.byref(uniform))

    @property
    def tonemapping(self):
        """The oracle's own AsvgfTonemapping (a live ctypes view: histogram[256], avgLuminance)."""
        return lib().asvgfOracleGetTonemapping(self._h).contents


def set_threads(n):
    lib().asvgfOracleSetThreads(n)


def get_threads():
    return lib().asvgfOracleGetThreads()
