"""B200-native A-SVGF denoiser (the pass chain behind RTGL1's Source/Denoiser.cpp) -- host-side mirror.

The compute lives in libasvgf.so (hand-written sm_100a CUDA behind the C ABI of include/asvgf.h); this package is
the thin ctypes mirror of the reference's Denoiser/Framebuffers/GlobalUniform interface.
"""
from . import layout  # noqa: F401
from .denoiser import AsvgfError, Denoiser  # noqa: F401
from .layout import GlobalUniform  # noqa: F401
