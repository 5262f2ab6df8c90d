"""ctypes binding of oracle/_ref/libasvgf_ref.so: the reference's own compute shaders (unmodified GLSL text compiled as
C++ by oracle/ref_glsl/build_ref.py) dispatched the way Source/Denoiser.cpp:87-254 dispatches them.

TEST INFRASTRUCTURE ONLY, and only where /root/reference exists (this container): it pins the hand-written oracle
(oracle/asvgf_oracle.cpp) and generates the golden fixtures under tests/golden/.  Nothing under raytracedgl1_b200/ may
import it.
"""
import ctypes
import math
import os
import subprocess
import sys

import numpy as np

from raytracedgl1_b200 import layout

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "_ref", "libasvgf_ref.so")
_lib = None

# glslshim::Format
_SHIM_FORMAT = {
    layout.R8_UINT: 1, layout.R32_UINT: 2, layout.R8G8_UNORM: 4, layout.R8G8B8A8_UNORM: 5, layout.R16_SFLOAT: 7,
    layout.R16G16_SFLOAT: 8, layout.R16G16B16A16_SFLOAT: 9, layout.R32G32B32A32_SFLOAT: 11, layout.B10G11R11_UFLOAT: 12,
}
FRAMEBUFFER_IMAGE_COUNT = 76


class _ImageDesc(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("format", ctypes.c_int32)]


class _DispatchContext(ctypes.Structure):
    _fields_ = [("images", ctypes.POINTER(_ImageDesc)), ("buffers", (ctypes.c_void_p * 4) * 8),
                ("specConstant", ctypes.c_uint32 * 4), ("pushConstant", ctypes.c_uint32 * 32)]


def available():
    return os.path.exists(_LIB_PATH)


def build(reference="/root/reference"):
    """Runs oracle/ref_glsl/build_ref.py (needs the reference tree; a no-op without it)."""
    if os.path.isdir(os.path.join(reference, "Source", "Shaders")):
        subprocess.run([sys.executable, os.path.join(_DIR, "ref_glsl", "build_ref.py"), "--reference", reference], check=True)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("oracle/_ref/libasvgf_ref.so is not built (python oracle/ref_glsl/build_ref.py; needs /root/reference)")
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _work_group_count(size, group):
    """Utils::GetWorkGroupCount(float, uint32_t), Source/Utils.cpp:462-476 (one group of over-coverage included)."""
    return 1 + (int(math.ceil(size)) + (group - 1)) // group


class RefDenoiser:
    """The image table (numpy arrays, one per physical image) plus Denoiser::Denoise as a list of shader dispatches.
    Same surface as oracle.Oracle: image(name, frame_index), run_pass(p, frame_index, uniform), denoise(...)."""

    SHADER_OF_PASS = {1: "CmASVGFGradientAtrous", 2: "CmASVGFGradientAtrous", 3: "CmASVGFGradientAtrous", 4: "CmASVGFGradientAtrous",
                      5: "CmSVGFTemporalAccumulation", 6: "CmAntiFirefly", 7: "CmSVGFEstimateVariance", 8: "CmSVGFAtrous_Iter0",
                      9: "CmSVGFAtrous", 10: "CmSVGFAtrous", 11: "CmSVGFAtrous"}

    def __init__(self, width, height):
        self.width, self.height = width, height
        self._arrays = {}
        for name, info in layout.IMAGES.items():
            dtype, shape = layout.image_numpy_shape(name, width, height)
            self._arrays[name] = np.zeros(shape, dtype=dtype)
        self.tonemapping = layout.Tonemapping()
        self._local = {}

    def close(self):
        self._arrays = {}

    def image(self, name, frame_index):
        """Writable numpy array of the PHYSICAL image the logical name maps to for frame_index (Framebuffers.cpp:35-50)."""
        from raytracedgl1_b200 import bands
        return self._arrays[bands.physical_name(name, frame_index)]

    def _descriptors(self, frame_index):
        descs = (_ImageDesc * FRAMEBUFFER_IMAGE_COUNT)()
        for name, info in layout.IMAGES.items():
            idx = info[0]
            if idx >= FRAMEBUFFER_IMAGE_COUNT:
                continue          # images this repo adds (FwdLum, DebugMask, ScatterCurrent) are not in the reference's table
            a = self.image(name, frame_index)
            descs[idx] = _ImageDesc(a.ctypes.data, a.shape[1], a.shape[0], _SHIM_FORMAT[info[1]])
        return descs

    def dispatch(self, shader, frame_index, uniform, size_x, size_y, spec_constant=0, push=None, lpm=None):
        L = lib()
        if shader not in self._local:
            xy = (ctypes.c_uint32 * 2)()
            getattr(L, "asvgfRefLocalSize_" + shader)(xy)
            self._local[shader] = (xy[0], xy[1])
        lx, ly = self._local[shader]
        ctx = _DispatchContext()
        descs = self._descriptors(frame_index)
        ctx.images = ctypes.cast(descs, ctypes.POINTER(_ImageDesc))
        u = ctypes.cast(ctypes.byref(uniform), ctypes.c_void_p)
        tm = ctypes.cast(ctypes.byref(self.tonemapping), ctypes.c_void_p)
        # descriptor sets as the shaders number them: global uniform = set 1 binding 0; tonemapping = set 2 binding 0
        ctx.buffers[1][0] = u
        ctx.buffers[2][0] = tm
        if lpm is not None:     # DESC_SET_LPM_PARAMS = 3 (CmPrepareFinal.comp:26), BINDING_LPM_PARAMS = 0
            ctx.buffers[3][0] = lpm.ctypes.data
        ctx.specConstant[0] = spec_constant
        if push is not None:
            raw = np.frombuffer(np.asarray(push, dtype=np.float32).tobytes(), dtype=np.uint32)
            for i, v in enumerate(raw):
                ctx.pushConstant[i] = int(v)
        fn = getattr(L, "asvgfRef_" + shader)
        fn.argtypes = [ctypes.POINTER(_DispatchContext), ctypes.c_uint32, ctypes.c_uint32]
        fn(ctypes.byref(ctx), _work_group_count(size_x, lx), _work_group_count(size_y, ly))

    def run_pass(self, p, frame_index, uniform):
        """One dispatch of Denoiser::Denoise (pass numbering of include/asvgf.h; pass 0, the ray-traced gradient-sample
        stage RtGradients.rgen, has no compute-shader form and is not available here)."""
        shader = self.SHADER_OF_PASS[p]
        W, H = uniform.renderWidth, uniform.renderHeight
        if 1 <= p <= 4:   # Denoiser.cpp:87-120
            self.dispatch(shader, frame_index, uniform, W / 3.0, H / 3.0, spec_constant=p - 1)
        elif p >= 9:      # Denoiser.cpp:198-254: iteration i >= 1 is CmSVGFAtrous.comp specialised with i
            self.dispatch(shader, frame_index, uniform, W, H, spec_constant=p - 8)
        else:
            self.dispatch(shader, frame_index, uniform, W, H)

    def denoise(self, frame_index, uniform):
        for p in range(1, 12):
            self.run_pass(p, frame_index, uniform)

    def resolve_checkerboard(self, frame_index, uniform):      # ImageComposition.cpp:157-222
        self.dispatch("CmCheckerboard", frame_index, uniform, uniform.renderWidth, uniform.renderHeight)

    def luminance_histogram(self, frame_index, uniform):       # Tonemapping.cpp:59-115
        self.dispatch("CmLuminanceHistogram", frame_index, uniform, uniform.renderWidth, uniform.renderHeight)

    def luminance_average(self, uniform):                      # Tonemapping.cpp:117-175: one work group
        L = lib()
        ctx = _DispatchContext()
        descs = self._descriptors(0)
        ctx.images = ctypes.cast(descs, ctypes.POINTER(_ImageDesc))
        ctx.buffers[1][0] = ctypes.cast(ctypes.byref(uniform), ctypes.c_void_p)
        ctx.buffers[2][0] = ctypes.cast(ctypes.byref(self.tonemapping), ctypes.c_void_p)
        fn = L.asvgfRef_CmLuminanceAvg
        fn.argtypes = [ctypes.POINTER(_DispatchContext), ctypes.c_uint32, ctypes.c_uint32]
        fn(ctypes.byref(ctx), 1, 1)

    def scatter_accumulate(self, frame_index, uniform, max_history_length):   # Volumetric.cpp:71-195
        self.dispatch("CmScatterAccum", frame_index, uniform, uniform.renderWidth, uniform.renderHeight,
                      push=[max_history_length])

    def prepare_final(self, frame_index, uniform, lpm_control):  # ImageComposition.cpp:110-155 (ApplyTonemapping)
        lpm = np.ascontiguousarray(lpm_control, dtype=np.uint32)
        assert lpm.size == 96
        self.dispatch("CmPrepareFinal", frame_index, uniform, uniform.renderWidth, uniform.renderHeight, lpm=lpm)


def lpm_setup(saturation=(0.0, 0.0, 0.0), crosstalk=(1.0, 1.0, 1.0)):
    """The 24 x uvec4 LPM control block as ImageComposition::SetupLpmParams computes it (ImageComposition.cpp:365-420; AMD's LpmSetup
    from the reference's vendored headers, compiled in their CPU mode by oracle/ref_glsl/lpm_setup_ref.cpp).  Defaults =
    RgDrawFrameTonemappingParams' defaults (Source/DrawFrameInfo.h:104-105)."""
    out = np.zeros(96, dtype=np.uint32)
    sat = (ctypes.c_float * 3)(*saturation)
    cro = (ctypes.c_float * 3)(*crosstalk)
    fn = lib().asvgfRefLpmSetup
    fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    fn(out.ctypes.data, sat, cro)
    return out
