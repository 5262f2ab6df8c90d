"""Python mirror of include/asvgf_layout.h: image table, storage formats, GlobalUniform.

Facts about the reference interface (indices: Source/Generated/ShaderCommonGLSL.h:408-483, formats:
Source/Generated/GenerateShaderCommon.py:802-887, uniform: Source/Generated/ShaderCommonC.h:209-304).
"""
import ctypes

import numpy as np

# format enum (AsvgfFormat)
R8_UINT, R8G8_UNORM, R8G8B8A8_UNORM, R16_SFLOAT, R16G16_SFLOAT, R16G16B16A16_SFLOAT, R32_UINT, \
    R32G32B32A32_SFLOAT, B10G11R11_UFLOAT = range(1, 10)

_FORMAT_NUMPY = {
    R8_UINT: (np.uint8, 1), R8G8_UNORM: (np.uint8, 2), R8G8B8A8_UNORM: (np.uint8, 4),
    R16_SFLOAT: (np.float16, 1), R16G16_SFLOAT: (np.float16, 2), R16G16B16A16_SFLOAT: (np.float16, 4),
    R32_UINT: (np.uint32, 1), R32G32B32A32_SFLOAT: (np.float32, 4), B10G11R11_UFLOAT: (np.uint32, 1),
}

# name -> (index, format, third_res, regular_layout, kind) ; kind: 'in' producer-written, 'state', 'scratch', 'out'
IMAGES = {
    "Albedo": (0, B10G11R11_UFLOAT, False, "in"),
    "IsSky": (1, R8_UINT, False, "in"),
    "Normal": (2, R32_UINT, False, "in"),
    "Normal_Prev": (3, R32_UINT, False, "in"),
    "MetallicRoughness": (4, R8G8_UNORM, False, "in"),
    "MetallicRoughness_Prev": (5, R8G8_UNORM, False, "in"),
    "DepthWorld": (6, R16_SFLOAT, False, "in"),
    "DepthWorld_Prev": (7, R16_SFLOAT, False, "in"),
    "DepthGrad": (8, R16_SFLOAT, False, "in"),
    "Motion": (10, R16G16B16A16_SFLOAT, False, "in"),
    "UnfilteredDirect": (11, R32_UINT, False, "in"),
    "UnfilteredSpecular": (12, R32_UINT, False, "in"),
    "UnfilteredIndir": (13, R32_UINT, False, "in"),
    "SurfacePosition": (14, R32G32B32A32_SFLOAT, False, "in"),
    "SurfacePosition_Prev": (15, R32G32B32A32_SFLOAT, False, "in"),
    "ViewDirection": (18, R16G16B16A16_SFLOAT, False, "in"),
    "Throughput": (20, R16G16B16A16_SFLOAT, False, "in"),
    "PreFinal": (21, B10G11R11_UFLOAT, False, "out"),
    "Final": (22, B10G11R11_UFLOAT, False, "out"),  # PreFinal after the checkerboard resolve (CmCheckerboard.comp)
    "AccumHistoryLength": (27, R16G16B16A16_SFLOAT, False, "state"),
    "AccumHistoryLength_Prev": (28, R16G16B16A16_SFLOAT, False, "state"),
    "DiffTemporary": (29, R32_UINT, False, "scratch"),
    "DiffAccumColor": (30, R32_UINT, False, "state"),
    "DiffAccumColor_Prev": (31, R32_UINT, False, "state"),
    "DiffAccumMoments": (32, R16G16_SFLOAT, False, "state"),
    "DiffAccumMoments_Prev": (33, R16G16_SFLOAT, False, "state"),
    "DiffColorHistory": (34, R16G16B16A16_SFLOAT, False, "state"),
    "DiffPingColorAndVariance": (35, R16G16B16A16_SFLOAT, False, "scratch"),
    "DiffPongColorAndVariance": (36, R16G16B16A16_SFLOAT, False, "scratch"),
    "SpecAccumColor": (37, R32_UINT, False, "state"),
    "SpecAccumColor_Prev": (38, R32_UINT, False, "state"),
    "SpecPingColor": (39, R32_UINT, False, "scratch"),
    "SpecPongColor": (40, R32_UINT, False, "scratch"),
    "IndirAccum": (41, R32_UINT, False, "state"),
    "IndirAccum_Prev": (42, R32_UINT, False, "state"),
    "IndirPing": (43, R32_UINT, False, "scratch"),
    "IndirPong": (44, R32_UINT, False, "scratch"),
    "AtrousFilteredVariance": (45, R16_SFLOAT, False, "state"),
    "HistogramInput": (46, R16_SFLOAT, False, "out"),
    "Scattering": (48, R16G16B16A16_SFLOAT, False, "state"),        # regular layout (CmScatterAccum.comp)
    "Scattering_Prev": (49, R16G16B16A16_SFLOAT, False, "state"),
    "ScatteringHistory": (50, R16_SFLOAT, False, "state"),
    "ScatteringHistory_Prev": (51, R16_SFLOAT, False, "state"),
    # the two other image pairs of the checkerboard resolve (CmCheckerboard.comp:62-77), regular layout
    "AcidFogRT": (52, B10G11R11_UFLOAT, False, "in"),
    "AcidFog": (53, B10G11R11_UFLOAT, False, "out"),
    "ScreenEmisRT": (54, B10G11R11_UFLOAT, False, "in"),
    "ScreenEmission": (55, B10G11R11_UFLOAT, False, "out"),
    "BloomInput": (56, B10G11R11_UFLOAT, False, "out"),            # asvgfPrepareFinal (CmPrepareFinal.comp:171)
    "GradientInputs": (70, R16G16_SFLOAT, False, "in"),
    "GradientInputs_Prev": (71, R16G16_SFLOAT, False, "in"),
    "DISPingGradient": (72, R8G8B8A8_UNORM, True, "scratch"),
    "DISPongGradient": (73, R8G8B8A8_UNORM, True, "scratch"),
    "DISGradientHistory": (74, R8G8B8A8_UNORM, True, "state"),
    "GradientPrevPix": (75, R8_UINT, True, "state"),
    "FwdLum": (76, R16G16_SFLOAT, False, "in"),
    "DebugMask": (77, R8_UINT, False, "out"),
    "ScatterCurrent": (78, R16G16B16A16_SFLOAT, False, "in"),       # extension: the VOLUMETRIC-mode scattering sample
}
INDEX = {name: v[0] for name, v in IMAGES.items()}
NAME_OF = {v[0]: name for name, v in IMAGES.items()}
IMAGE_COUNT = 79

# images a G-buffer producer writes every frame (what synth generates / what the host uploads)
PRODUCER_IMAGES = ["Albedo", "IsSky", "Normal", "MetallicRoughness", "DepthWorld", "DepthGrad", "Motion",
                   "UnfilteredDirect", "UnfilteredSpecular", "UnfilteredIndir", "SurfacePosition", "ViewDirection",
                   "Throughput", "GradientInputs", "FwdLum"]
# E5B9G9R9-packed R32_UINT images (the others are encoded normals)
E5_IMAGES = {"UnfilteredDirect", "UnfilteredSpecular", "UnfilteredIndir", "DiffTemporary", "DiffAccumColor",
             "DiffAccumColor_Prev", "SpecAccumColor", "SpecAccumColor_Prev", "SpecPingColor", "SpecPongColor",
             "IndirAccum", "IndirAccum_Prev", "IndirPing", "IndirPong"}

PASS_GRADIENT_SAMPLES = 0
PASS_GRADIENT_ATROUS_0 = 1
PASS_TEMPORAL = 5
PASS_ANTIFIREFLY = 6
PASS_VARIANCE = 7
PASS_ATROUS_0 = 8
PASS_GRADIENT_ATROUS_ALL = 100  # the four gradient a-trous dispatches in one call
PASS_GUIDE = 101                # decoded geometry plane of the frame (not a reference dispatch)


def image_extent(name_or_index, width, height):
    """(w, h) of an image for a render size (Source/Framebuffers.cpp:639-700)."""
    name = NAME_OF[name_or_index] if isinstance(name_or_index, int) else name_or_index
    if IMAGES[name][2]:
        return max(1, (width + 1) // 3), max(1, (height + 1) // 3)
    return width, height


def image_numpy_shape(name_or_index, width, height):
    name = NAME_OF[name_or_index] if isinstance(name_or_index, int) else name_or_index
    dtype, ch = _FORMAT_NUMPY[IMAGES[name][1]]
    w, h = image_extent(name, width, height)
    return (np.dtype(dtype), (h, w) if ch == 1 else (h, w, ch))


def bytes_per_pixel(name_or_index):
    name = NAME_OF[name_or_index] if isinstance(name_or_index, int) else name_or_index
    dtype, ch = _FORMAT_NUMPY[IMAGES[name][1]]
    return np.dtype(dtype).itemsize * ch


class GlobalUniform(ctypes.Structure):
    """Byte-compatible with ShGlobalUniform (2096 bytes); only the members the path reads are named."""
    _fields_ = [
        ("view", ctypes.c_float * 16), ("invView", ctypes.c_float * 16), ("viewPrev", ctypes.c_float * 16),
        ("projection", ctypes.c_float * 16), ("invProjection", ctypes.c_float * 16),
        ("projectionPrev", ctypes.c_float * 16),
        ("_reserved0", ctypes.c_uint32 * 65),
        ("lightmapScreenCoverage", ctypes.c_float),
        ("_reserved1", ctypes.c_uint32),
        ("renderWidth", ctypes.c_float), ("renderHeight", ctypes.c_float),
        ("frameId", ctypes.c_uint32), ("timeDelta", ctypes.c_float),
        ("minLogLuminance", ctypes.c_float), ("maxLogLuminance", ctypes.c_float),
        ("luminanceWhitePoint", ctypes.c_float), ("stopEyeAdaptation", ctypes.c_uint32),
        ("_reserved2", ctypes.c_uint32 * 9),
        ("cameraPosition", ctypes.c_float * 4), ("cameraPositionPrev", ctypes.c_float * 4),
        ("debugShowFlags", ctypes.c_uint32),
        ("_reserved3", ctypes.c_uint32 * 4), ("emissionMaxScreenColor", ctypes.c_float),
        ("_reserved3a", ctypes.c_uint32 * 7), ("bloomThreshold", ctypes.c_float), ("bloomEmissionMultiplier", ctypes.c_float),
        ("_reserved3c", ctypes.c_uint32 * 5),
        ("volumeEnableType", ctypes.c_uint32), ("volumeScattering", ctypes.c_float),
        ("_reserved3b", ctypes.c_uint32 * 23),
        ("jitterX", ctypes.c_float), ("jitterY", ctypes.c_float),
        ("_reserved4", ctypes.c_uint32 * 6),
        ("gradientMultDiffuse", ctypes.c_float), ("gradientMultIndirect", ctypes.c_float),
        ("gradientMultSpecular", ctypes.c_float),
        ("_reserved5", ctypes.c_uint32 * 3),
        ("antiFireflyEnabled", ctypes.c_uint32),
        ("volumeAmbient", ctypes.c_float * 4),
        ("_reserved6", ctypes.c_uint32 * 272),
    ]


assert ctypes.sizeof(GlobalUniform) == 2096
assert GlobalUniform.minLogLuminance.offset == 668 and GlobalUniform.stopEyeAdaptation.offset == 680
assert GlobalUniform.volumeEnableType.offset == 832 and GlobalUniform.volumeAmbient.offset == 992


class Tonemapping(ctypes.Structure):
    """Byte-compatible with ShTonemapping (Source/Generated/ShaderCommonC.h:344-348)."""
    _fields_ = [("histogram", ctypes.c_uint32 * 256), ("avgLuminance", ctypes.c_float)]


assert ctypes.sizeof(Tonemapping) == 1028
assert GlobalUniform.renderWidth.offset == 652 and GlobalUniform.antiFireflyEnabled.offset == 988


# ---- numpy codecs for COMPARING images (decode only; tests use them on both sides alike) ----
def decode_e5(u32):
    u32 = np.asarray(u32, dtype=np.uint32)
    scale = np.ldexp(np.float32(1.0), (u32 >> 27).astype(np.int32) - 24).astype(np.float32)
    return np.stack([(u32 & 0x1FF).astype(np.float32) * scale, ((u32 >> 9) & 0x1FF).astype(np.float32) * scale,
                     ((u32 >> 18) & 0x1FF).astype(np.float32) * scale], axis=-1)


def _decode_ufloat(code, mbits):
    code = code.astype(np.uint32)
    e = (code >> mbits).astype(np.int32)
    m = (code & ((1 << mbits) - 1)).astype(np.float32)
    den = np.float32(1 << mbits)
    normal = np.ldexp(np.float32(1.0) + m / den, e - 15)
    sub = np.ldexp(m / den, -14)
    return np.where(e == 0, sub, normal).astype(np.float32)


def decode_b10g11r11(u32):
    u32 = np.asarray(u32, dtype=np.uint32)
    return np.stack([_decode_ufloat(u32 & 0x7FF, 6), _decode_ufloat((u32 >> 11) & 0x7FF, 6),
                     _decode_ufloat(u32 >> 22, 5)], axis=-1)


def luminance(rgb):
    return 0.2125 * rgb[..., 0] + 0.7154 * rgb[..., 1] + 0.0721 * rgb[..., 2]
