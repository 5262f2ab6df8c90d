"""Host-side mirror of RTGL1's Denoiser + Framebuffers over the C ABI of libasvgf.so (include/asvgf.h).

Reference interface mirrored here (same call order, argument meaning and error behaviour):
  RTGL1::Denoiser::Denoise(cmd, frameIndex, uniform)            Source/Denoiser.h:45-47
  RTGL1::PathTracer::CalculateGradientsSamples(params)          Source/PathTracer.cpp:196-219
  RTGL1::Framebuffers::GetImage(fbImageIndex, frameIndex)       Source/Framebuffers.cpp:35-50 (cur/prev swap)
`cmd` becomes a CUDA stream handle (an int, e.g. torch.cuda.current_stream().cuda_stream); calls only enqueue.
There is no CPU fallback: constructing a Denoiser without libasvgf.so or without a CUDA device raises.
"""
import ctypes
import os

import numpy as np

from . import layout

_LIB_PATH = os.environ.get("ASVGF_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libasvgf.so")  # ASVGF_LIB: A/B builds
_lib = None
_TEST_ONLY_ALLOW_HOST_BUILD = False  # set by tests/conftest.py in the child pytest of tests/test_hostsim.py, by nothing else

SUCCESS, WRONG_ARGUMENT, CUDA_ERROR, OUT_OF_MEMORY, UNSUPPORTED, NOT_BOUND, HALO_EXCEEDED = range(7)
CREATE_EXPORT_DEBUG_MASK = 1 << 0
CREATE_EXTERNAL_IMAGES = 1 << 1
CREATE_LITERAL_KERNELS = 1 << 2
CREATE_EAGER_GRADIENT_ATROUS = 1 << 4
CREATE_BAND_APRON = 1 << 3

EXPORTED_SYMBOLS = [
    "asvgfCreate", "asvgfDestroy", "asvgfGetImageInfo", "asvgfGetImage", "asvgfBindExternalImage", "asvgfUploadImage",
    "asvgfDownloadImage", "asvgfResetHistory", "asvgfGradientSamples", "asvgfDenoise", "asvgfDenoisePass",
    "asvgfSynchronize", "asvgfGetLaunchCount", "asvgfPassName", "asvgfPassBytesPerPixel", "asvgfResultString",
    "asvgfLastErrorMessage", "asvgfPassKernelName", "asvgfBandApronRows", "asvgfResolveCheckerboard", "asvgfCalculateExposure", "asvgfLuminanceHistogram",
    "asvgfLuminanceAverage", "asvgfGetTonemappingBuffer", "asvgfUploadTonemapping", "asvgfDownloadTonemapping",
    "asvgfScatterAccumulate", "asvgfPrepareFinal", "asvgfGetStateSize", "asvgfSaveState", "asvgfLoadState",
    "asvgfBandSetCreate", "asvgfBandSetDestroy", "asvgfBandSetGetBandCount", "asvgfBandSetGetBand", "asvgfBandSetFrame",
    "asvgfBandSetSynchronize", "asvgfBandSetBeginTiming", "asvgfBandSetEndTiming",
    "asvgfImportExternalMemoryFd", "asvgfExternalMemoryGetBuffer", "asvgfDestroyExternalMemory", "asvgfImportExternalSemaphoreFd",
    "asvgfWaitExternalSemaphore", "asvgfSignalExternalSemaphore", "asvgfDestroyExternalSemaphore",
]


class CreateInfo(ctypes.Structure):
    _fields_ = [("renderWidth", ctypes.c_uint32), ("renderHeight", ctypes.c_uint32),
                ("atrousIterations", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("device", ctypes.c_int32),
                ("bandRowBegin", ctypes.c_uint32), ("bandRowEnd", ctypes.c_uint32), ("bandHaloRows", ctypes.c_uint32)]


class BandSetCreateInfo(ctypes.Structure):
    _fields_ = [("renderWidth", ctypes.c_uint32), ("renderHeight", ctypes.c_uint32), ("atrousIterations", ctypes.c_uint32),
                ("flags", ctypes.c_uint32), ("deviceCount", ctypes.c_uint32), ("pDevices", ctypes.POINTER(ctypes.c_int32)),
                ("pBandRowBoundaries", ctypes.POINTER(ctypes.c_uint32)), ("bandHaloRows", ctypes.c_uint32)]


class AsvgfError(RuntimeError):
    def __init__(self, code, where):
        L = lib()
        self.code = code
        super().__init__("%s: %s (%s)" % (where, L.asvgfResultString(code).decode(),
                                          L.asvgfLastErrorMessage().decode()))


def lib():
    """Loads libasvgf.so; raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError("libasvgf.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                              "the A-SVGF path has no CPU fallback")
        L = ctypes.CDLL(_LIB_PATH)
        if hasattr(L, "asvgfHostBuildMarker") and not _TEST_ONLY_ALLOW_HOST_BUILD:
            # the test suite's g++ build of the library source (tests/hostsim) is a checker, like the oracle: never a way to run the product
            raise ImportError("%s is the test suite's host build of the library source, not libasvgf.so: the A-SVGF path has no "
                              "CPU fallback" % _LIB_PATH)
        vp, u32, sz = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t
        pu32 = ctypes.POINTER(u32)
        puni = ctypes.POINTER(layout.GlobalUniform)
        L.asvgfCreate.argtypes = [ctypes.POINTER(CreateInfo), ctypes.POINTER(vp)]
        L.asvgfDestroy.argtypes = [vp]
        L.asvgfGetImageInfo.argtypes = [u32, u32, u32, pu32, pu32, pu32, pu32]
        L.asvgfGetImage.argtypes = [vp, u32, u32, ctypes.POINTER(vp), ctypes.POINTER(sz), pu32, pu32, pu32]
        L.asvgfBindExternalImage.argtypes = [vp, u32, vp, sz]
        L.asvgfUploadImage.argtypes = [vp, vp, u32, u32, vp, sz]
        L.asvgfDownloadImage.argtypes = [vp, vp, u32, u32, vp, sz]
        L.asvgfResetHistory.argtypes = [vp, vp]
        L.asvgfGradientSamples.argtypes = [vp, vp, u32, puni]
        L.asvgfDenoise.argtypes = [vp, vp, u32, puni]
        L.asvgfDenoisePass.argtypes = [vp, vp, u32, puni, u32]
        L.asvgfSynchronize.argtypes = [vp, vp]
        for name in ("asvgfResolveCheckerboard", "asvgfCalculateExposure", "asvgfLuminanceHistogram", "asvgfLuminanceAverage"):
            getattr(L, name).argtypes = [vp, vp, u32, puni]
        L.asvgfScatterAccumulate.argtypes = [vp, vp, u32, puni, ctypes.c_float]
        L.asvgfPrepareFinal.argtypes = [vp, vp, u32, puni, ctypes.c_void_p]
        L.asvgfGetStateSize.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
        L.asvgfSaveState.argtypes = [vp, vp, ctypes.c_void_p, ctypes.c_size_t]
        L.asvgfLoadState.argtypes = [vp, vp, ctypes.c_void_p, ctypes.c_size_t]
        L.asvgfGetTonemappingBuffer.argtypes = [vp, ctypes.POINTER(vp)]
        L.asvgfUploadTonemapping.argtypes = [vp, vp, ctypes.POINTER(layout.Tonemapping)]
        L.asvgfDownloadTonemapping.argtypes = [vp, vp, ctypes.POINTER(layout.Tonemapping)]
        L.asvgfGetLaunchCount.argtypes = [vp]
        L.asvgfGetLaunchCount.restype = ctypes.c_uint64
        L.asvgfPassName.argtypes = [u32]
        L.asvgfPassName.restype = ctypes.c_char_p
        L.asvgfBandApronRows.argtypes = [u32, u32]
        L.asvgfBandApronRows.restype = ctypes.c_int32
        L.asvgfPassKernelName.argtypes = [u32, u32]
        L.asvgfPassKernelName.restype = ctypes.c_char_p
        L.asvgfPassBytesPerPixel.argtypes = [u32, u32]
        L.asvgfPassBytesPerPixel.restype = ctypes.c_double
        L.asvgfBandSetCreate.argtypes = [ctypes.POINTER(BandSetCreateInfo), ctypes.POINTER(vp)]
        L.asvgfBandSetDestroy.argtypes = [vp]
        L.asvgfBandSetGetBandCount.argtypes = [vp]
        L.asvgfBandSetGetBandCount.restype = u32
        L.asvgfBandSetGetBand.argtypes = [vp, u32, ctypes.POINTER(vp), ctypes.POINTER(vp), pu32, pu32]
        L.asvgfBandSetFrame.argtypes = [vp, u32, puni]
        L.asvgfBandSetSynchronize.argtypes = [vp]
        L.asvgfBandSetBeginTiming.argtypes = [vp]
        L.asvgfBandSetEndTiming.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
        L.asvgfImportExternalMemoryFd.argtypes = [ctypes.c_int32, ctypes.c_int, ctypes.c_uint64, u32, ctypes.POINTER(vp)]
        L.asvgfExternalMemoryGetBuffer.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(vp)]
        L.asvgfDestroyExternalMemory.argtypes = [vp]
        L.asvgfImportExternalSemaphoreFd.argtypes = [ctypes.c_int32, ctypes.c_int, u32, ctypes.POINTER(vp)]
        L.asvgfWaitExternalSemaphore.argtypes = [vp, vp, ctypes.c_uint64]
        L.asvgfSignalExternalSemaphore.argtypes = [vp, vp, ctypes.c_uint64]
        L.asvgfDestroyExternalSemaphore.argtypes = [vp]
        L.asvgfResultString.argtypes = [ctypes.c_int]
        L.asvgfResultString.restype = ctypes.c_char_p
        L.asvgfLastErrorMessage.restype = ctypes.c_char_p
        for name in EXPORTED_SYMBOLS:
            f = getattr(L, name)
            if f.restype is ctypes.c_int and name not in ("asvgfGetLaunchCount",):
                f.restype = ctypes.c_int
        _lib = L
    return _lib


def _check(code, where):
    if code != SUCCESS:
        raise AsvgfError(code, where)


def pass_bytes_per_pixel(p, iterations=4):
    return lib().asvgfPassBytesPerPixel(p, iterations)


def band_apron_rows(iterations, p):
    """ASVGF_CREATE_BAND_APRON: rows a pass computes beyond the owned band (gradient rows for the gradient passes)."""
    return lib().asvgfBandApronRows(iterations, p)


def chain_bytes_per_pixel(iterations=4, with_gradient_samples=True):
    """Algorithmic HBM bytes per pixel of one frame (408 for the reference chain)."""
    first = 0 if with_gradient_samples else 1
    return sum(pass_bytes_per_pixel(p, iterations) for p in range(first, layout.PASS_ATROUS_0 + iterations))


class Denoiser:
    """One view stream (= one RTGL1 instance's Denoiser + its Framebuffers)."""

    def __init__(self, width, height, atrous_iterations=4, flags=0, device=-1, band=None, halo_rows=0, _borrowed=None):
        self._h = ctypes.c_void_p()
        self._owns = _borrowed is None
        self.width, self.height, self.iterations = width, height, atrous_iterations
        if _borrowed is not None:          # a band handle that belongs to a BandSet
            self._h = ctypes.c_void_p(_borrowed)
            return
        info = CreateInfo(width, height, atrous_iterations, flags, device,
                          band[0] if band else 0, band[1] if band else 0, halo_rows)
        _check(lib().asvgfCreate(ctypes.byref(info), ctypes.byref(self._h)), "asvgfCreate")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if self._owns:
                lib().asvgfDestroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def _index(name):
        return layout.INDEX[name] if isinstance(name, str) else int(name)

    def image(self, name, frame_index):
        """(device pointer, pitch bytes, width, first stored row, stored rows), cur/prev swap applied."""
        p, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        w, r0, rows = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        _check(lib().asvgfGetImage(self._h, self._index(name), frame_index, ctypes.byref(p), ctypes.byref(pitch),
                                   ctypes.byref(w), ctypes.byref(r0), ctypes.byref(rows)), "asvgfGetImage")
        return p.value, pitch.value, w.value, r0.value, rows.value

    def bind_external(self, name, device_ptr, pitch_bytes):
        _check(lib().asvgfBindExternalImage(self._h, self._index(name), device_ptr, pitch_bytes),
               "asvgfBindExternalImage")

    def upload(self, name, frame_index, host, stream=0):
        """host: C-contiguous numpy array (or a host pointer int) holding the stored rows, tightly packed."""
        ptr = host if isinstance(host, int) else host.ctypes.data
        _check(lib().asvgfUploadImage(self._h, stream, self._index(name), frame_index, ptr, 0), "asvgfUploadImage")

    def download(self, name, frame_index, stream=0, out=None):
        idx = self._index(name)
        _, _, w, r0, rows = self.image(idx, frame_index)
        dtype, shape = layout.image_numpy_shape(idx, self.width, self.height)
        shape = (rows,) + tuple(shape[1:])
        if out is None:
            out = np.empty(shape, dtype=dtype)
        ptr = out if isinstance(out, int) else out.ctypes.data
        _check(lib().asvgfDownloadImage(self._h, stream, idx, frame_index, ptr, 0), "asvgfDownloadImage")
        if not isinstance(out, int):
            self.synchronize(stream)
        return out

    def save_state(self, stream=0):
        """The temporal state (history images + exposure) as a numpy byte array (asvgfSaveState)."""
        n = ctypes.c_size_t()
        _check(lib().asvgfGetStateSize(self._h, ctypes.byref(n)), "asvgfGetStateSize")
        buf = np.empty(n.value, dtype=np.uint8)
        _check(lib().asvgfSaveState(self._h, stream, buf.ctypes.data, buf.nbytes), "asvgfSaveState")
        self.synchronize(stream)
        return buf

    def load_state(self, buf, stream=0):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        _check(lib().asvgfLoadState(self._h, stream, buf.ctypes.data, buf.nbytes), "asvgfLoadState")
        self.synchronize(stream)

    def reset_history(self, stream=0):
        _check(lib().asvgfResetHistory(self._h, stream), "asvgfResetHistory")

    def gradient_samples(self, frame_index, uniform, stream=0):
        _check(lib().asvgfGradientSamples(self._h, stream, frame_index, ctypes.byref(uniform)), "asvgfGradientSamples")

    def denoise(self, frame_index, uniform, stream=0):
        _check(lib().asvgfDenoise(self._h, stream, frame_index, ctypes.byref(uniform)), "asvgfDenoise")

    def denoise_pass(self, p, frame_index, uniform, stream=0):
        _check(lib().asvgfDenoisePass(self._h, stream, frame_index, ctypes.byref(uniform), p), "asvgfDenoisePass")

    # ---- the two dispatches after the denoiser ----
    def resolve_checkerboard(self, frame_index, uniform, stream=0):
        """= ImageComposition::PrepareForRaster (Source/ImageComposition.cpp:87-92): PreFinal -> Final."""
        _check(lib().asvgfResolveCheckerboard(self._h, stream, frame_index, ctypes.byref(uniform)), "asvgfResolveCheckerboard")

    def prepare_final(self, frame_index, uniform, lpm_control, stream=0):
        """lpm_control: the 96 uint32 of the host-side LPM setup (ImageComposition::SetupLpmParams)."""
        lpm = np.ascontiguousarray(lpm_control, dtype=np.uint32)
        if lpm.size != 96:
            raise ValueError("the LPM control block holds 24 x uvec4")
        _check(lib().asvgfPrepareFinal(self._h, stream, frame_index, ctypes.byref(uniform), lpm.ctypes.data), "asvgfPrepareFinal")

    def calculate_exposure(self, frame_index, uniform, stream=0):
        """= Tonemapping::CalculateExposure (Source/Tonemapping.cpp:59-175): histogram + eye adaptation."""
        _check(lib().asvgfCalculateExposure(self._h, stream, frame_index, ctypes.byref(uniform)), "asvgfCalculateExposure")

    def luminance_histogram(self, frame_index, uniform, stream=0):
        _check(lib().asvgfLuminanceHistogram(self._h, stream, frame_index, ctypes.byref(uniform)), "asvgfLuminanceHistogram")

    def luminance_average(self, frame_index, uniform, stream=0):
        _check(lib().asvgfLuminanceAverage(self._h, stream, frame_index, ctypes.byref(uniform)), "asvgfLuminanceAverage")

    def scatter_accumulate(self, frame_index, uniform, max_history_length, stream=0):
        """= Volumetric::ProcessScattering (Source/Volumetric.cpp:71-195), CmScatterAccum.comp."""
        _check(lib().asvgfScatterAccumulate(self._h, stream, frame_index, ctypes.byref(uniform), max_history_length),
               "asvgfScatterAccumulate")

    def tonemapping(self, stream=0):
        """Host copy of the handle's AsvgfTonemapping (histogram[256], avgLuminance)."""
        t = layout.Tonemapping()
        _check(lib().asvgfDownloadTonemapping(self._h, stream, ctypes.byref(t)), "asvgfDownloadTonemapping")
        self.synchronize(stream)
        return t

    def set_tonemapping(self, t, stream=0):
        _check(lib().asvgfUploadTonemapping(self._h, stream, ctypes.byref(t)), "asvgfUploadTonemapping")
        self.synchronize(stream)

    def synchronize(self, stream=0):
        _check(lib().asvgfSynchronize(self._h, stream), "asvgfSynchronize")

    @property
    def launch_count(self):
        return lib().asvgfGetLaunchCount(self._h)


class BandSet:
    """One frame split into horizontal bands across GPUs (include/asvgf.h, asvgfBandSet*): a band handle, a stream and the
    halo pushes per entry of `devices` (a device may repeat: several bands on one GPU)."""

    def __init__(self, width, height, devices, atrous_iterations=4, flags=0, boundaries=None, halo_rows=0):
        self._h = ctypes.c_void_p()
        self.width, self.height, self.iterations = width, height, atrous_iterations
        n = len(devices)
        dev = (ctypes.c_int32 * n)(*devices)
        bnd = (ctypes.c_uint32 * (n + 1))(*boundaries) if boundaries is not None else None
        info = BandSetCreateInfo(width, height, atrous_iterations, flags, n, dev,
                                 ctypes.cast(bnd, ctypes.POINTER(ctypes.c_uint32)) if bnd is not None else None, halo_rows)
        _check(lib().asvgfBandSetCreate(ctypes.byref(info), ctypes.byref(self._h)), "asvgfBandSetCreate")
        self.bands = []
        for k in range(n):
            h, st, r0, r1 = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_uint32()
            _check(lib().asvgfBandSetGetBand(self._h, k, ctypes.byref(h), ctypes.byref(st), ctypes.byref(r0), ctypes.byref(r1)),
                   "asvgfBandSetGetBand")
            d = Denoiser(width, height, atrous_iterations, _borrowed=h.value)
            self.bands.append({"denoiser": d, "stream": st.value or 0, "rows": (r0.value, r1.value), "device": devices[k]})

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            for b in self.bands:
                b["denoiser"].close()
            lib().asvgfBandSetDestroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def stored_rows(self, k, name, frame_index=0):
        """(first stored row, number of stored rows) of an image of band k: its rows plus the halo."""
        _, _, _, r0, rows = self.bands[k]["denoiser"].image(name, frame_index)
        return r0, rows

    def upload(self, name, frame_index, full_image):
        """full_image: numpy array of the WHOLE frame; every band gets the rows it stores, on its own stream."""
        for k, b in enumerate(self.bands):
            r0, rows = self.stored_rows(k, name, frame_index)
            b["denoiser"].upload(name, frame_index, np.ascontiguousarray(full_image[r0:r0 + rows]), stream=b["stream"])
        # (the host arrays above are temporaries: finish the copies before they go away)
        for b in self.bands:
            lib().asvgfSynchronize(b["denoiser"]._h, b["stream"])

    def download(self, name, frame_index):
        """The whole frame, assembled from the rows every band OWNS."""
        dtype, shape = layout.image_numpy_shape(name, self.width, self.height)
        third = layout.IMAGES[name][2] if isinstance(name, str) else False
        out = np.zeros(shape, dtype=dtype)
        for k, b in enumerate(self.bands):
            r0, rows = self.stored_rows(k, name, frame_index)
            a = b["denoiser"].download(name, frame_index, stream=b["stream"])
            o0, o1 = b["rows"]
            if third:
                o0, o1 = o0 // 3, (shape[0] if o1 == self.height else o1 // 3)
            out[o0:o1] = a[o0 - r0:o1 - r0]
        return out

    def frame(self, frame_index, uniform):
        """= CalculateGradientsSamples + Denoiser::Denoise of the whole frame (enqueues only)."""
        _check(lib().asvgfBandSetFrame(self._h, frame_index, ctypes.byref(uniform)), "asvgfBandSetFrame")

    def synchronize(self):
        _check(lib().asvgfBandSetSynchronize(self._h), "asvgfBandSetSynchronize")

    def begin_timing(self):
        _check(lib().asvgfBandSetBeginTiming(self._h), "asvgfBandSetBeginTiming")

    def end_timing(self):
        """(longest band interval, [each band's interval]) in milliseconds since begin_timing()."""
        worst = ctypes.c_float()
        per = (ctypes.c_float * len(self.bands))()
        _check(lib().asvgfBandSetEndTiming(self._h, ctypes.byref(worst), per), "asvgfBandSetEndTiming")
        return worst.value, list(per)


class ExternalMemory:
    """CUDA half of the Vulkan interop: an exported memory object (opaque fd) imported and mapped as linear device memory."""

    def __init__(self, fd, size, device=-1, dedicated=False):
        self._h = ctypes.c_void_p()
        self.size = size
        _check(lib().asvgfImportExternalMemoryFd(device, fd, size, 1 if dedicated else 0, ctypes.byref(self._h)),
               "asvgfImportExternalMemoryFd")

    def buffer(self, offset=0, size=None):
        p = ctypes.c_void_p()
        _check(lib().asvgfExternalMemoryGetBuffer(self._h, offset, self.size - offset if size is None else size, ctypes.byref(p)),
               "asvgfExternalMemoryGetBuffer")
        return p.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().asvgfDestroyExternalMemory(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close


class ExternalSemaphore:
    """An exported Vulkan semaphore (binary or timeline, fd) imported for stream-ordered wait / signal."""

    def __init__(self, fd, device=-1, timeline=False):
        self._h = ctypes.c_void_p()
        _check(lib().asvgfImportExternalSemaphoreFd(device, fd, 1 if timeline else 0, ctypes.byref(self._h)),
               "asvgfImportExternalSemaphoreFd")

    def wait(self, stream=0, value=0):
        _check(lib().asvgfWaitExternalSemaphore(self._h, ctypes.c_void_p(stream), value), "asvgfWaitExternalSemaphore")

    def signal(self, stream=0, value=0):
        _check(lib().asvgfSignalExternalSemaphore(self._h, ctypes.c_void_p(stream), value), "asvgfSignalExternalSemaphore")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().asvgfDestroyExternalSemaphore(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close
