"""Cases that only make sense against tests/hostsim/libasvgf_hostsim.so (the library's source compiled for the host): sizes small
enough for one host thread per CUDA thread.  Not collected by the normal runs (the file name does not match test_*.py);
tests/test_hostsim.py passes it to a child pytest that has ASVGF_TEST_HOSTSIM=1 set (tests/conftest.py)."""
import os

import numpy as np
import pytest

from _harness import Pair, assert_all_ok
from test_bandset_gpu import _run as run_band_set
from test_parity_gpu import OUTPUTS_PER_PASS

pytestmark = pytest.mark.skipif(os.environ.get("ASVGF_TEST_HOSTSIM") != "1", reason="host-compiled library only")


def test_the_library_under_test_is_the_host_build():
    from raytracedgl1_b200 import denoiser
    assert denoiser._LIB_PATH.endswith("libasvgf_hostsim.so")


@pytest.mark.parametrize("size", [(130, 49), (64, 66)])
def test_per_pass_parity_from_identical_state(size):
    """Every dispatch of the chain from the oracle's state (as tests/test_parity_gpu.py does at 320x180 ... 1280x720 on the GPU): the
    outputs of that pass within the single-pass tolerance, the integer images bit-exact."""
    W, H = size
    p = Pair(W, H)
    try:
        for frame in range(3):
            fi = frame & 1
            u = p.feed_inputs(frame, fi)
            for pas in range(12):
                p.sync_state_to_cuda()
                p.oracle.run_pass(pas, fi, u)
                p.cuda.denoise_pass(pas, fi, u)
                p.cuda.synchronize()
                res = p.compare(fi, OUTPUTS_PER_PASS[pas], whole_frame=False)
                assert_all_ok(res, "%dx%d frame %d pass %d" % (W, H, frame, pas))
                if pas in (0, 1, 2, 3, 4):      # the gradient chain is integer work end to end
                    for name in OUTPUTS_PER_PASS[pas]:
                        assert res[name]["exact_frac"] == 1.0, (name, res[name])
    finally:
        p.close()


@pytest.mark.parametrize("bands,boundaries,size", [(2, None, (64, 138)), (3, [0, 66, 135, 201], (34, 201)), (1, None, (34, 40))])
def test_band_set_is_bit_identical_to_the_single_handle(bands, boundaries, size):
    """asvgfBandSet* (apron schedule, the four push points, per-band OpenMP enqueue) with the bands on different pretend devices."""
    run_band_set(size[0], size[1], list(range(bands)), boundaries, frames=4)


def run_tight_external(W, H, iterations=4, extra=0):
    """Every image caller-owned (asvgfBindExternalImage) with a TIGHT row pitch -- row bytes rounded up to 16, plus `extra` x 16..48
    bytes that differ from image to image -- instead of the library's align256(width) texels: what a Vulkan host's linear images
    look like.  The images then hold different numbers of texels per row, which selects the UNIFORM = false forms of the temporal
    and a-trous kernels (one offset per image instead of one shared texel offset).  Bit-identical to the owned-image handle."""
    import ctypes
    import synth
    from raytracedgl1_b200 import denoiser as dn, layout
    own = dn.Denoiser(W, H, iterations, flags=dn.CREATE_EXPORT_DEBUG_MASK)
    ext = dn.Denoiser(W, H, iterations, flags=dn.CREATE_EXTERNAL_IMAGES | dn.CREATE_EXPORT_DEBUG_MASK)
    keep = {}
    try:
        for name in layout.IMAGES:
            idx = layout.INDEX[name]
            f, w, h, b = [ctypes.c_uint32() for _ in range(4)]
            assert dn.lib().asvgfGetImageInfo(idx, W, H, ctypes.byref(f), ctypes.byref(w), ctypes.byref(h), ctypes.byref(b)) == 0
            pitch = ((w.value * b.value + 15) // 16) * 16 + 16 * extra * (1 + idx % 3)
            buf = np.zeros(pitch * h.value + 16, dtype=np.uint8)     # (under the host build "device memory" is the heap)
            keep[name] = buf
            ext.bind_external(name, buf.ctypes.data + (-buf.ctypes.data) % 16, pitch)
        for frame in range(3):
            fi = frame & 1
            g, u = synth.generate(W, H, frame), synth.uniform(W, H, frame)
            for d in (own, ext):
                for name, arr in g.items():
                    d.upload(name, fi, arr)
                d.gradient_samples(fi, u)
                d.denoise(fi, u)
                d.synchronize()
            for name in layout.IMAGES:
                a, b = own.download(name, fi), ext.download(name, fi)
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (W, H, frame, name)
    finally:
        own.close()
        ext.close()


@pytest.mark.parametrize("size,extra", [((34, 10), 0), ((66, 35), 1), ((130, 49), 1)])
def test_tight_and_unequal_external_pitches(size, extra):
    run_tight_external(size[0], size[1], extra=extra)


def test_motion_beyond_the_halo_is_reported():
    from raytracedgl1_b200 import denoiser as dn
    W, H = 34, 240

    def fast_motion(frame, g):
        g["Motion"][..., 1] = np.float16(100.0 / H)

    with pytest.raises(dn.AsvgfError) as e:
        run_band_set(W, H, [0, 1], None, frames=2, mutate=fast_motion)
    assert e.value.code == dn.HALO_EXCEEDED


def _edit(case):
    def general_projection(u):       # entries outside m0, m5, m8..m11, m14: the full mat4 products of the virtual-point reprojection
        u.projection[1], u.projection[4], u.projection[12], u.projection[13] = 0.02, -0.015, 0.03, -0.01
        u.projectionPrev[1], u.projectionPrev[4], u.projectionPrev[12], u.projectionPrev[13] = 0.021, -0.014, 0.031, -0.012

    def time_delta_zero(u):
        u.timeDelta = 0.0

    def time_delta_negative(u):
        u.timeDelta = -1.0

    def time_delta_large(u):
        u.timeDelta = 10.0

    def coverage_one(u):
        u.lightmapScreenCoverage = 1.0

    def gradient_multipliers(u):
        u.gradientMultDiffuse, u.gradientMultIndirect, u.gradientMultSpecular = 5.0, 0.0, 100.0

    def camera_jump(u):
        for i in range(3):
            u.cameraPositionPrev[i] = u.cameraPosition[i] + 1000.0

    def all_debug_flags(u):
        u.debugShowFlags = 0xffffffff & ~(1 << 8)
    return locals()[case]


@pytest.mark.parametrize("case", ["general_projection", "time_delta_zero", "time_delta_negative", "time_delta_large", "coverage_one",
                                  "gradient_multipliers", "camera_jump", "all_debug_flags"])
def test_uniform_edges(case):
    """Uniform values no GPU test sets (the general-projection branch of the temporal pass among them): whole frames against the oracle."""
    p = Pair(66, 35)
    try:
        for frame in range(3 if case == "general_projection" else 2):
            fi, u = p.run_frame(frame, uniform_edit=_edit(case))
            res = p.compare(fi)
            assert_all_ok(res, "%s frame %d" % (case, frame))
            if case == "general_projection" and frame:
                assert int(((p.oracle.image("DebugMask", fi) >> 5) & 1).sum()) > 500     # the virtual-point branch is taken
            p.sync_state_to_cuda()
    finally:
        p.close()


def test_call_orders():
    """asvgfDenoise without asvgfGradientSamples (the host may keep its own gradient-sample pass), and
    asvgfResetHistory in the middle of a sequence."""
    p = Pair(66, 35)
    try:
        for frame in range(3):
            fi = frame & 1
            u = p.feed_inputs(frame, fi)
            p.oracle.denoise(fi, u)
            p.cuda.denoise(fi, u)
            p.cuda.synchronize()
            assert_all_ok(p.compare(fi), "denoise only, frame %d" % frame)
    finally:
        p.close()
    p = Pair(66, 35)
    try:
        for frame in range(4):
            if frame == 2:
                p.oracle.reset_history()
                p.cuda.reset_history()
            fi, u = p.run_frame(frame)
            assert_all_ok(p.compare(fi), "reset at 2, frame %d" % frame)
    finally:
        p.close()
