"""Pins the hand-written oracle (oracle/asvgf_oracle.cpp) to the reference itself.

oracle/_ref executes the UNMODIFIED text of the reference's compute shaders (/root/reference/Source/Shaders/Cm*.comp,
compiled as C++ over the reference's vendored GLM by oracle/ref_glsl/build_ref.py).  Here every dispatch of
Denoiser::Denoise (Source/Denoiser.cpp:87-254) and the post dispatches are run by both the oracle and the reference
shaders from IDENTICAL state, and every image of the table must come out bit-identical: integer images, packed colours,
fp16 -- no tolerance.  (The gradient-sample stage RtGradients.rgen is a ray-generation shader; it has no compute form
and stays pinned by its literal restatement and KATs only.)

These tests need /root/reference (this container).  On the GPU box they are skipped and the committed fixtures that the
same executor generated (tests/golden/asvgf_ref_96x54.npz, tests/test_golden.py) carry the pin.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle as oracle_mod
import synth
from oracle import ref
from raytracedgl1_b200 import layout

pytestmark = pytest.mark.skipif(not (ref.available() or os.path.isdir("/root/reference/Source/Shaders")),
                                reason="the reference tree is not present: oracle/_ref cannot be built here")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not ref.available():
        ref.build()


def _lock(r, o):
    for n in layout.IMAGES:
        r.image(n, 0)[...] = o.image(n, 0)
    ctypes.memmove(ctypes.byref(r.tonemapping), ctypes.byref(o.tonemapping), ctypes.sizeof(layout.Tonemapping))


def _assert_tables_equal(o, r, context):
    for n in layout.IMAGES:
        if n == "DebugMask":      # this repo's export of the temporal pass's decisions; the shader has no such image
            continue
        a, b = o.image(n, 0), r.image(n, 0)
        if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
            bad = np.argwhere((a.view(np.uint8) != b.view(np.uint8)).reshape(a.shape[0], a.shape[1], -1).any(axis=-1))
            raise AssertionError("%s: %s differs in %d texels, first at (y, x) = %s" % (context, n, len(bad), bad[0]))


def _run_sequence(W, H, frames, edit=None, mutate=None):
    o = oracle_mod.Oracle(W, H, 4)
    r = ref.RefDenoiser(W, H)
    try:
        for frame in range(frames):
            fi = frame & 1
            g = synth.generate(W, H, frame)
            rng = np.random.default_rng(99 + frame)
            g["Throughput"][..., 3] = (rng.random((H, W)) < 0.4).astype(np.float16)       # split pixels: the resolve has work to do
            for name in ("ScreenEmisRT", "AcidFogRT"):                                    # finite random B10G11R11 texels
                g[name] = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) & np.uint32(0x7FDFFBFF)
            if mutate:
                mutate(frame, g)
            for n, a in g.items():
                o.image(n, fi)[...] = a
            u = synth.uniform(W, H, frame)
            u.volumeEnableType, u.volumeScattering = 1, 1500.0
            # the final-prepare uniforms at the reference's defaults (Source/DrawFrameInfo.h:103,118-119,177)
            u.luminanceWhitePoint, u.bloomThreshold, u.bloomEmissionMultiplier, u.emissionMaxScreenColor = 10.0, 4.0, 16.0, 1.5
            if edit:
                edit(u)
            o.gradient_samples(fi, u)
            _lock(r, o)
            for p in range(1, 12):
                o.run_pass(p, fi, u)
                r.run_pass(p, fi, u)
                _assert_tables_equal(o, r, "%dx%d frame %d pass %d" % (W, H, frame, p))
            for call in ("resolve_checkerboard", "luminance_histogram"):
                getattr(o, call)(fi, u)
                getattr(r, call)(fi, u)
            assert list(o.tonemapping.histogram) == list(r.tonemapping.histogram), "frame %d histogram" % frame
            o.luminance_average(u)
            r.luminance_average(u)
            assert np.float32(o.tonemapping.avgLuminance).tobytes() == np.float32(r.tonemapping.avgLuminance).tobytes()
            o.scatter_accumulate(fi, u, 16.0)
            r.scatter_accumulate(fi, u, 16.0)
            _assert_tables_equal(o, r, "%dx%d frame %d post passes" % (W, H, frame))
            # CmPrepareFinal.comp with the LPM control block of the reference's own host-side setup (defaults, and a tinted one)
            lpm = ref.lpm_setup() if frame % 2 == 0 else ref.lpm_setup((0.1, -0.2, 0.05), (1.0, 0.5, 0.25))
            o.prepare_final(fi, u, lpm)
            r.prepare_final(fi, u, lpm)
            _assert_tables_equal(o, r, "%dx%d frame %d final prepare" % (W, H, frame))
    finally:
        o.close()
        r.close()


@pytest.mark.parametrize("size,frames", [((320, 180), 8), ((322, 182), 4), ((66, 35), 3), ((34, 10), 2), ((6, 7), 2)])
def test_every_dispatch_bit_identical_to_the_reference_shaders(size, frames):
    """322x182: W/3, W/2 and H/3 do not divide evenly; the tiny sizes are smaller than one work group."""
    _run_sequence(size[0], size[1], frames)


def test_light_switch_frames():
    """Frames 22..26 span the synthetic light switch (non-zero anti-lag gradients feed the temporal pass)."""
    W, H = 160, 90
    o = oracle_mod.Oracle(W, H, 4)
    r = ref.RefDenoiser(W, H)
    try:
        for frame in range(22, 27):
            fi = frame & 1
            for n, a in synth.generate(W, H, frame).items():
                o.image(n, fi)[...] = a
            u = synth.uniform(W, H, frame)
            o.gradient_samples(fi, u)
            _lock(r, o)
            o.denoise(fi, u)
            r.denoise(fi, u)
            _assert_tables_equal(o, r, "frame %d" % frame)
        assert o.image("DISPingGradient", 0).max() > 0 or o.image("DISPingGradient", 1).max() > 0
    finally:
        o.close()
        r.close()


@pytest.mark.parametrize("case", ["antifirefly_off", "debug_white_albedo", "classic_shading_half", "debug_only_specular",
                                  "debug_only_direct", "debug_only_indirect", "volume_none", "debug_motion", "debug_gradients",
                                  "debug_unfiltered_diffuse", "debug_unfiltered_specular", "debug_unfiltered_indirect", "debug_normals"])
def test_uniform_edges(case):
    def edit(u):
        if case == "antifirefly_off":
            u.antiFireflyEnabled = 0
        if case == "debug_white_albedo":
            u.debugShowFlags = 1 << 9
        if case == "debug_only_direct":
            u.debugShowFlags = 1 << 5
        if case == "debug_only_specular":
            u.debugShowFlags = 1 << 6
        if case == "debug_only_indirect":
            u.debugShowFlags = 1 << 7
        if case == "classic_shading_half":
            u.lightmapScreenCoverage = 0.5
        if case == "volume_none":
            u.volumeEnableType = 0
        # the debug views of CmPrepareFinal.comp:175-226
        flag = {"debug_motion": 0, "debug_gradients": 1, "debug_unfiltered_diffuse": 2, "debug_unfiltered_specular": 3,
                "debug_unfiltered_indirect": 4, "debug_normals": 10}.get(case)
        if flag is not None:
            u.debugShowFlags = 1 << flag
    _run_sequence(128, 72, 3, edit=edit)


def test_degenerate_inputs():
    """All-sky frame, NaN / Inf / negative garbage in the float inputs, random packed normals, huge motion vectors."""
    def mutate(frame, g):
        if frame == 1:
            g["IsSky"][...] = 1
            g["DepthWorld"][...] = np.float16(20000.0)
            g["Normal"][...] = 0x80000000
        if frame in (2, 3):
            rng = np.random.default_rng(7)
            m = g["Motion"]
            m[::7, ::5, 0] = np.float16(np.inf)
            m[::5, ::7, 1] = np.float16(-60000.0)
            m[3::11, ::3, 2] = np.float16(np.nan)
            g["DepthWorld"][::9, ::4] = np.float16(-5.0)
            g["DepthWorld"][1::9, ::4] = np.float16(0.0)
            g["DepthGrad"][::3, ::8] = np.float16(np.nan)
            g["SurfacePosition"][::6, ::6, :3] = 0.0
            g["ViewDirection"][::4, ::9, 3] = np.float16(np.inf)
            g["UnfilteredDirect"][::10, ::10] = 0xFFFFFFFF
            g["Normal"][::5, ::5] = rng.integers(0, 2 ** 32, size=g["Normal"][::5, ::5].shape, dtype=np.uint64).astype(np.uint32)
    _run_sequence(128, 72, 5, mutate=mutate)
