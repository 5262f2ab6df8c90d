#!/usr/bin/env python
"""Regenerates tests/golden/asvgf_ref_96x54.npz: outputs of the REFERENCE'S OWN SHADERS for four frames of the seeded
synthetic sequence at 96x54.

The images are produced by oracle/_ref -- the unmodified text of /root/reference/Source/Shaders/Cm*.comp compiled as
C++ over the reference's vendored GLM (oracle/ref_glsl/build_ref.py) and dispatched in the order of
Source/Denoiser.cpp:87-254 -- free-running over the sequence, i.e. every frame's history is itself reference output.
Only the gradient-sample stage (RtGradients.rgen, a ray-generation shader with no compute form) is taken from the
hand-written oracle.  These fixtures therefore PIN the oracle and the CUDA path to the reference: tests/test_golden.py
requires the oracle to reproduce them bit for bit (CPU) and the CUDA path to match them within tolerance (GPU),
also where /root/reference does not exist.

    python tests/golden/make_golden_ref.py        (needs /root/reference; rebuilds oracle/_ref)
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402
from raytracedgl1_b200 import layout  # noqa: E402

W, H, N, FRAMES = 96, 54, 4, 4
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "asvgf_ref_96x54.npz")
# everything the chain writes (state, scratch and outputs), plus the post passes' outputs
NAMES = [n for n, info in layout.IMAGES.items() if info[3] != "in" and n not in ("DebugMask", "FwdLum", "ScatterCurrent")]
POST = ["Final", "ScreenEmission", "AcidFog", "Scattering", "ScatteringHistory", "BloomInput"]
# LPM control blocks of the reference's host-side setup (oracle.ref.lpm_setup = AMD's LpmSetup as ImageComposition::SetupLpmParams
# calls it): RTGL1's default tonemapping parameters, and a tinted set; stored in the fixture so that they travel
LPM_PARAMS = {"lpm_default": ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), "lpm_tinted": ((0.1, -0.2, 0.05), (1.0, 0.5, 0.25))}


def edit_uniform(u):
    u.volumeEnableType, u.volumeScattering = 1, 1500.0           # VOLUME_ENABLE_SIMPLE
    u.volumeAmbient[0], u.volumeAmbient[1], u.volumeAmbient[2] = 0.7, 0.8, 1.0
    # CmPrepareFinal.comp's uniforms at the reference's defaults (Source/DrawFrameInfo.h:103,118-119,177); a classic-shading strip
    u.luminanceWhitePoint, u.bloomThreshold, u.bloomEmissionMultiplier, u.emissionMaxScreenColor = 10.0, 4.0, 16.0, 1.5


def inputs(frame):
    g = synth.generate(W, H, frame)
    rng = np.random.default_rng(2024 + frame)
    g["Throughput"][..., 3] = (rng.random((H, W)) < 0.4).astype(np.float16)   # split (reflect/refract) pixels
    # the rasteriser's / fog producer's checkerboarded images that the resolve also handles (CmCheckerboard.comp:63-64):
    # random finite B10G11R11 texels (top exponent bit of every channel cleared)
    finite = np.uint32(~((1 << 10) | (1 << 21) | (1 << 31)) & 0xFFFFFFFF)
    for name in ("ScreenEmisRT", "AcidFogRT"):
        g[name] = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) & finite
    return g


def run(side, lpm):
    """side: an object with the Oracle surface (oracle.Oracle or oracle.ref.RefDenoiser); gradient samples come from
    `gradient_source` when the side has none.  lpm: {name: 96 x uint32} control blocks (LPM_PARAMS)."""
    out = {k: np.array(v, dtype=np.uint32) for k, v in lpm.items()}
    grad = oracle.Oracle(W, H, N) if not hasattr(side, "gradient_samples") else None
    for f in range(FRAMES):
        fi = f & 1
        for k, v in inputs(f).items():
            side.image(k, fi)[...] = v
        u = synth.uniform(W, H, f)
        edit_uniform(u)
        if grad is None:
            side.gradient_samples(fi, u)
        else:   # RtGradients.rgen has no compute form: the oracle runs it on a copy of this side's state
            for n in layout.IMAGES:
                grad.image(n, 0)[...] = side.image(n, 0)
            grad.gradient_samples(fi, u)
            for n in ("DISPingGradient", "GradientPrevPix"):
                side.image(n, fi)[...] = grad.image(n, fi)
        side.denoise(fi, u)
        for n in NAMES:
            if n in POST:
                continue
            out["f%d_%s" % (f, n)] = np.array(side.image(n, fi), copy=True)
        side.scatter_accumulate(fi, u, 8.0)
        side.luminance_histogram(fi, u)
        tm = side.tonemapping
        out["f%d_histogram" % f] = np.array(tm.histogram, dtype=np.uint32)
        side.luminance_average(u)
        out["f%d_avgLuminance" % f] = np.array([side.tonemapping.avgLuminance], dtype=np.float32)
        side.resolve_checkerboard(fi, u)
        for n in POST:
            if n != "BloomInput":
                out["f%d_%s" % (f, n)] = np.array(side.image(n, fi), copy=True)
        # ImageComposition::Finalize (CmPrepareFinal.comp): Final is tonemapped in place, the bloom input written
        side.prepare_final(fi, u, lpm["lpm_default" if f % 2 == 0 else "lpm_tinted"])
        out["f%d_FinalPrepared" % f] = np.array(side.image("Final", fi), copy=True)
        out["f%d_BloomInput" % f] = np.array(side.image("BloomInput", fi), copy=True)
    if grad is not None:
        grad.close()
    return out


if __name__ == "__main__":
    from oracle import ref
    oracle.build(); synth.build(); ref.build()
    r = ref.RefDenoiser(W, H)
    np.savez_compressed(PATH, **run(r, {k: ref.lpm_setup(*v) for k, v in LPM_PARAMS.items()}))
    print("wrote", PATH, os.path.getsize(PATH), "bytes")
