"""Analytic known-answer tests of the oracle's passes (CPU only).  The reference ships no golden vectors for this
path, so these pin the restatement against closed-form expectations derived from the shader source."""
import ctypes

import numpy as np
import pytest

import oracle
import synth
from raytracedgl1_b200 import layout

W, H = 256, 120


def e5(r, g, b):
    return oracle.lib().asvgfOracleEncodeE5B9G9R9(r, g, b)


def enc_normal(x, y, z):
    return oracle.lib().asvgfOracleEncodeNormal(x, y, z)


def flat_scene(o, fi, depth=10.0, color=(1.0, 0.5, 0.25), rough=1.0, metallic=0.0):
    """A fronto-parallel flat wall filling the frame, zero motion, history-free."""
    o.image("IsSky", fi)[...] = 0
    o.image("Normal", fi)[...] = enc_normal(0.0, 0.0, 1.0)
    o.image("Normal_Prev", fi)[...] = enc_normal(0.0, 0.0, 1.0)
    o.image("DepthWorld", fi)[...] = np.float16(depth)
    o.image("DepthWorld_Prev", fi)[...] = np.float16(depth)
    o.image("DepthGrad", fi)[...] = np.float16(0.0)
    o.image("Motion", fi)[...] = np.float16(0.0)
    o.image("MetallicRoughness", fi)[...] = np.array([round(metallic * 255), round(rough * 255)], dtype=np.uint8)
    for n in ("UnfilteredDirect", "UnfilteredSpecular", "UnfilteredIndir"):
        o.image(n, fi)[...] = e5(*color)
    o.image("SurfacePosition", fi)[...] = np.array([0, 0, -depth, 0], dtype=np.float32)
    o.image("SurfacePosition_Prev", fi)[...] = np.array([0, 0, -depth, 0], dtype=np.float32)
    o.image("ViewDirection", fi)[...] = np.array([0, 0, -1, 5.0], dtype=np.float16)
    o.image("Throughput", fi)[...] = np.array([1, 1, 1, 0], dtype=np.float16)
    o.image("Albedo", fi)[...] = oracle.lib().asvgfOraclePackUfloat111110(0.5, 0.5, 0.5)
    lum = 0.2125 * color[0] + 0.7154 * color[1] + 0.0721 * color[2]
    o.image("GradientInputs", fi)[...] = np.float16(lum)
    o.image("GradientInputs_Prev", fi)[...] = np.float16(lum)
    o.image("FwdLum", fi)[...] = np.float16(lum)
    u = synth.uniform(W, H, 0)
    u.viewPrev[:] = u.view[:]
    u.projectionPrev[:] = u.projection[:]
    u.cameraPositionPrev[:] = u.cameraPosition[:]
    return u


def test_static_scene_history_grows_by_one_per_frame_and_converges():
    o = oracle.Oracle(W, H)
    # deep interior of both checkerboard halves: the reference's filters blend in out-of-image zero texels at the
    # image border and that darkening creeps inwards by up to 16 px per frame
    rows, cols = slice(40, H - 40), np.r_[40:W // 2 - 40, W // 2 + 40:W - 40]
    for frame in range(8):
        fi = frame & 1
        u = flat_scene(o, fi)
        o.gradient_samples(fi, u)
        o.denoise(fi, u)
        hl = o.image("AccumHistoryLength", fi).astype(np.float32)
        assert np.all(hl[..., 0] == frame + 1) and np.all(hl[..., 1] == frame + 1)       # :339-340, alpha = 0
        assert np.all(o.image("DebugMask", fi) & 0x10)   # every pixel finds consistent history (flat_scene sets *_Prev)
        assert np.all(o.image("DISPingGradient", fi) == 0)
        if frame < 2:
            # a constant field is a fixed point of every filter in the chain
            assert np.all(o.image("DiffAccumColor", fi)[rows][:, cols] == e5(1.0, 0.5, 0.25))
            d = o.image("DiffColorHistory", fi).astype(np.float32)[rows][:, cols]
            assert np.allclose(d[..., :3], [1.0, 0.5, 0.25], rtol=2e-3)
            assert np.all(d[..., 3] <= 1e-3)             # no variance in a constant field
            # composition KAT (CmSVGFAtrous.comp:294-311): rough=1 -> specular := indirect; metallic=0
            pf = layout.decode_b10g11r11(o.image("PreFinal", fi))[rows, 80:W - 80]
            c = np.array([1.0, 0.5, 0.25])
            want = (c + c) * 0.5 + c * 0.04
            assert np.allclose(pf, want, rtol=2.0 ** -5)
            hist = o.image("HistogramInput", fi).astype(np.float32)[rows, 80:W - 80]
            assert np.allclose(hist, layout.luminance(want), rtol=2e-3)
    o.close()


def test_full_disocclusion_resets_history():
    o = oracle.Oracle(W, H)
    u = flat_scene(o, 0)
    o.gradient_samples(0, u); o.denoise(0, u)
    u = flat_scene(o, 1, depth=20.0, color=(0.25, 2.0, 4.0))      # every surface moved 10 units: depth test fails
    o.image("DepthWorld_Prev", 1)[...] = np.float16(10.0)
    o.gradient_samples(1, u); o.denoise(1, u)
    assert np.all(o.image("DebugMask", 1) == 0)
    hl = o.image("AccumHistoryLength", 1).astype(np.float32)
    assert np.all(hl[..., :3] == 1.0) and np.all(hl[..., 3] == 5.0)                      # :310-315, :368-376
    assert np.all(o.image("DiffTemporary", 1) == e5(0.25, 2.0, 4.0))
    assert np.all(o.image("SpecAccumColor", 1) == e5(0.25, 2.0, 4.0))
    lum = np.float32(0.2125) * np.float32(0.25) + np.float32(0.7154) * np.float32(2.0) + np.float32(0.0721) * np.float32(4.0)
    m = o.image("DiffAccumMoments", 1).astype(np.float32)
    assert np.allclose(m[..., 0], lum, rtol=1e-3) and np.allclose(m[..., 1], lum * lum, rtol=1e-3)
    o.close()


def test_antilag_alpha_one_drops_history():
    o = oracle.Oracle(W, H)
    for frame in range(3):
        u = flat_scene(o, frame & 1)
        o.gradient_samples(frame & 1, u); o.denoise(frame & 1, u)
    fi = 1
    u = flat_scene(o, fi, color=(3.0, 3.0, 3.0))
    o.image("DISPingGradient", fi)[...] = 255                                            # alpha_D = alpha_I = alpha_S = 1
    o.run_pass(layout.PASS_TEMPORAL, fi, u)
    # (W=256: the last pixel column reads gradient texel 255/3 = 85, outside the 85-wide gradient image -> alpha 0)
    hl = o.image("AccumHistoryLength", fi).astype(np.float32)[:, :-1]
    assert np.all(hl[..., 0] == 1.0) and np.all(hl[..., 1] == 1.0)                        # len*(1-1)^10 + 1
    assert np.all(o.image("DiffTemporary", fi)[:, :-1] == e5(3.0, 3.0, 3.0))              # mix(prev, cur, 1)
    assert np.all(o.image("IndirAccum", fi)[:, :-1] == e5(3.0, 3.0, 3.0))
    assert np.all(o.image("AccumHistoryLength", fi).astype(np.float32)[:, -1, 0] == 4.0)
    o.close()


def test_antifirefly_clamps_single_outlier_and_is_idempotent_on_flat_input():
    o = oracle.Oracle(W, H)
    u = flat_scene(o, 0)
    for n in ("DiffTemporary", "SpecAccumColor", "IndirAccum"):
        o.image(n, 0)[...] = e5(1.0, 1.0, 1.0)
    o.image("DiffTemporary", 0)[20, 30] = e5(500.0, 1.0, 0.001)
    o.run_pass(layout.PASS_ANTIFIREFLY, 0, u)
    out = o.image("DiffAccumColor", 0)
    inner = out[2:-2, 2:W // 2 - 2]
    assert np.all(inner == e5(1.0, 1.0, 1.0))            # the firefly is clamped into its neighbours' [min,max]
    assert np.all(o.image("SpecPingColor", 0)[2:-2, 2:-2] == e5(1.0, 1.0, 1.0))
    # out-of-image zeros participate (no isInside mask in CmAntiFirefly.comp): border pixels stay within [0,1] = unchanged
    assert np.all(out[0, :] == e5(1.0, 1.0, 1.0))
    u.antiFireflyEnabled = 0
    o.run_pass(layout.PASS_ANTIFIREFLY, 0, u)
    assert o.image("DiffAccumColor", 0)[20, 30] == e5(500.0, 1.0, 0.001)                  # pass-through :109-117
    o.close()


def test_gradient_atrous_of_constant_is_constant_and_routes_history():
    o = oracle.Oracle(W, H)
    u = flat_scene(o, 0)
    o.image("DISPingGradient", 0)[...] = np.array([51, 102, 204, 0], dtype=np.uint8)
    for it in range(4):
        o.run_pass(layout.PASS_GRADIENT_ATROUS_0 + it, 0, u)
    for n in ("DISPingGradient", "DISPongGradient", "DISGradientHistory"):
        assert np.all(o.image(n, 0)[..., :3] == [51, 102, 204]), n                       # result lands in Ping (:103)
    o.close()


def test_gradient_samples_select_brightest_and_never_repeat_a_slot():
    o = oracle.Oracle(W, H)
    u = flat_scene(o, 0)
    gi = o.image("GradientInputs_Prev", 0)
    gi[...] = np.float16(0.1)
    gi[4::3, 4::3] = np.float16(5.0)                    # slot (1,1) = 4 of every stratum is brightest
    o.image("FwdLum", 0)[...] = np.float16(0.1)
    o.image("FwdLum", 0)[4::3, 4::3] = np.float16(2.5)  # it got darker by half
    o.run_pass(layout.PASS_GRADIENT_SAMPLES, 0, u)
    gp = o.image("GradientPrevPix", 0)
    assert np.all(gp[1:-1, 1:-1] == 4)
    g = o.image("DISPingGradient", 0)[1:-1, 1:-1]
    # alpha_D = alpha_S = |2.5-5|/5 * 0.5 = 0.25, mixed 50/50 with a zero history -> 0.125 -> unorm8 32
    # alpha_I = max(0, 5-2.5)/5 * (1 - smoothstep(1,10,0)) * 0.2 = 0.1 -> 0.05 -> unorm8 13
    assert np.all(g[..., 0] == 32) and np.all(g[..., 2] == 32) and np.all(g[..., 1] == 13)
    o.run_pass(layout.PASS_GRADIENT_SAMPLES, 0, u)      # slot 4 was used last frame: must pick another one (:75-78)
    assert np.all(o.image("GradientPrevPix", 0)[1:-1, 1:-1] != 4)
    u.gradientMultDiffuse = u.gradientMultIndirect = u.gradientMultSpecular = 0.0
    o.run_pass(layout.PASS_GRADIENT_SAMPLES, 0, u)
    assert np.all(o.image("GradientPrevPix", 0) == 10) and np.all(o.image("DISPingGradient", 0) == 0)
    o.close()


def test_variance_switches_to_temporal_after_four_frames():
    o = oracle.Oracle(W, H)
    u = flat_scene(o, 0)
    o.image("DiffAccumColor", 0)[...] = e5(1.0, 1.0, 1.0)
    o.image("DiffAccumMoments", 0)[...] = np.array([1.0, 3.0], dtype=np.float16)          # variance 2
    o.image("AccumHistoryLength", 0)[...] = np.array([8, 8, 8, 1], dtype=np.float16)
    o.run_pass(layout.PASS_VARIANCE, 0, u)
    v = o.image("DiffPingColorAndVariance", 0).astype(np.float32)
    assert np.allclose(v[..., 3], 2.0 / 4.0)                                             # /(histLen-4), :153-161
    o.image("AccumHistoryLength", 0)[...] = np.array([2, 2, 2, 1], dtype=np.float16)
    o.run_pass(layout.PASS_VARIANCE, 0, u)
    v = o.image("DiffPingColorAndVariance", 0).astype(np.float32)
    # spatial path, interior: 9 equal samples + temporal moments added to the centre (:165-166):
    # mu1 = (9*1 + 1)/9, mu2 = (9*1 + 3)/9, boost = 1 + 2*(1 - 2/4) = 2
    mu1, mu2 = 10.0 / 9.0, 12.0 / 9.0
    assert np.allclose(v[5:-5, 5:W // 2 - 5, 3], max(0.0, mu2 - mu1 * mu1) * 2.0, rtol=2e-3)
    o.close()


def test_sky_pixels_pass_tests_and_compose_albedo():
    o = oracle.Oracle(W, H)
    u = flat_scene(o, 0)
    o.image("IsSky", 0)[...] = 1
    o.image("DepthWorld", 0)[...] = np.float16(20000.0)
    o.image("DepthWorld_Prev", 0)[...] = np.float16(20000.0)
    o.image("Normal", 0)[...] = 0x80000000
    o.image("Normal_Prev", 0)[...] = 0x80000000
    o.gradient_samples(0, u); o.denoise(0, u)
    assert np.all(o.image("DebugMask", 0) & 0x10)        # z = 20000 in both frames is "consistent" (SURVEY a-3)
    assert np.all(o.image("GradientPrevPix", 0) == 10)   # sky strata are rejected (:98-105)
    pf = layout.decode_b10g11r11(o.image("PreFinal", 0))
    assert np.allclose(pf, 0.5)                          # albedo * throughput (:326-329)
    assert np.all(o.image("AtrousFilteredVariance", 0) == 0)   # never written for invalid depth (stale zeros)
    o.close()


def test_five_iterations_extension_runs_and_differs_only_by_extra_blur():
    a, b = oracle.Oracle(W, H, 4), oracle.Oracle(W, H, 5)
    for frame in range(2):
        g = synth.generate(W, H, frame)
        u = synth.uniform(W, H, frame)
        for o in (a, b):
            for k, v in g.items():
                o.image(k, frame)[...] = v
            o.gradient_samples(frame, u); o.denoise(frame, u)
    assert np.array_equal(a.image("DiffColorHistory", 1), b.image("DiffColorHistory", 1))   # iterations 0..2 identical
    pa, pb = layout.decode_b10g11r11(a.image("PreFinal", 1)), layout.decode_b10g11r11(b.image("PreFinal", 1))
    assert np.isfinite(pb).all() and abs(pa.mean() - pb.mean()) < 0.05 * pa.mean()
    a.close(); b.close()
