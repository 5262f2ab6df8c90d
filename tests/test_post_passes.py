"""The two dispatches that consume the denoiser's outputs: checkerboard resolve
(CmCheckerboard.comp, PreFinal -> Final) and exposure (CmLuminanceHistogram.comp + CmLuminanceAvg.comp).

CPU part: analytic known-answer tests of the oracle.  GPU part: bit-exact parity of the CUDA path (C ABI) against it."""
import ctypes
import math

import numpy as np
import pytest

import oracle as oracle_mod
import synth
from raytracedgl1_b200 import layout


def _pack111110(r, g, b):
    return oracle_mod.lib().asvgfOraclePackUfloat111110(r, g, b)


def _bin(lum, lo, hi):
    """getColorBin restated with numpy fp32 (CmLuminanceHistogram.comp:38-55)."""
    lum = np.float32(lum)
    if lum < np.float32(0.001):
        return 0
    inv = np.float32(1.0) / (np.float32(hi) - np.float32(lo))
    t = (np.float32(math.log2(float(lum))) - np.float32(lo)) * inv
    t = min(max(t, np.float32(0.0)), np.float32(1.0))
    return int(np.float32(t * np.float32(254.0) + np.float32(1.0)))


def test_oracle_histogram_known_answers():
    W, H = 64, 36
    o = oracle_mod.Oracle(W, H)
    try:
        u = synth.uniform(W, H, 0)
        img = o.image("HistogramInput", 0)
        img[...] = 0
        img[0, :] = np.float16(0.125)     # exactly 2^minLog -> t = 0 -> bin 1
        img[1, :] = np.float16(128.0)     # exactly 2^maxLog -> t = 1 -> bin 255
        img[2, :] = np.float16(65504.0)   # above the range -> clamped to bin 255
        img[3, :] = np.float16(0.0005)    # below LUMINANCE_EPS -> bin 0
        img[4, :] = np.float16(4.0)       # log2 = 2 -> t = 0.5 -> uint(0.5*254 + 1) = 128
        img[5, :] = np.float16(np.inf)    # fp16 overflow of the composed luminance -> bin 255
        o.luminance_histogram(0, u)
        h = np.array(o.tonemapping.histogram)
        assert h.sum() == W * H
        assert h[1] == W and h[255] == 3 * W and h[128] == W and h[0] == W * H - 5 * W
        # every fp16 value against the numpy restatement
        vals = np.arange(0, 0x7c01, 37, dtype=np.uint16).view(np.float16)
        o.reset_history()
        img = o.image("HistogramInput", 0)
        flat = img.reshape(-1)
        flat[:] = 0
        flat[:len(vals)] = vals[:flat.size]
        n = min(len(vals), flat.size)
        o.luminance_histogram(0, u)
        want = np.zeros(256, dtype=np.int64)
        for v in flat:
            want[_bin(v, u.minLogLuminance, u.maxLogLuminance)] += 1
        assert n > 800 and np.array_equal(np.array(o.tonemapping.histogram), want)
    finally:
        o.close()


def test_oracle_average_known_answers():
    W, H = 64, 36
    o = oracle_mod.Oracle(W, H)
    try:
        u = synth.uniform(W, H, 0)
        o.image("HistogramInput", 0)[...] = np.float16(4.0)  # every pixel in bin 128
        o.tonemapping.avgLuminance = 1.0
        o.calculate_exposure(0, u)
        tm = o.tonemapping
        assert not any(tm.histogram)  # cleared for the next frame (CmLuminanceAvg.comp:48)
        # weightedLogAverage = 128 - 1 = 127 -> 2^(127/254*10 - 3) = 4; one step of the 1.1/s adaptation from 1
        t = 1.0 - math.exp(-(1.0 / 60.0) * 1.1)
        assert abs(tm.avgLuminance - (1.0 + (4.0 - 1.0) * t)) < 1e-6
        # converges to the scene luminance and never exceeds 2^maxLog
        for _ in range(2000):
            o.calculate_exposure(0, u)
        assert abs(o.tonemapping.avgLuminance - 4.0) < 1e-3
        u.stopEyeAdaptation = 1
        before = o.tonemapping.avgLuminance
        o.image("HistogramInput", 0)[...] = np.float16(100.0)
        o.calculate_exposure(0, u)
        assert o.tonemapping.avgLuminance == before and not any(o.tonemapping.histogram)
        # all-black frame: pixelCount - blackPixelCount clamps to 1, weighted average -1 -> below the range
        u.stopEyeAdaptation = 0
        o.image("HistogramInput", 0)[...] = 0
        o.tonemapping.avgLuminance = 0.0
        o.calculate_exposure(0, u)
        t1 = 2.0 ** ((-1.0 / 254.0) * 10.0 - 3.0)
        assert abs(o.tonemapping.avgLuminance - t1 * t) < 1e-6
    finally:
        o.close()


def test_oracle_resolve_known_answers():
    W, H = 16, 8
    o = oracle_mod.Oracle(W, H)
    try:
        u = synth.uniform(W, H, 0)
        pf = o.image("PreFinal", 0)
        tp = o.image("Throughput", 0)
        rng = np.random.default_rng(3)
        vals = rng.integers(0, 64, size=(H, W, 3)).astype(np.float32)  # small integers: exactly representable in uf11/uf10
        for y in range(H):
            for x in range(W):
                pf[y, x] = _pack111110(*vals[y, x])
        tp[...] = 0
        o.resolve_checkerboard(0, u)
        assert np.array_equal(o.image("Final", 0), pf)  # no split pixel: pass-through
        tp[..., 3] = np.float16(1.0)                   # every pixel split
        o.resolve_checkerboard(0, u)
        fin = layout.decode_b10g11r11(o.image("Final", 0))
        out = np.zeros(4, dtype=np.int32)
        sep = W // 2
        for (x, y) in ((5, 3), (0, 0), (W - 1, H - 1), (1, 2)):
            odd = (x + y % 2) % 2
            cbx = odd * sep + x // 2
            fl = 0.0 if cbx in (0, sep) else 1.0
            fr = 0.0 if cbx in (sep, W) else 1.0
            g = lambda xx, yy: vals[yy, xx] if 0 <= xx < W and 0 <= yy < H else np.zeros(3, np.float32)
            cross = 0.25 * g(x - 1, y) * fl + 0.25 * g(x + 1, y) * fr + 0.25 * g(x, y + 1) + 0.25 * g(x, y - 1)
            want = 0.5 * vals[y, x] + 0.5 * cross
            # uf11 (6-bit mantissa) / uf10 (5-bit) quantisation of the stored result
            assert np.all(np.abs(fin[y, x] - want) <= np.maximum(want, 1e-3) * np.array([2.0 ** -6, 2.0 ** -6, 2.0 ** -5])), (x, y)
        del out
    finally:
        o.close()


def test_oracle_scatter_accumulation_known_answers():
    """CmScatterAccum.comp on a static scene (zero motion): every tap is consistent, the history length counts
    1, 2, 3, ... up to the push constant and the accumulated value is the running mean of the per-frame samples;
    SIMPLE mode writes ambient * (1 - exp(-depth * density)) and the transmittance; sky pixels get (0,0,0,1)."""
    W, H = 32, 16
    o = oracle_mod.Oracle(W, H)
    try:
        u = synth.uniform(W, H, 0)
        u.volumeEnableType = 2  # VOLUMETRIC: the sample comes from ScatterCurrent
        for fi in (0, 1):
            o.image("Motion", fi)[...] = 0
            o.image("DepthWorld", fi)[...] = np.float16(10.0)
            o.image("Normal", fi)[...] = oracle_mod.lib().asvgfOracleEncodeNormal(0.0, 0.0, 1.0)
        samples = [1.0, 3.0, 5.0, 7.0, 9.0, 11.0]
        mean = 0.0
        for f, v in enumerate(samples):
            fi = f & 1
            o.image("ScatterCurrent", fi)[...] = np.float16(v)
            o.scatter_accumulate(fi, u, 4.0)
            n = min(f + 1, 4)
            mean = mean + (v - mean) / n          # mix(accum, cur, 1/historyLen)
            assert np.all(o.image("ScatteringHistory", fi) == np.float16(n)), f
            got = o.image("Scattering", fi).astype(np.float32)
            assert np.allclose(got, mean, rtol=2e-3), (f, got[0, 0], mean)
            mean = float(np.float16(mean))        # the state is stored as fp16
        # disocclusion: the depth jumps, no tap is consistent, history restarts at 1 with the current sample
        o.image("DepthWorld", 0)[...] = np.float16(100.0)
        o.image("ScatterCurrent", 0)[...] = np.float16(2.0)
        o.scatter_accumulate(0, u, 4.0)
        assert np.all(o.image("ScatteringHistory", 0) == np.float16(1.0)) and np.all(o.image("Scattering", 0) == np.float16(2.0))
        # SIMPLE mode on fresh history
        o.reset_history()
        u.volumeEnableType = 1
        u.volumeScattering = 2000.0
        u.volumeAmbient[0], u.volumeAmbient[1], u.volumeAmbient[2] = 0.5, 0.25, 1.0
        o.image("DepthWorld", 0)[...] = np.float16(10.0)
        o.image("DepthWorld", 0)[0, :] = np.float16(20000.0)  # sky
        o.scatter_accumulate(0, u, 8.0)
        f = math.exp(-10.0 * 2000.0 * 0.00005)
        got = o.image("Scattering", 0).astype(np.float32)
        # row 0 of the checkerboard image is row 0 of the regular image
        assert np.allclose(got[1:, :, 3], f, rtol=2e-3) and np.allclose(got[1:, :, 0], 0.5 * (1 - f), rtol=2e-3)
        assert np.all(got[0, :, 3] == 1.0) and np.all(got[0, :, :3] == 0.0)
        u.volumeEnableType = 0
        o.reset_history()
        o.scatter_accumulate(0, u, 8.0)
        got = o.image("Scattering", 0).astype(np.float32)
        assert np.all(got[..., 3] == 1.0) and np.all(got[..., :3] == 0.0)
    finally:
        o.close()


def _lpm_blocks():
    """The LPM control blocks of the reference's host-side setup, as stored in the reference-generated fixture."""
    import os
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asvgf_ref_96x54.npz"))
    return ref["lpm_default"], ref["lpm_tinted"]


def test_oracle_prepare_final_known_answers():
    """CmPrepareFinal.comp on hand-computed pixels.  Classic-shading columns: Uncharted 2 of 2 * hdr over its white-point value
    (:75-93); LPM columns: a grey pixel keeps its hue and lands on the tone curve luma^c / (luma^c * a + b) of the control block
    (ffx_lpm.h:759-761 with ratio = 1); bloom input = (hdr + emission * 2^(EV+m-3)) * exposure * smoothstep (:123-140);
    exposure = 1 / (1.2 * 2^EV100), EV100 = log2(avg * 100 / 12.5) (Exposure.h:36-58)."""
    W, H = 16, 4
    lpm, _ = _lpm_blocks()
    o = oracle_mod.Oracle(W, H)
    try:
        u = synth.uniform(W, H, 0)
        u.luminanceWhitePoint, u.bloomThreshold, u.bloomEmissionMultiplier, u.emissionMaxScreenColor = 10.0, 4.0, 16.0, 1.5
        u.lightmapScreenCoverage = 0.5             # columns 0..7 classic, 8..15 ray traced
        o.tonemapping.avgLuminance = 0.125         # EV100 = log2(1) = 0 -> exposure = 1 / 1.2
        fin, sc = o.image("Final", 0), o.image("Scattering", 0)
        grey = [0.0, 0.5, 1.0, 2.0, 8.0, 0.25, 4.0, 16.0]
        for x in range(W):
            fin[:, x] = _pack111110(grey[x % 8], grey[x % 8], grey[x % 8])
        sc[...] = np.array([0, 0, 0, 1], np.float16)   # no in-scattering, full transmittance
        o.image("ScreenEmission", 0)[...] = 0
        o.image("AcidFog", 0)[...] = 0
        o.prepare_final(0, u, lpm)
        out = layout.decode_b10g11r11(o.image("Final", 0))
        bloom = layout.decode_b10g11r11(o.image("BloomInput", 0))
        f32 = np.float32

        def unch(x):
            A, B, C, D, E, F = map(f32, (0.15, 0.50, 0.10, 0.20, 0.02, 0.30))
            return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F
        exposure = f32(1.0) / (f32(1.2) * f32(1.0))
        k = lpm.view(np.float32)
        contrast, tsb_x, tsb_y = k[3], k[4], k[5]
        for x in range(W):
            v = f32(grey[x % 8])
            if x < 8:
                want = unch(f32(2.0) * v) / unch(f32(10.0))
            else:
                luma = np.power(v * exposure * f32(k[6] + k[7] + k[8]), contrast) if v > 0 else f32(0)   # lumaT sums to ~1 on a grey pixel
                want = min(max(luma / (luma * tsb_x + tsb_y), f32(0)), f32(1))
            if v == 0 and x >= 8:
                continue    # black: 0 * (1/0) is NaN in the shader; whatever the pack makes of it is covered by the reference-pinned tests
            assert abs(out[0, x, 1] - want) <= max(float(want), 1e-3) * 2.0 ** -5, (x, out[0, x], want)
            assert abs(out[0, x, 0] - out[0, x, 1]) <= max(float(want), 1e-3) * 2.0 ** -5   # grey stays grey
            b = v * exposure                     # no emission
            t = min(max(b / f32(4.0), f32(0)), f32(1))
            wb = b * (t * t * (f32(3) - f32(2) * t))
            assert abs(bloom[0, x, 1] - wb) <= max(float(wb), 1e-3) * 2.0 ** -5, (x, bloom[0, x], wb)
        # avgLuminance = 0: EV100 = -inf, the luminous exposure is defined as 0 (Exposure.h:50-54) -> black frame, no NaN
        o.tonemapping.avgLuminance = 0.0
        for x in range(W):
            fin[:, x] = _pack111110(1.0, 2.0, 3.0)
        u.lightmapScreenCoverage = 0.0
        o.prepare_final(0, u, lpm)
        assert not o.image("BloomInput", 0).any()
    finally:
        o.close()


def test_uniform_and_tonemapping_layout():
    assert layout.GlobalUniform.volumeEnableType.offset == 832 and layout.GlobalUniform.volumeScattering.offset == 836
    assert layout.GlobalUniform.volumeAmbient.offset == 992
    assert (layout.INDEX["Scattering"], layout.INDEX["ScatteringHistory_Prev"]) == (48, 51)
    assert layout.GlobalUniform.minLogLuminance.offset == 668
    assert layout.GlobalUniform.maxLogLuminance.offset == 672
    assert layout.GlobalUniform.stopEyeAdaptation.offset == 680
    assert ctypes.sizeof(layout.Tonemapping) == 1028 and layout.Tonemapping.avgLuminance.offset == 1024
    assert layout.INDEX["Final"] == 22


# ---------------------------------------------------------------- GPU parity ----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("size", [(320, 180), (322, 182), (1280, 720)])
def test_post_passes_parity(size):
    """After a denoised frame: Final bit-identical, histogram bit-identical, avgLuminance identical to 1e-6 relative
    (its exp2/exp run in double on both sides), over several frames so that the adaptation state carries."""
    from _harness import Pair
    W, H = size
    p = Pair(W, H)
    try:
        rng = np.random.default_rng(11)
        for frame in range(3):
            def mutate(g):
                # mark a third of the pixels as split (reflect/refract), RaygenPrimary writes 1 there
                g["Throughput"][..., 3] = (rng.random((H, W)) < 0.33).astype(np.float16)
            fi, u = p.run_frame(frame, mutate=mutate)
            p.sync_state_to_cuda()  # both sides continue from the oracle's PreFinal / HistogramInput
            p.oracle.resolve_checkerboard(fi, u)
            p.cuda.resolve_checkerboard(fi, u)
            p.cuda.synchronize()
            assert np.array_equal(p.oracle.image("Final", fi), p.cuda.download("Final", fi)), "Final frame %d" % frame
            p.oracle.luminance_histogram(fi, u)
            p.cuda.luminance_histogram(fi, u)
            got = p.cuda.tonemapping()
            want = p.oracle.tonemapping
            assert sum(want.histogram) == W * H
            assert list(got.histogram) == list(want.histogram), "histogram frame %d" % frame
            p.oracle.luminance_average(u)
            p.cuda.luminance_average(fi, u)
            got = p.cuda.tonemapping()
            want = p.oracle.tonemapping
            assert not any(got.histogram)
            assert abs(got.avgLuminance - want.avgLuminance) <= 1e-6 * abs(want.avgLuminance), (got.avgLuminance, want.avgLuminance)
            assert want.avgLuminance > 0
    finally:
        p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [2, 1, 0])
def test_scatter_accumulation_parity(mode):
    """CmScatterAccum.comp over a moving sequence (camera motion, disocclusions): Scattering and ScatteringHistory
    bit-identical to the oracle every frame, state carried on both sides independently."""
    from _harness import Pair
    W, H = 322, 182
    p = Pair(W, H)
    try:
        rng = np.random.default_rng(5)
        for frame in range(5):
            fi = frame & 1
            u = p.feed_inputs(frame, fi)
            u.volumeEnableType = mode
            u.volumeScattering = 1500.0
            u.volumeAmbient[0], u.volumeAmbient[1], u.volumeAmbient[2] = 0.7, 0.8, 1.0
            cur = rng.random((H, W, 4)).astype(np.float16)
            p.oracle.image("ScatterCurrent", fi)[...] = cur
            p.cuda.upload("ScatterCurrent", fi, cur)
            p.oracle.scatter_accumulate(fi, u, 16.0)
            p.cuda.scatter_accumulate(fi, u, 16.0)
            p.cuda.synchronize()
            for name in ("Scattering", "ScatteringHistory"):
                want, got = p.oracle.image(name, fi), p.cuda.download(name, fi)
                assert np.array_equal(want.view(np.uint16), got.view(np.uint16)), (mode, frame, name, int((want.view(np.uint16) != got.view(np.uint16)).sum()))
            if frame >= 2:
                assert p.oracle.image("ScatteringHistory", fi).max() >= 3.0  # history does accumulate
    finally:
        p.close()


@pytest.mark.gpu
def test_histogram_every_half_pattern_and_flat_image():
    """Every fp16 bit pattern as input (negatives, denormals, inf, NaN) and a flat image (all lanes in one bin: the
    warp-aggregated atomic path) give the oracle's histogram exactly."""
    from _harness import Pair
    W, H = 512, 128  # 65536 texels
    p = Pair(W, H)
    try:
        u = synth.uniform(W, H, 0)
        allbits = np.arange(65536, dtype=np.uint16).view(np.float16).reshape(H, W)
        for img in (allbits, np.full((H, W), 3.0, dtype=np.float16)):
            p.oracle.reset_history()
            p.oracle.image("HistogramInput", 0)[...] = img
            p.cuda.upload("HistogramInput", 0, np.ascontiguousarray(img))
            p.cuda.set_tonemapping(layout.Tonemapping())
            p.oracle.luminance_histogram(0, u)
            p.cuda.luminance_histogram(0, u)
            assert list(p.cuda.tonemapping().histogram) == list(p.oracle.tonemapping.histogram)
    finally:
        p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["default", "tinted_classic_half", "debug_motion", "debug_gradients", "debug_unfiltered_specular", "debug_normals"])
def test_prepare_final_parity(case):
    """asvgfPrepareFinal (CmPrepareFinal.comp) against the oracle from identical state, after a full frame + scattering + exposure +
    resolve.  pow() on the device is ex2(y lg2 x) on the MUFU pipe, the oracle's is libm's: the 6/6/5-bit outputs agree within one code, >= 99.95 %
    of the texels bit for bit (gate 99.95 %, achieved 99.999 %); the debug views involve no transcendental and are bit-exact."""
    from _harness import Pair, compare_image
    from raytracedgl1_b200 import denoiser as dn
    W, H = 642, 362
    lpm_default, lpm_tinted = _lpm_blocks()
    p = Pair(W, H)
    try:
        rng = np.random.default_rng(5)
        for frame in range(3):
            def mutate(g):
                g["Throughput"][..., 3] = (rng.random((H, W)) < 0.33).astype(np.float16)
                finite = np.uint32(~((1 << 10) | (1 << 21) | (1 << 31)) & 0xFFFFFFFF) & np.uint32(0x3BDEF7BF | 0x3FF)   # modest magnitudes
                for name in ("ScreenEmisRT", "AcidFogRT"):
                    g[name] = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) & finite
            fi, u = p.run_frame(frame, mutate=mutate)
            u.volumeEnableType, u.volumeScattering = 1, 1500.0
            u.luminanceWhitePoint, u.bloomThreshold, u.bloomEmissionMultiplier, u.emissionMaxScreenColor = 10.0, 4.0, 16.0, 1.5
            lpm = lpm_default
            if case == "tinted_classic_half":
                u.lightmapScreenCoverage, lpm = 0.5, lpm_tinted
            flag = {"debug_motion": 0, "debug_gradients": 1, "debug_unfiltered_specular": 3, "debug_normals": 10}.get(case)
            p.sync_state_to_cuda()
            for side in (p.oracle, p.cuda):
                side.scatter_accumulate(fi, u, 16.0)
                side.resolve_checkerboard(fi, u)
                side.luminance_histogram(fi, u)
            p.oracle.luminance_average(u)
            p.cuda.luminance_average(fi, u)
            p.cuda.set_tonemapping(p.oracle.tonemapping)      # identical exposure state on both sides
            if flag is not None:
                u.debugShowFlags = 1 << flag
            p.oracle.prepare_final(fi, u, lpm)
            p.cuda.prepare_final(fi, u, lpm)
            p.cuda.synchronize()
            for name in ("Final", "BloomInput"):
                ref, got = p.oracle.image(name, fi), p.cuda.download(name, fi)
                r = compare_image(name, ref, got)
                assert r["ok"] and r["exact_frac"] >= 0.9995, (case, frame, name, r)   # achieved: 0.99999
                if flag is not None and name == "Final":
                    assert np.array_equal(ref, got), (case, frame)
        u.debugShowFlags = 1 << 8
        with pytest.raises(dn.AsvgfError) as e:
            p.cuda.prepare_final(fi, u, lpm_default)
        assert e.value.code == dn.UNSUPPORTED
    finally:
        p.close()
