"""Known-answer tests that pin the oracle's primitives (the reference itself ships no test vectors for them).

E5B9G9R9: the Khronos EXT_texture_shared_exponent algorithm that Utils.h:150 cites; normal codec Utils.h:108-140;
checkerboard maps ShaderCommonGLSLFunc.h:321-342; getPrevScreenPos :271-277.
"""
import ctypes

import numpy as np
import pytest

import oracle
from raytracedgl1_b200 import layout

L = oracle.lib


def enc(r, g, b):
    return L().asvgfOracleEncodeE5B9G9R9(r, g, b)


def dec(p):
    out = (ctypes.c_float * 3)()
    L().asvgfOracleDecodeE5B9G9R9(p, out)
    return tuple(out)


def test_e5_known_answers():
    assert enc(0, 0, 0) == 0
    # 1.0 = 256 * 2^(16-15-9): exponent 16, mantissa 256
    assert enc(1.0, 0.0, 0.0) == (16 << 27) | 256
    assert enc(1.0, 1.0, 1.0) == (16 << 27) | (256 << 18) | (256 << 9) | 256
    assert enc(0.5, 0.25, 0.125) == (15 << 27) | (64 << 18) | (128 << 9) | 256
    # clamp to sharedexp_max = 65408 = 511 * 2^7, negatives and NaN to 0
    assert enc(1e9, -3.0, float("nan")) == (31 << 27) | 511
    assert dec(enc(1e9, 0, 0))[0] == 65408.0
    # mantissa that rounds up to 512 bumps the exponent (Utils.h:178-183)
    x = np.nextafter(np.float32(2.0), np.float32(0.0))
    assert enc(float(x), 0, 0) == (17 << 27) | 256
    # values below 2^-16 share the minimum exponent 0: unit 2^-24
    assert enc(2.0 ** -24, 0, 0) == 1
    assert enc(2.0 ** -26, 0, 0) == 0          # rounds to zero mantissa
    assert dec(1) == (2.0 ** -24, 0.0, 0.0)


def test_e5_round_trip_random():
    rng = np.random.default_rng(1)
    v = np.exp(rng.uniform(-12, 11, size=(2000, 3))).astype(np.float32)
    for r, g, b in v:
        p = enc(float(r), float(g), float(b))
        d = np.array(dec(p), dtype=np.float64)
        m = max(min(float(max(r, g, b)), 65408.0), 2.0 ** -16)
        step = 2.0 ** (np.floor(np.log2(m)) + 1 - 9)
        assert np.all(np.abs(d - np.minimum([r, g, b], 65408.0)) <= step * 0.5 * 1.0001 + 1e-30)
        assert enc(*[float(x) for x in d]) == p          # canonical: decode -> encode is the identity
        assert np.allclose(layout.decode_e5(np.array([p], dtype=np.uint32))[0], d)


def test_every_e5_component_is_exact_in_fp16():
    """The a-trous kernel keeps specular / indirect colours of the staged tile as fp16 (asvgf_atrous_half.cuh) and the anti-firefly
    kernel runs its min / max on fp16 planes: both rely on every decoded E5B9G9R9 component m * 2^(e-24) (m < 512, e < 32;
    Utils.h:194-207) surviving f32 -> f16 -> f32 unchanged, subnormal halves included."""
    m = np.arange(512, dtype=np.float64)
    for e in range(32):
        v = (m * 2.0 ** (e - 24)).astype(np.float32)
        assert np.array_equal(v.astype(np.float16).astype(np.float32), v), e
    # ... and so do the tile's depth / roughness words: an r16f depth and a byte 0..255
    assert np.array_equal(np.arange(256, dtype=np.float32).astype(np.float16).astype(np.float32), np.arange(256, dtype=np.float32))


def test_half_conversions_match_ieee():
    allh = np.arange(65536, dtype=np.uint16)
    ref = allh.view(np.float16).astype(np.float32)
    got = np.array([L().asvgfOracleHalfToFloat(int(h)) for h in allh], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)) or np.array_equal(np.isnan(got), np.isnan(ref))
    fin = ~np.isnan(ref)
    assert np.array_equal(got[fin], ref[fin])
    rng = np.random.default_rng(2)
    f = np.concatenate([
        np.exp(rng.uniform(-30, 13, 20000)).astype(np.float32) * rng.choice([-1, 1], 20000).astype(np.float32),
        np.array([0, -0.0, 65504, 65519.99, 65520, 1e9, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 2.0 ** -14,
                  np.inf, -np.inf, 20000.0, 1 + 2.0 ** -11, 1 + 3 * 2.0 ** -11], dtype=np.float32)])
    with np.errstate(over="ignore"):
        want = f.astype(np.float16).view(np.uint16)
    have = np.array([L().asvgfOracleFloatToHalf(float(x)) for x in f], dtype=np.uint16)
    assert np.array_equal(want, have)


def test_ufloat_11_11_10():
    pack, unpack = L().asvgfOraclePackUfloat111110, L().asvgfOracleUnpackUfloat111110
    out = (ctypes.c_float * 3)()
    assert pack(0, 0, 0) == 0 and pack(-1, float("nan"), -0.0) == 0
    assert pack(1.0, 1.0, 1.0) == (15 << 6) | ((15 << 6) << 11) | ((15 << 5) << 22)
    unpack(pack(1.0, 0.5, 2.0), out)
    assert tuple(out) == (1.0, 0.5, 2.0)
    unpack(pack(1e9, 1e9, 1e9), out)                      # overflow clamps to max finite
    assert tuple(out) == (65024.0, 65024.0, 64512.0)
    # RTNE on a 6-bit mantissa: 1 + 1/128 is a tie between 1 and 1+1/64 -> even (1.0); 1 + 3/128 -> 1 + 2/64
    unpack(pack(1 + 1 / 128, 1 + 3 / 128, 1 + 1 / 64 + 1 / 64), out)
    assert out[0] == 1.0 and out[1] == 1 + 2 / 64 and out[2] == 1 + 1 / 32
    rng = np.random.default_rng(3)
    v = np.exp(rng.uniform(-9, 10, size=(500, 3))).astype(np.float32)
    p = np.array([pack(float(a), float(b), float(c)) for a, b, c in v], dtype=np.uint32)
    d = layout.decode_b10g11r11(p)
    assert np.all(np.abs(d - v) <= v * np.array([2.0 ** -7, 2.0 ** -7, 2.0 ** -6]) * 1.0001)


def test_unorm8():
    f = L().asvgfOracleFloatToUnorm8
    assert f(0.0) == 0 and f(1.0) == 255 and f(2.0) == 255 and f(-1.0) == 0 and f(float("nan")) == 0
    assert f(0.5) == 128 and f(0.5 / 255.0) == 0 and f(1.5 / 255.0) == 2   # ties to even


def test_normal_codec():
    encn, decn = L().asvgfOracleEncodeNormal, L().asvgfOracleDecodeNormal
    out = (ctypes.c_float * 3)()
    decn(0, out)
    assert tuple(out) == (0.0, 0.0, 1.0)                                  # an all-zero (out-of-bounds) texel
    decn(0xFFFF0000, out)
    assert abs(out[2] + 1.0) < 1e-7 and abs(out[0]) < 1e-6               # south pole
    assert encn(0.0, 0.0, 0.0) == 0x80000000                              # the sky normal, RaygenPrimary.inl:119
    rng = np.random.default_rng(4)
    n = rng.normal(size=(2000, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    for x, y, z in n.astype(np.float32):
        decn(encn(float(x), float(y), float(z)), out)
        assert np.linalg.norm(np.array(out) - [x, y, z]) < 1.5e-4        # 16-bit angles
    # theta seam: k wraps modulo 65536
    assert encn(1.0, -1e-7, 0.0) & 0xFFFF in (0, 65535)


def test_checkerboard_maps_are_inverse():
    cb, rg = L().asvgfOracleCheckerboardPix, L().asvgfOracleRegularPix
    o = (ctypes.c_int * 2)()
    W = 16
    seen = set()
    for y in range(6):
        for x in range(W):
            cb(x, y, W, o)
            cx, cy = o[0], o[1]
            assert cy == y and (cx >= W // 2) == bool((x + y % 2) % 2)
            rg(cx, cy, W, o)
            assert (o[0], o[1]) == (x, y)
            seen.add((cx, cy))
    assert len(seen) == 6 * W


def test_prev_screen_pos():
    f = L().asvgfOraclePrevScreenPos
    o = (ctypes.c_float * 2)()
    f(0.0, 0.0, 10, 20, 1280.0, 720.0, o)
    assert abs(o[0] - 10.5) < 1e-4 and abs(o[1] - 20.5) < 1e-4
    f(1.0 / 1280.0, -2.0 / 720.0, 700, 20, 1280.0, 720.0, o)            # motion is in full-screen UV; x is half-res
    assert abs(o[0] - 701.0) < 1e-3 and abs(o[1] - 18.5) < 1e-3


def test_uniform_layout_matches_reference_offsets():
    U = layout.GlobalUniform
    assert ctypes.sizeof(U) == 2096
    want = dict(view=0, invView=64, viewPrev=128, projection=192, invProjection=256, projectionPrev=320,
                lightmapScreenCoverage=644, renderWidth=652, renderHeight=656, frameId=660, timeDelta=664,
                cameraPosition=720, cameraPositionPrev=736, debugShowFlags=752, jitterX=932, jitterY=936,
                gradientMultDiffuse=964, gradientMultIndirect=968, gradientMultSpecular=972, antiFireflyEnabled=988)
    for k, v in want.items():
        assert getattr(U, k).offset == v, k
