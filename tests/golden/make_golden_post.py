#!/usr/bin/env python
"""Regenerates tests/golden/asvgf_post_64x36.npz: outputs of the CPU oracle for the dispatches after the denoiser
(checkerboard resolve, exposure histogram + eye adaptation, scattering accumulation) over
three frames of the seeded synthetic sequence at 64x36.  Like asvgf_64x36.npz these freeze the ORACLE's behaviour
(regression on the CPU); they do not pin it to the reference, whose shaders cannot run in this image.

    python tests/golden/make_golden_post.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402

W, H, N, FRAMES = 64, 36, 4, 3


def run():
    oracle.build(); synth.build()
    o = oracle.Oracle(W, H, N)
    out = {}
    rng = np.random.default_rng(2024)
    for f in range(FRAMES):
        fi = f & 1
        g = synth.generate(W, H, f)
        g["Throughput"][..., 3] = (rng.random((H, W)) < 0.4).astype(np.float16)   # split (reflect/refract) pixels
        for k, v in g.items():
            o.image(k, fi)[...] = v
        u = synth.uniform(W, H, f)
        u.volumeEnableType, u.volumeScattering = 1, 1500.0                           # VOLUME_ENABLE_SIMPLE
        u.volumeAmbient[0], u.volumeAmbient[1], u.volumeAmbient[2] = 0.7, 0.8, 1.0
        o.gradient_samples(fi, u)
        o.denoise(fi, u)
        o.scatter_accumulate(fi, u, 8.0)
        o.luminance_histogram(fi, u)
        out["f%d_histogram" % f] = np.array(o.tonemapping.histogram, dtype=np.uint32)
        o.luminance_average(u)
        out["f%d_avgLuminance" % f] = np.array([o.tonemapping.avgLuminance], dtype=np.float32)
        o.resolve_checkerboard(fi, u)
        for n in ("Final", "Scattering", "ScatteringHistory"):
            out["f%d_%s" % (f, n)] = np.array(o.image(n, fi), copy=True)
    o.close()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "asvgf_post_64x36.npz")
    np.savez_compressed(path, **run())
    print("wrote", path, os.path.getsize(path), "bytes")
