#!/usr/bin/env python
"""Regenerates tests/golden/asvgf_64x36.npz: outputs of the CPU oracle for three frames of the seeded synthetic
sequence at 64x36 (4 a-trous iterations).

The reference ships no golden vectors for this path and its shaders cannot be executed in this image,
so these fixtures do NOT pin the oracle to the reference; they freeze the oracle's own behaviour so that an
accidental change of the oracle (or of the synthetic generator) is caught on the CPU, and they give the GPU tests a
fixed target that does not depend on building the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import synth  # noqa: E402

W, H, N, FRAMES = 64, 36, 4, 3
NAMES = ["PreFinal", "HistogramInput", "DebugMask", "GradientPrevPix", "DISPingGradient", "DiffColorHistory",
         "AccumHistoryLength", "SpecAccumColor", "IndirAccum", "AtrousFilteredVariance"]


def run():
    oracle.build(); synth.build()
    o = oracle.Oracle(W, H, N)
    out = {}
    for f in range(FRAMES):
        fi = f & 1
        for k, v in synth.generate(W, H, f).items():
            o.image(k, fi)[...] = v
        u = synth.uniform(W, H, f)
        o.gradient_samples(fi, u)
        o.denoise(fi, u)
        for n in NAMES:
            out["f%d_%s" % (f, n)] = np.array(o.image(n, fi), copy=True)
    o.close()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "asvgf_64x36.npz")
    np.savez_compressed(path, **run())
    print("wrote", path, os.path.getsize(path), "bytes")
