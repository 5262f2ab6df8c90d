"""Golden fixtures made by the REFERENCE'S OWN SHADER TEXT (tests/golden/asvgf_ref_96x54.npz, generated in the build
container by tests/golden/make_golden_ref.py through oracle/_ref, see oracle/ref_glsl/build_ref.py):

  * the hand-written oracle must reproduce them bit for bit over the free-running four-frame sequence (CPU; this is
    what pins the oracle to the reference also on the GPU box, where /root/reference does not exist);
  * the CUDA path must match them within the stated tolerance, integer images bit-exact (GPU).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_ref as mg  # noqa: E402


def test_oracle_reproduces_reference_shader_fixtures():
    import oracle
    ref = np.load(mg.PATH)
    o = oracle.Oracle(mg.W, mg.H, mg.N)
    try:
        got = mg.run(o, {k: ref[k] for k in mg.LPM_PARAMS})
    finally:
        o.close()
    assert sorted(ref.files) == sorted(got)
    for k in ref.files:
        assert ref[k].dtype == got[k].dtype and ref[k].shape == got[k].shape, k
        assert np.array_equal(ref[k].view(np.uint8), got[k].view(np.uint8)), k
    last = mg.FRAMES - 1
    assert int(ref["f%d_histogram" % last].sum()) == mg.W * mg.H and float(ref["f%d_avgLuminance" % last][0]) > 0.0
    assert float(ref["f%d_ScatteringHistory" % last].max()) == float(mg.FRAMES)   # history where nothing was disoccluded


@pytest.mark.gpu
def test_cuda_path_matches_reference_shader_fixtures():
    """Every frame from the fixture's own frame-start state (the previous frame's reference outputs)."""
    import synth
    from _harness import compare_image, OUTLIER_FACTOR, OUTLIER_FRAC, branch_flip_allowance
    from raytracedgl1_b200 import denoiser as dn
    from raytracedgl1_b200 import layout
    ref = np.load(mg.PATH)
    d = dn.Denoiser(mg.W, mg.H, mg.N)
    chain = [n for n in mg.NAMES if n not in mg.POST]
    try:
        for f in range(mg.FRAMES):
            fi = f & 1
            if f > 0:
                for n in mg.NAMES:
                    d.upload(n, fi ^ 1, np.ascontiguousarray(ref["f%d_%s" % (f - 1, n)]))
                tm = layout.Tonemapping()
                tm.avgLuminance = float(ref["f%d_avgLuminance" % (f - 1)][0])
                d.set_tonemapping(tm)
            for k, v in mg.inputs(f).items():
                d.upload(k, fi, v)
            u = synth.uniform(mg.W, mg.H, f)
            mg.edit_uniform(u)
            d.gradient_samples(fi, u)
            d.denoise(fi, u)
            d.synchronize()
            for n in chain:
                lum = None
                if n == "AtrousFilteredVariance":   # a luminance standard deviation: tolerance relative to luminance
                    lum = layout.luminance(np.abs(ref["f%d_DiffColorHistory" % f][..., :3].astype(np.float32)))
                r = compare_image(n, ref["f%d_%s" % (f, n)], d.download(n, fi), outlier_factor=OUTLIER_FACTOR,
                                  outlier_frac=OUTLIER_FRAC, lum_scale=lum, bad_allow=branch_flip_allowance(mg.W * mg.H))
                assert r["ok"], (f, n, r)
                if r["kind"] == "int":
                    assert r["n_bad"] == 0, (f, n)      # GradientPrevPix exact; DebugMask bits 0-4 exact (bit 5: _harness.py)
            # the post dispatches are bit-exact by construction: run them on the fixture's own PreFinal / HistogramInput
            for n in ("PreFinal", "HistogramInput"):
                d.upload(n, fi, np.ascontiguousarray(ref["f%d_%s" % (f, n)]))
            d.scatter_accumulate(fi, u, 8.0)
            d.luminance_histogram(fi, u)
            d.synchronize()
            assert list(d.tonemapping().histogram) == list(ref["f%d_histogram" % f])
            d.luminance_average(fi, u)
            d.resolve_checkerboard(fi, u)
            d.synchronize()
            assert np.float32(d.tonemapping().avgLuminance) == ref["f%d_avgLuminance" % f][0]
            for n in mg.POST:
                if n != "BloomInput":
                    assert np.array_equal(ref["f%d_%s" % (f, n)], d.download(n, fi)), (f, n)
            # the final prepare (LPM tonemap: pow() on the device is not libm's): within one code of the 6/6/5-bit mantissas
            d.prepare_final(fi, u, ref["lpm_default" if f % 2 == 0 else "lpm_tinted"])
            d.synchronize()
            for n, key in (("Final", "FinalPrepared"), ("BloomInput", "BloomInput")):
                r = compare_image(n, ref["f%d_%s" % (f, key)], d.download(n, fi))
                assert r["ok"] and r["exact_frac"] >= 0.9995, (f, n, r)
    finally:
        d.close()
