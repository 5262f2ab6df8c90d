"""Golden fixtures (tests/golden/asvgf_64x36.npz, made by tests/golden/make_golden.py from the CPU oracle): the oracle
must still reproduce them bit for bit (CPU), and the CUDA path must match them within the stated tolerance (GPU)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asvgf_64x36.npz")


def test_oracle_reproduces_golden_fixtures():
    ref = np.load(GOLDEN)
    got = make_golden.run()
    assert sorted(ref.files) == sorted(got)
    for k in ref.files:
        assert ref[k].dtype == got[k].dtype and ref[k].shape == got[k].shape, k
        assert np.array_equal(ref[k].view(np.uint8), got[k].view(np.uint8)), k


def test_oracle_reproduces_post_pass_fixtures():
    """asvgf_post_64x36.npz (tests/golden/make_golden_post.py): resolve, exposure and scattering accumulation."""
    import make_golden_post
    ref = np.load(os.path.join(os.path.dirname(GOLDEN), "asvgf_post_64x36.npz"))
    got = make_golden_post.run()
    assert sorted(ref.files) == sorted(got)
    for k in ref.files:
        assert ref[k].dtype == got[k].dtype and ref[k].shape == got[k].shape, k
        assert np.array_equal(ref[k].view(np.uint8), got[k].view(np.uint8)), k
    assert int(ref["f2_histogram"].sum()) == make_golden_post.W * make_golden_post.H and float(ref["f2_avgLuminance"][0]) > 0.0
    assert float(ref["f2_ScatteringHistory"].max()) == 3.0      # three frames of history where nothing was disoccluded


@pytest.mark.gpu
def test_cuda_path_matches_golden_fixtures():
    import synth
    from _harness import compare_image
    from raytracedgl1_b200 import denoiser as dn
    from raytracedgl1_b200 import layout
    ref = np.load(GOLDEN)
    W, H, N = make_golden.W, make_golden.H, make_golden.N
    d = dn.Denoiser(W, H, N, flags=dn.CREATE_EXPORT_DEBUG_MASK)
    try:
        for f in range(make_golden.FRAMES):
            fi = f & 1
            for k, v in synth.generate(W, H, f).items():
                d.upload(k, fi, v)
            u = synth.uniform(W, H, f)
            d.gradient_samples(fi, u)
            d.denoise(fi, u)
            d.synchronize()
            for n in make_golden.NAMES:
                lum = None
                if n == "AtrousFilteredVariance":   # a luminance standard deviation: tolerance relative to luminance
                    lum = layout.luminance(np.abs(ref["f%d_DiffColorHistory" % f][..., :3].astype(np.float32)))
                r = compare_image(n, ref["f%d_%s" % (f, n)], d.download(n, fi), outlier_factor=6.0, outlier_frac=1e-3,
                                  lum_scale=lum)
                assert r["ok"], (f, n, r)
                if r["kind"] == "int":
                    assert r["n_bad"] == 0, (f, n)
    finally:
        d.close()
