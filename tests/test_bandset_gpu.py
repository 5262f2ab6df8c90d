"""The multi-GPU band mode behind the C ABI (include/asvgf.h, asvgfBandSet*; csrc/asvgf_bandset.inl) on ONE device: a band
set may list a device several times, so two, three and four bands on cuda:0 exercise the whole schedule -- apron passes,
the four push points, the event ordering, unequal bands -- without a second GPU.  The correctness gate is the one of
SURVEY.md 8(e): the banded frame is BIT-IDENTICAL to the single-handle frame, every image, every frame."""
import numpy as np
import pytest

import synth
from raytracedgl1_b200 import layout

pytestmark = pytest.mark.gpu

OUTPUTS = ["PreFinal", "HistogramInput", "DiffColorHistory", "AccumHistoryLength", "DiffAccumMoments", "SpecAccumColor", "IndirAccum",
           "DISGradientHistory", "GradientPrevPix", "AtrousFilteredVariance", "DebugMask", "DiffAccumColor"]


def _run(W, H, devices, boundaries, frames, mutate=None, iterations=4, halo_rows=0):
    from raytracedgl1_b200 import denoiser as dn
    single = dn.Denoiser(W, H, iterations, flags=dn.CREATE_EXPORT_DEBUG_MASK)
    bs = dn.BandSet(W, H, devices, iterations, flags=dn.CREATE_EXPORT_DEBUG_MASK, boundaries=boundaries, halo_rows=halo_rows)
    try:
        for frame in range(frames):
            fi = frame & 1
            g = synth.generate(W, H, frame)
            if mutate:
                mutate(frame, g)
            u = synth.uniform(W, H, frame)
            for name, arr in g.items():
                single.upload(name, fi, arr)
                bs.upload(name, fi, arr)
            single.gradient_samples(fi, u)
            single.denoise(fi, u)
            single.synchronize()
            bs.frame(fi, u)
            bs.synchronize()
            for name in OUTPUTS:
                a, b = single.download(name, fi), bs.download(name, fi)
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), "frame %d: %s differs between %d bands and the single handle" % (
                    frame, name, len(devices))
    finally:
        bs.close()
        single.close()


@pytest.mark.parametrize("bands,boundaries", [(1, None), (2, None), (4, None), (3, [0, 99, 231, 360]), (2, [0, 291, 360])])
def test_band_set_is_bit_identical_to_the_single_handle(bands, boundaries):
    _run(640, 360, [0] * bands, boundaries, frames=5)


def test_band_set_across_devices_is_bit_identical():
    """The same gate with the bands on DIFFERENT GPUs (peer copies over NVLink) when the box has several; on a one-GPU box the
    device list degenerates to [0, 0] (two bands of a 640x720 frame on one device) instead of skipping."""
    import torch
    n = torch.cuda.device_count()
    _run(640, 720, list(range(min(n, 4))) if n >= 2 else [0, 0], None, frames=5)


def test_band_set_five_iterations_and_ragged_size():
    """N = 5 (stride-16 apron: 30 + 2 + 1 + 1 rows) on a frame whose width and height do not divide by 3."""
    _run(322, 362, [0, 0], None, frames=3, iterations=5)


def test_motion_beyond_the_halo_is_reported_not_hidden():
    """Vertical motion of 100 rows per frame across the band seam: the temporal pass of a band would have to read rows the
    band does not store.  asvgfBandSetSynchronize must return ASVGF_HALO_EXCEEDED (the single-GPU frame cannot be
    reproduced), and a halo sized for that motion must make the frames identical again."""
    from raytracedgl1_b200 import denoiser as dn
    W, H = 320, 720

    def fast_motion(frame, g):
        g["Motion"][..., 1] = np.float16(100.0 / H)      # screen-UV units: +100 rows

    with pytest.raises(dn.AsvgfError) as e:
        _run(W, H, [0, 0], None, frames=2, mutate=fast_motion)
    assert e.value.code == dn.HALO_EXCEEDED
    _run(W, H, [0, 0], None, frames=3, mutate=fast_motion, halo_rows=132)   # 18 apron rows + 100 rows of motion, multiple of 3


def test_band_set_argument_errors():
    from raytracedgl1_b200 import denoiser as dn
    with pytest.raises(dn.AsvgfError):
        dn.BandSet(640, 360, [0, 0], boundaries=[0, 100, 360])        # 100 is not a multiple of 3
    with pytest.raises(dn.AsvgfError):
        dn.BandSet(640, 360, [0] * 8)                                  # 45-row bands cannot hold a 66-row halo
    with pytest.raises(dn.AsvgfError):
        dn.BandSet(640, 360, [])
