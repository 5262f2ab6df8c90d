"""Caller-owned images whose row pitches are TIGHT and differ from image to image -- what a Vulkan host's linear images look like
 -- instead of the library's align256(width) texels per row: selects the UNIFORM = false forms of the
temporal and a-trous kernels (one offset per image instead of one shared texel offset).  Gate: every image bit-identical to the
owned-image handle.  Written in the last session of round 2 from tests/hostsim_cases.py::test_tight_and_unequal_external_pitches,
which passes on the host build (addresses and vector alignment checked there under the sanitizers); passed on a B200 with the
round's last GPU seconds (profiles/r02_final_gpu_check_b.log)."""
import ctypes

import numpy as np
import pytest

import synth
from raytracedgl1_b200 import layout

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,extra", [((192, 108), 0), ((322, 182), 1)])
def test_tight_and_unequal_external_pitches(size, extra):
    import torch
    from raytracedgl1_b200 import denoiser as dn
    W, H = size
    own = dn.Denoiser(W, H, flags=dn.CREATE_EXPORT_DEBUG_MASK)
    ext = dn.Denoiser(W, H, flags=dn.CREATE_EXTERNAL_IMAGES | dn.CREATE_EXPORT_DEBUG_MASK)
    keep = {}
    try:
        for name in layout.IMAGES:
            idx = layout.INDEX[name]
            f, w, h, b = [ctypes.c_uint32() for _ in range(4)]
            assert dn.lib().asvgfGetImageInfo(idx, W, H, ctypes.byref(f), ctypes.byref(w), ctypes.byref(h), ctypes.byref(b)) == 0
            pitch = ((w.value * b.value + 15) // 16) * 16 + 16 * extra * (1 + idx % 3)
            t = torch.zeros(pitch * h.value, dtype=torch.uint8, device="cuda")      # (the allocator's base is 512-byte aligned)
            keep[name] = t
            ext.bind_external(name, t.data_ptr(), pitch)
        torch.cuda.synchronize()
        for frame in range(3):
            fi = frame & 1
            g, u = synth.generate(W, H, frame), synth.uniform(W, H, frame)
            for d in (own, ext):
                for name, arr in g.items():
                    d.upload(name, fi, arr)
                d.gradient_samples(fi, u)
                d.denoise(fi, u)
                d.synchronize()
            for name in layout.IMAGES:
                a, b2 = own.download(name, fi), ext.download(name, fi)
                assert np.array_equal(a.view(np.uint8), b2.view(np.uint8)), (W, H, frame, name)
    finally:
        own.close()
        ext.close()
