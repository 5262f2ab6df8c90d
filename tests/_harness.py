"""Shared parity harness: drives the CPU oracle and the CUDA path (through the C ABI) on identical inputs/state and
compares every image of the table in its storage format.

Tolerances (north_star: "fp16-output max abs error <= 1e-3 relative to luminance, integer/boolean outputs bit-exact"):
  * integer images (GradientPrevPix, DebugMask, IsSky) ............ bit-exact
  * fp16 images ..................................................... |d| <= REL_TOL * max(|ref| scale, EPS)
  * packed E5B9G9R9 / B10G11R11 / unorm8 images ..................... the same bound on decoded values PLUS one
    quantisation step of the storage format (9-, 6/5- and 8-bit mantissas are coarser than 1e-3, so a 1-ulp fp32
    difference before the store can legitimately flip one code), and >= MIN_EXACT of the texels bit-identical.
"""
import numpy as np

import oracle as oracle_mod
import synth
from raytracedgl1_b200 import layout

REL_TOL = 1e-3
EPS = 1e-3
MIN_EXACT = 0.99

STATE_AND_SCRATCH = [n for n, v in layout.IMAGES.items()]


class Pair:
    def __init__(self, width, height, iterations=4, seed=synth.DEFAULT_SEED, flags=None, with_cuda=True):
        from raytracedgl1_b200 import denoiser as dn
        self.W, self.H, self.N, self.seed = width, height, iterations, seed
        self.oracle = oracle_mod.Oracle(width, height, iterations)
        self.cuda = None
        if with_cuda:
            self.cuda = dn.Denoiser(width, height, iterations,
                                    flags=dn.CREATE_EXPORT_DEBUG_MASK if flags is None else flags)

    def close(self):
        self.oracle.close()
        if self.cuda:
            self.cuda.close()

    def feed_inputs(self, frame, fi, mutate=None):
        g = synth.generate(self.W, self.H, frame, self.seed)
        if mutate:
            mutate(g)
        for name, arr in g.items():
            self.oracle.image(name, fi)[...] = arr
            if self.cuda:
                self.cuda.upload(name, fi, arr)
        return synth.uniform(self.W, self.H, frame)

    def sync_state_to_cuda(self):
        """Make the CUDA side's whole image table bit-identical to the oracle's."""
        for name in layout.IMAGES:
            self.cuda.upload(name, 0, np.ascontiguousarray(self.oracle.image(name, 0)))

    def run_frame(self, frame, uniform_edit=None, mutate=None):
        fi = frame & 1
        u = self.feed_inputs(frame, fi, mutate)
        if uniform_edit:
            uniform_edit(u)
        self.oracle.gradient_samples(fi, u)
        self.oracle.denoise(fi, u)
        if self.cuda:
            self.cuda.gradient_samples(fi, u)
            self.cuda.denoise(fi, u)
            self.cuda.synchronize()
        return fi, u

    def compare(self, fi, names=None):
        out = {}
        for name in (names or layout.IMAGES):
            ref = self.oracle.image(name, fi)
            got = self.cuda.download(name, fi)
            out[name] = compare_image(name, ref, got)
        return out


def _decode(name, arr):
    fmt = layout.IMAGES[name][1]
    if fmt == layout.R32_UINT and name in layout.E5_IMAGES:
        return layout.decode_e5(arr), "e5"
    if fmt == layout.B10G11R11_UFLOAT:
        return layout.decode_b10g11r11(arr), "uf"
    if fmt in (layout.R8G8_UNORM, layout.R8G8B8A8_UNORM):
        return arr.astype(np.float32) / 255.0, "unorm"
    if fmt in (layout.R16_SFLOAT, layout.R16G16_SFLOAT, layout.R16G16B16A16_SFLOAT, layout.R32G32B32A32_SFLOAT):
        return arr.astype(np.float32), "float"
    return None, "int"


def compare_image(name, ref, got):
    """-> dict(exact_frac, max_err_over_tol, ok, n_bad)."""
    assert ref.shape == got.shape and ref.dtype == got.dtype, (name, ref.shape, got.shape)
    same = (ref.view(np.uint8) == got.view(np.uint8))
    exact_frac = float(same.mean()) if same.size else 1.0
    dref, kind = _decode(name, ref)
    if kind == "int":
        nbad = int((ref != got).sum())
        return dict(kind=kind, exact_frac=exact_frac, n_bad=nbad, ok=nbad == 0, worst=float(nbad > 0))
    dgot, _ = _decode(name, got)
    with np.errstate(invalid="ignore", over="ignore"):
        both_nan = np.isnan(dref) & np.isnan(dgot)
        both_inf = np.isinf(dref) & (dref == dgot)
        err = np.abs(dgot - dref)
        err[both_nan | both_inf] = 0.0
        err[np.isnan(err)] = np.inf
        if dref.ndim == 3 and dref.shape[-1] >= 3 and kind in ("e5", "uf"):
            scale = np.maximum(layout.luminance(np.abs(dref)), EPS)[..., None]
        else:
            scale = np.maximum(np.abs(dref), EPS)
        if not np.all(np.isfinite(scale)):
            scale = np.where(np.isfinite(scale), scale, 1.0)
        tol = REL_TOL * scale
        if kind == "e5":      # one step of the 9-bit shared-exponent mantissa
            tol = tol + np.max(np.abs(dref), axis=-1, keepdims=True) * 2.0 ** -8 + 2.0 ** -24
        elif kind == "uf":    # one step of a 6-bit (R,G) / 5-bit (B) mantissa
            tol = tol + np.abs(dref) * np.array([2.0 ** -6, 2.0 ** -6, 2.0 ** -5], dtype=np.float32) + 2.0 ** -20
        elif kind == "unorm":
            tol = tol + 1.0 / 255.0
        elif ref.dtype == np.float16:   # half an fp16 ulp of slack for a flipped rounding
            tol = tol + np.abs(dref) * 2.0 ** -11 + 2.0 ** -24
        ratio = err / tol
    worst = float(np.max(ratio)) if ratio.size else 0.0
    nbad = int((ratio > 1.0).sum())
    ok = nbad == 0 and (kind == "float" or exact_frac >= MIN_EXACT)
    return dict(kind=kind, exact_frac=exact_frac, n_bad=nbad, ok=ok, worst=worst)


def assert_all_ok(results, context=""):
    bad = {k: v for k, v in results.items() if not v["ok"]}
    assert not bad, "%s parity failures: %s" % (context, bad)
