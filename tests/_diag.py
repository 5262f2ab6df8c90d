"""Debug helper (not a test): print the worst mismatching texels of one image between oracle and CUDA."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from _harness import Pair, _decode


def dump(p, fi, name, n=8):
    ref = p.oracle.image(name, fi); got = p.cuda.download(name, fi)
    a, _ = _decode(name, ref); b, _ = _decode(name, got)
    with np.errstate(invalid="ignore"):
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    idx = np.argwhere(bad)
    print(name, "mismatching comps:", len(idx))
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b); d[np.isnan(d)] = np.inf
    order = np.argsort(-d[bad])[:n]
    for k in order:
        i = tuple(idx[k])
        print("  at", i, "ref", a[i[:2]] if a.ndim == 3 else a[i], "got", b[i[:2]] if b.ndim == 3 else b[i])


if __name__ == "__main__":
    case = sys.argv[1]
    if case == "five":
        p = Pair(256, 144, iterations=5)
        fi, u = p.run_frame(0)
        for name in ("DiffPingColorAndVariance", "DiffPongColorAndVariance", "DiffColorHistory", "AtrousFilteredVariance"):
            dump(p, fi, name)
        dv = p.oracle.image("DiffPingColorAndVariance", fi).astype(np.float32)
        print("ref nonfinite count", (~np.isfinite(dv)).sum())


def atrous_taps(p, fi, it, x, y):
    """Recompute the oracle's per-tap weights of a-trous iteration `it` at checkerboard pixel (x,y) in numpy fp32."""
    import ctypes
    import oracle as om
    from raytracedgl1_b200 import layout
    o = p.oracle
    step = 1 << it
    src = {1: "DiffColorHistory"}.get(it, "DiffPongColorAndVariance" if it % 2 else "DiffPingColorAndVariance")
    dv = o.image(src, fi).astype(np.float32)
    depth = o.image("DepthWorld", fi).astype(np.float32)
    nrm = o.image("Normal", fi)
    mr = o.image("MetallicRoughness", fi)
    fv = o.image("AtrousFilteredVariance", fi).astype(np.float32)
    gd = o.image("DepthGrad", fi).astype(np.float32)
    out = (ctypes.c_float * 3)()
    def dn(v):
        om.lib().asvgfOracleDecodeNormal(int(v), out); return np.array(out, dtype=np.float32)
    n = dn(nrm[y, x]); l = layout.luminance(dv[y, x, :3])
    print("centre", (x, y), "dv", dv[y, x], "depth", depth[y, x], "sigma", fv[y, x], "gd", gd[y, x], "lum", l)
    for yy in (-1, 0, 1):
        for xx in (-1, 0, 1):
            if xx == 0 and yy == 0: continue
            qx, qy = x + xx * step, y + yy * step
            if not (0 <= qx < dv.shape[1] and 0 <= qy < dv.shape[0]):
                print("  tap", (xx, yy), "OOB"); continue
            nq = dn(nrm[qy, qx]); nn = max(0.0, float(np.dot(n, nq)))
            wz = abs(depth[y, x] - depth[qy, qx]) / max(gd[y, x] * (abs(xx) + abs(yy)), 0.01)
            lq = layout.luminance(dv[qy, qx, :3])
            wl = abs(l - lq) / (4 * fv[y, x] + 1e-5)
            print("  tap", (xx, yy), "varQ", dv[qy, qx, 3], "nN", nn, "wz", wz, "wl", wl, "lq", lq, "depthQ", depth[qy, qx])


if __name__ == "__main__" and sys.argv[1] == "five2":
    from raytracedgl1_b200 import layout
    p = Pair(256, 144, iterations=5)
    u = p.feed_inputs(0, 0)
    for ps in range(0, 13):
        p.sync_state_to_cuda()
        if ps == 11:
            break
        p.oracle.run_pass(ps, 0, u)
    # state before pass 11 (iteration 3, step 8) is identical on both sides
    import copy
    pre = {n: p.oracle.image(n, 0).copy() for n in ("DiffPongColorAndVariance",)}
    p.oracle.run_pass(11, 0, u); p.cuda.denoise_pass(11, 0, u); p.cuda.synchronize()
    ref = p.oracle.image("DiffPingColorAndVariance", 0).astype(np.float32); got = p.cuda.download("DiffPingColorAndVariance", 0).astype(np.float32)
    bad = np.argwhere(np.isinf(ref[..., 3]) & np.isnan(got[..., 3]))
    print("inf-vs-nan:", len(bad), bad[:12].tolist())
    # the oracle reads its own (already overwritten?) images: DiffPong is the source, untouched by pass 11
    for (y, x) in bad[:2]:
        atrous_taps(p, 0, 3, int(x), int(y))
