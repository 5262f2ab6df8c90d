"""One large frame split into horizontal bands, one band per GPU / process (BASELINE config 4).

The reference has no multi-GPU path; this is the host-side schedule that makes N handles of
libasvgf.so -- each created with `bandRowBegin/bandRowEnd/bandHaloRows` -- produce exactly the single-GPU frame.
A pass COMPUTES the rows its rank owns and READS up to `reach` rows beyond them; before the pass, the rows of the
images it is about to read that the two neighbouring ranks computed are exchanged (`torch.distributed` send/recv:
NCCL over NVLink on GPUs, gloo in the CPU tests).  Reach per pass, from the reference shaders:

    gradient a-trous i ......... 2^i gradient rows of Ping/Pong            CmASVGFGradientAtrous.comp:57-60
    gradient samples, temporal . the motion-vector reach (<= halo) of the previous frame's state images
                                                             RtGradients.rgen:82-94, CmSVGFTemporalAccumulation.comp:149-178
    anti-firefly ............... 2 rows of the temporal outputs            CmAntiFirefly.comp:31
    variance ................... 1 row of DiffAccumColor                   CmSVGFEstimateVariance.comp:59
    a-trous i .................. 2^i rows of the three colour images        CmSVGFAtrous.comp:61,126-140

Two schedules.  `frame_schedule`: one exchange before every pass that reads beyond the band (13 per frame; each is a
few hundred KB, so its cost is NCCL launch latency, not bandwidth).  `frame_schedule_apron` (the default of
BandedDenoiser): the handle is created with ASVGF_CREATE_BAND_APRON, every pass also computes the few rows beyond the
band that the later passes will read (18 rows for the temporal pass at 4 a-trous iterations, 3.3 % of a 540-row band),
and a frame needs only two exchanges: the previous frame's state images before the first pass and the gradient samples
after the first.  Both produce the single-GPU frame bit for bit.

Producer images (the G-buffer) are uploaded with their halo rows, like a renderer that rasterises the band plus an
apron; only images the denoiser itself writes are exchanged.  All images are torch tensors bound with
asvgfBindExternalImage, so a halo is a contiguous row slice and needs no packing.
"""
import numpy as np

from . import layout

STRATA = 3


def band_rows(height, world, row_cost=None):
    """Row boundaries b[0..world]: multiples of 3 (a 3x3 gradient stratum never straddles two bands), b[world] = height.
    row_cost: optional relative cost of each of the `height` rows; the bands then get equal summed cost instead of
    equal height (the passes skip sky pixels and take a longer path on disoccluded ones, so rows are not equally
    expensive: at 8K the top half of the synthetic scene costs 3.0 ms and the bottom half 3.9 ms)."""
    if row_cost is None:
        b = [min(height, int(round(k * height / world / STRATA)) * STRATA) for k in range(world)] + [height]
    else:
        c = np.cumsum(np.asarray(row_cost, dtype=np.float64))
        if c.shape[0] != height or not np.all(np.diff(c) >= 0) or c[-1] <= 0:
            raise ValueError("row_cost must hold one non-negative cost per row")
        b = [0]
        for k in range(1, world):
            y = int(np.searchsorted(c, c[-1] * k / world)) + 1   # rows [0, y) hold k/world of the cost
            b.append(min(height, int(round(y / STRATA)) * STRATA))
        b.append(height)
    for k in range(world):
        if b[k + 1] <= b[k]:
            raise ValueError("frame of %d rows is too small for %d bands" % (height, world))
    return b


def rebalanced_rows(boundaries, band_ms, min_rows=0):
    """New boundaries from the measured kernel time of each band: the cost per row is taken as constant inside each
    of the old bands (band_ms[k] / rows of band k) and the new bands get equal summed cost."""
    height, world = boundaries[-1], len(boundaries) - 1
    cost = np.zeros(height)
    for k in range(world):
        cost[boundaries[k]:boundaries[k + 1]] = max(band_ms[k], 1e-6) / (boundaries[k + 1] - boundaries[k])
    b = band_rows(height, world, cost)
    for k in range(1, world):  # keep every band tall enough to hold its neighbours' halos
        b[k] = max(b[k], b[k - 1] + (min_rows + STRATA - 1) // STRATA * STRATA)
    for k in range(world - 1, 0, -1):
        b[k] = min(b[k], b[k + 1] - (min_rows + STRATA - 1) // STRATA * STRATA)
    return b


class BandPlan:
    """Rows a rank owns and stores, in full-resolution and in gradient (1/3) resolution."""

    def __init__(self, height, world, rank, halo_rows=48, boundaries=None):
        if halo_rows % STRATA:
            raise ValueError("halo_rows must be a multiple of 3")
        self.height, self.world, self.rank, self.halo = height, world, rank, halo_rows
        b = list(boundaries) if boundaries is not None else band_rows(height, world)
        if len(b) != world + 1 or b[0] != 0 or b[-1] != height or any(x % STRATA for x in b[:-1]):
            raise ValueError("band boundaries must start at 0, end at the frame height and be multiples of 3")
        self.boundaries = b
        if world > 1 and min(b[k + 1] - b[k] for k in range(world)) < halo_rows:
            raise ValueError("bands must be at least halo_rows (%d) rows tall: halos come from the direct neighbours only" % halo_rows)
        self.own = (b[rank], b[rank + 1])
        self.stored = (max(0, self.own[0] - halo_rows), min(height, self.own[1] + halo_rows))
        gh = max(1, (height + 1) // STRATA)
        self.gown = (self.own[0] // STRATA, gh if self.own[1] == height else self.own[1] // STRATA)
        self.gstored = (max(0, self.gown[0] - halo_rows // STRATA), min(gh, self.gown[1] + halo_rows // STRATA))
        self.up = rank - 1 if rank > 0 else None
        self.down = rank + 1 if rank + 1 < world else None

    def rows(self, third_res):
        return (self.gown, self.gstored) if third_res else (self.own, self.stored)


def exchange_halos(tensors, depth, plan, third_res=False, dist=None, group=None):
    """tensors: list of 2-D (stored rows, pitch bytes) tensors of images this rank just wrote on its OWN rows.
    Sends the `depth` outermost own rows to each neighbour and receives the neighbours' into the halo rows."""
    exchange_halos_multi([(tensors, depth, third_res)], plan, dist, group)


def exchange_halos_multi(specs, plan, dist=None, group=None):
    """specs: list of (tensors, depth, third_res); all of them travel in ONE batch of sends/receives."""
    if plan.world == 1:
        return
    import torch.distributed as td
    dist = dist or td
    ops = []
    for tensors, depth, third_res in specs:
        ops += _halo_ops(tensors, depth, plan, third_res, dist, group)
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def _halo_ops(tensors, depth, plan, third_res, dist, group):
    (o0, o1), (s0, s1) = plan.rows(third_res)
    ops = []
    for t in tensors:
        if plan.up is not None:
            d = min(depth, o0 - s0)
            # the neighbour above owns [.., o0): it wants my top rows, I want its bottom rows
            ops.append(dist.P2POp(dist.isend, t[o0 - s0:o0 - s0 + min(depth, o1 - o0)], plan.up, group))
            if d > 0:
                ops.append(dist.P2POp(dist.irecv, t[o0 - s0 - d:o0 - s0], plan.up, group))
        if plan.down is not None:
            d = min(depth, s1 - o1)
            ops.append(dist.P2POp(dist.isend, t[o1 - s0 - min(depth, o1 - o0):o1 - s0], plan.down, group))
            if d > 0:
                ops.append(dist.P2POp(dist.irecv, t[o1 - s0:o1 - s0 + d], plan.down, group))
    return ops


def atrous_routing(iteration):
    """(source, destination) colour images of a-trous iteration i (CmSVGFAtrous_Iter0.comp:312-334,
    CmSVGFAtrous.comp:237-276; the same table as runAtrous in csrc/asvgf.cu)."""
    if iteration == 0:
        return (("DiffPingColorAndVariance", "SpecPingColor", "IndirPing"), ("DiffColorHistory", "SpecPongColor", "IndirPong"))
    if iteration % 2 == 1:
        return (("DiffColorHistory" if iteration == 1 else "DiffPongColorAndVariance", "SpecPongColor", "IndirPong"),
                ("DiffPingColorAndVariance", "SpecPingColor", "IndirPing"))
    return (("DiffPingColorAndVariance", "SpecPingColor", "IndirPing"), ("DiffPongColorAndVariance", "SpecPongColor", "IndirPong"))


def frame_schedule(iterations):
    """The frame as a list of steps: ("pass", pass id) or ("exchange", [logical image names], rows, third_res).
    Image names are LOGICAL (cur/prev swap still to be applied for the frame index)."""
    full = None  # depth None = the whole halo
    s = [("exchange", ["IndirAccum_Prev", "DiffColorHistory", "DiffAccumMoments_Prev", "AccumHistoryLength_Prev",
                       "SpecAccumColor_Prev"], full, False),
         ("exchange", ["DISGradientHistory"], full, True),
         ("pass", layout.PASS_GRADIENT_SAMPLES)]
    for i in range(4):
        s.append(("exchange", ["DISPingGradient" if i % 2 == 0 else "DISPongGradient"], 1 << i, True))
        s.append(("pass", layout.PASS_GRADIENT_ATROUS_0 + i))
    s += [("pass", layout.PASS_TEMPORAL),
          ("exchange", ["DiffTemporary", "SpecAccumColor", "IndirAccum"], 2, False),
          ("pass", layout.PASS_ANTIFIREFLY),
          ("exchange", ["DiffAccumColor", "SpecPingColor", "IndirPing"], 1, False),
          ("pass", layout.PASS_VARIANCE),
          ("exchange", ["DiffPingColorAndVariance"], 1, False)]
    for i in range(iterations):
        if i > 0:
            s.append(("exchange", list(atrous_routing(i)[0]), 1 << i, False))
        s.append(("pass", layout.PASS_ATROUS_0 + i))
    return s


def frame_schedule_apron(iterations, apron_rows):
    """Schedule for handles created with ASVGF_CREATE_BAND_APRON.  apron_rows(pass) = asvgfBandApronRows.  A step
    ("exchange", [(names, rows, third_res), ...]) is ONE batch of sends/receives."""
    full = None
    s = [("exchange", [(["IndirAccum_Prev", "DiffColorHistory", "DiffAccumMoments_Prev", "AccumHistoryLength_Prev",
                         "SpecAccumColor_Prev"], full, False), (["DISGradientHistory"], full, True)]),
         ("pass", layout.PASS_GRADIENT_SAMPLES),
         ("exchange", [(["DISPingGradient"], apron_rows(layout.PASS_GRADIENT_SAMPLES), True)])]
    s += [("pass", p) for p in range(layout.PASS_GRADIENT_ATROUS_0, layout.PASS_ATROUS_0 + iterations)]
    return s


def apron_halo_rows(iterations, apron_rows, motion_rows=48):
    """Smallest halo (multiple of 3) that holds the gradient-sample exchange and the temporal pass's apron plus a
    motion-vector reach of `motion_rows` rows into the previous frame's state."""
    need = max(STRATA * apron_rows(layout.PASS_GRADIENT_SAMPLES), apron_rows(layout.PASS_TEMPORAL) + motion_rows)
    return (need + STRATA - 1) // STRATA * STRATA


_SWAPPABLE = {n for n in layout.IMAGES if n + "_Prev" in layout.IMAGES}


def physical_name(name, frame_index):
    """Framebuffers.cpp:35-50: with frameIndex 1 a swappable image and its _Prev twin trade places."""
    if frame_index == 0:
        return name
    if name in _SWAPPABLE:
        return name + "_Prev"
    if name.endswith("_Prev") and name[:-5] in _SWAPPABLE:
        return name[:-5]
    return name


class BandedDenoiser:
    """One rank's band of a frame.  Mirrors Denoiser (gradient_samples + denoise = `frame`), plus the halo exchanges."""

    def __init__(self, width, height, atrous_iterations=4, rank=0, world=1, device=0, halo_rows=None, flags=0, apron=True,
                 boundaries=None):
        import torch
        from . import denoiser as dn
        self.torch = torch
        self.apron = bool(apron) and world > 1
        rows_of = lambda p: dn.band_apron_rows(atrous_iterations, p)  # noqa: E731
        if halo_rows is None:
            halo_rows = apron_halo_rows(atrous_iterations, rows_of) if self.apron else 48
        self.plan = BandPlan(height, world, rank, halo_rows, boundaries)
        need = 8 << max(0, atrous_iterations - 4)
        if world > 1 and halo_rows < max(need, 45):
            raise ValueError("halo_rows must cover the a-trous reach (%d rows) and the gradient a-trous reach (45)" % need)
        if self.apron and halo_rows < max(STRATA * rows_of(layout.PASS_GRADIENT_SAMPLES), rows_of(layout.PASS_TEMPORAL)):
            raise ValueError("halo_rows too small for the apron schedule")
        self.width, self.height, self.iterations = width, height, atrous_iterations
        band = None if world == 1 else self.plan.own
        if self.apron:
            flags |= dn.CREATE_BAND_APRON
        self.d = dn.Denoiser(width, height, atrous_iterations, flags=flags | dn.CREATE_EXTERNAL_IMAGES, device=device,
                             band=band, halo_rows=halo_rows if world > 1 else 0)
        self.images = {}
        for name in layout.IMAGES:
            if name == "DebugMask" and not (flags & dn.CREATE_EXPORT_DEBUG_MASK):
                continue
            _, pitch, _, _, rows = self.d.image(name, 0)
            t = torch.zeros((rows, pitch), dtype=torch.uint8, device=torch.device("cuda", device))
            self.images[name] = t
            self.d.bind_external(name, t.data_ptr(), pitch)
        self.schedule = frame_schedule_apron(atrous_iterations, rows_of) if self.apron else frame_schedule(atrous_iterations)

    def close(self):
        self.d.close()

    def stored_rows(self, name):
        third = layout.IMAGES[name][2]
        return self.plan.rows(third)[1]

    def upload(self, name, frame_index, full_image):
        """full_image: numpy array of the WHOLE frame; the rows this rank stores (band + halo) are uploaded."""
        s0, s1 = self.stored_rows(name)
        self.d.upload(name, frame_index, np.ascontiguousarray(full_image[s0:s1]))

    def download_own(self, name, frame_index):
        """The rows this rank owns, as a numpy array."""
        third = layout.IMAGES[name][2]
        (o0, o1), (s0, _) = self.plan.rows(third)
        return self.d.download(name, frame_index)[o0 - s0:o1 - s0]

    def frame(self, frame_index, uniform, stream=0, events=None):
        """= CalculateGradientsSamples + Denoiser::Denoise for this band, with the halo exchanges in between.
        events: optional list; receives (kind, start event, end event) per step (current torch stream) for profiling."""
        for step in self.schedule:
            if events is not None:
                e0 = self.torch.cuda.Event(enable_timing=True)
                e0.record()
            self._step(step, frame_index, uniform, stream)
            if events is not None:
                e1 = self.torch.cuda.Event(enable_timing=True)
                e1.record()
                events.append((step[0], e0, e1))

    def _step(self, step, frame_index, uniform, stream):
        if step[0] == "pass":
            if step[1] == layout.PASS_GRADIENT_SAMPLES:
                self.d.gradient_samples(frame_index, uniform, stream=stream)
            else:
                self.d.denoise_pass(step[1], frame_index, uniform, stream=stream)
        else:
            specs = step[1] if self.apron else [(step[1], step[2], step[3])]
            batch = []
            for names, depth, third in specs:
                tensors = [self.images[physical_name(n, frame_index)] for n in names]
                depth = (self.plan.halo // STRATA if third else self.plan.halo) if depth is None else depth
                batch.append((tensors, depth, third))
            exchange_halos_multi(batch, self.plan)
