#!/usr/bin/env python
"""bench.py -- A-SVGF denoise throughput (Mpixel/s, ms/frame) on N B200s, with the HBM roofline of the dominant
kernel and the CPU oracle timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 4k|4k-n5|1080p|720p]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one frame of the hot path: asvgfGradientSamples + the Denoiser::Denoise pass chain (gradient a-trous x4,
temporal, anti-firefly, variance, a-trous xN with composition) over one synthetic G-buffer.
  value : whole-job Mpixel/s with inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e   : the same metric through the C ABI with HOST buffers: every step uploads that frame's producer images from
          pinned host memory and reads PreFinal + HistogramInput back, copies inside the timed region
  N > 1 : one process per GPU, each denoising its own independent view stream (weak scaling, no data-path collective)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, a-trous iterations, description)
    "4k": (3840, 2160, 4, "3840x2160 synthetic G-buffer, full A-SVGF chain (gradient samples + gradient a-trous x4 + "
                          "temporal + anti-firefly + variance + 4 a-trous iterations + composition), 1 view stream per GPU"),
    "4k-n5": (3840, 2160, 5, "3840x2160 synthetic G-buffer, full chain with 5 a-trous iterations (BASELINE config 3)"),
    "1080p": (1920, 1080, 4, "1920x1080 synthetic G-buffer, full chain, 4 a-trous iterations"),
    "720p": (1280, 720, 4, "1280x720 synthetic G-buffer, full chain, 4 a-trous iterations"),
    # BASELINE config 4: ONE frame split into horizontal bands across the ranks (strong scaling, halo exchange per pass)
    "8k-bands": (7680, 4320, 4, "7680x4320 synthetic G-buffer, full A-SVGF chain (4 a-trous iterations), one frame split into "
                                "horizontal bands across the GPUs with per-pass halo exchange over NCCL/NVLink"),
    "4k-bands": (3840, 2160, 4, "3840x2160 synthetic G-buffer, full chain, one frame split into horizontal bands across the GPUs"),
}
METRIC = "asvgf_denoise_throughput"
UNIT = "Mpixel/s"


def measured_traffic(kernel, W, H, N):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full
    capture (profiles/); None when no capture exists for this workload."""
    try:
        name = "r02_dram_traffic_4k.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_dram_traffic_4k.json")) else "r01_dram_traffic_4k.json"
        with open(os.path.join(ROOT, "profiles", name)) as f:
            t = json.load(f)
        if (W, H, N) == (3840, 2160, 4):
            return t["kernels"][kernel]["bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:  # noqa: BLE001
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        while self.ok and not self._stop_evt.is_set():
            try:
                self.sample()
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


REFERENCE_KIND = ("the reference's own compute shaders (unmodified GLSL text of Source/Shaders/Cm*.comp compiled as C++ over its "
                  "vendored GLM: oracle/_ref) on the host cores; Vulkan/lavapipe does not exist in this image")


class _RefSide:
    """The reference arm's implementation: oracle/_ref (kind 'reference') when its prebuilt library travelled with the
    repo, else the hand-written oracle port.  The gradient-sample stage (a ray-generation shader) has no compute form:
    with _ref it is run by the oracle on a copy of the state (a few percent of the step)."""

    def __init__(self, W, rows, N):
        import oracle
        from oracle import ref
        from raytracedgl1_b200 import layout
        self.layout = layout
        self.o = oracle.Oracle(W, rows, N)
        self.kind = "port"
        self.r = None
        if N == 4 and ref.available():          # CmSVGFAtrous.comp is written for exactly four iterations (its #error guard)
            self.r = ref.RefDenoiser(W, rows)
            self.kind = "reference"

    def image(self, name, fi):
        return (self.r or self.o).image(name, fi)

    def frame(self, fi, u):
        if self.r is None:
            self.o.gradient_samples(fi, u)
            self.o.denoise(fi, u)
            return
        for n in ("Motion", "GradientInputs_Prev", "IsSky", "IndirAccum_Prev", "FwdLum", "DISGradientHistory", "GradientPrevPix"):
            self.o.image(n, fi)[...] = self.r.image(n, fi)
        self.o.gradient_samples(fi, u)
        for n in ("DISPingGradient", "GradientPrevPix"):
            self.r.image(n, fi)[...] = self.o.image(n, fi)
        self.r.denoise(fi, u)

    def close(self):
        self.o.close()


def run_reference(args, rank, world):
    """The reference arm for this tier: the reference's own CPU-executable form of the path -- oracle/_ref, i.e. the
    unmodified shader text compiled for the host (kind 'reference'); the oracle port when that library is absent -- on all
    host threads, each step a bounded band of the same workload.  The line also carries the (faster) oracle port's rate,
    so that the ratio the driver computes can be read against both."""
    if rank != 0:
        return
    import oracle
    import synth
    W, H, N, desc = WORKLOADS[args.workload]
    oracle.set_threads(0)
    cores = oracle.get_threads()

    def rate_of(side_factory, cal_w=640, cal_h=180):
        side = side_factory(cal_w, cal_h, N)
        g = synth.generate(cal_w, cal_h, 1)
        for k, v in g.items():
            side.image(k, 1)[...] = v
        u = synth.uniform(cal_w, cal_h, 1)
        side.frame(1, u)
        t0 = time.perf_counter()
        side.frame(1, u)
        r = cal_w * cal_h / (time.perf_counter() - t0)          # pixels per second
        kind = side.kind
        side.close()
        return r, kind

    rate, kind = rate_of(_RefSide)
    budget_s = 100.0                                            # (steps + warmup) steps take about this long
    rows = int(rate * budget_s / (args.steps + args.warmup) / W)
    rows = max(36, min(H, rows - rows % 3))
    side = _RefSide(W, rows, N)
    frames = []
    for f in (0, 1):                                            # the top `rows` rows of two consecutive full-width frames
        frames.append((synth.generate(W, rows, f), synth.uniform(W, rows, f)))

    def step(i):
        fi = i & 1
        g, u = frames[fi]
        for k, v in g.items():
            side.image(k, fi)[...] = v
        side.frame(fi, u)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    side.close()
    value = W * rows * args.steps / dt / 1e6
    sample = "%dx%d band (top rows of the %dx%d frame) per step, %d steps" % (W, rows, W, H, args.steps)
    port = None
    if kind == "reference":                                     # the hand-written port, for context (about 10 s)
        class _Port(_RefSide):
            def __init__(self, W, rows, N):
                import oracle as om
                self.o, self.r, self.kind = om.Oracle(W, rows, N), None, "port"
        prate, _ = rate_of(_Port, 1280, 360)
        port = {"value": prate / 1e6, "unit": UNIT, "cores": cores, "kind": "port", "sample": "one 1280x360 frame with history"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": run_config(args.workload, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "what": REFERENCE_KIND if kind == "reference" else "CPU oracle port of the GLSL shaders (oracle/_ref not present)"},
        "cpu_port": port,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_config(workload, n_frames):
    """The `config` object both arms print (the driver compares them key by key)."""
    W, H, N, desc = WORKLOADS[workload]
    return {"workload": desc, "width": W, "height": H, "atrous_iterations": N, "streams_per_gpu": 1,
            "parallelism": "independent view streams, 1 per GPU, no collective",
            "l2": "no flush: per-stream images total %.2f GB, each pass streams > 126 MB L2" % (201 * W * H / 1e9)}


def cpu_baseline(W, H, N, frames):
    """Oracle (kind 'port') on all host cores over a bounded sample of the same workload: one warm-up frame and one
    timed frame of the top band of the frame, sized from a quick calibration to about 10-20 s of CPU work."""
    import oracle
    oracle.set_threads(0)
    cores = oracle.get_threads()
    o = oracle.Oracle(640, 360, N)
    import synth
    g = synth.generate(640, 360, 1)
    for k, v in g.items():
        o.image(k, 1)[...] = v
    u = synth.uniform(640, 360, 1)
    o.gradient_samples(1, u); o.denoise(1, u)
    t0 = time.perf_counter()
    o.gradient_samples(1, u); o.denoise(1, u)
    rate = 640 * 360 / (time.perf_counter() - t0)
    o.close()
    rows = int(rate * 7.0 / W)
    rows = max(36, min(H, rows - rows % 3))
    o = oracle.Oracle(W, rows, N)
    t_tot = 0.0
    for f in (0, 1):
        g = synth.generate(W, rows, f)
        u = synth.uniform(W, rows, f)
        for k, v in g.items():
            o.image(k, f)[...] = v
        t0 = time.perf_counter()
        o.gradient_samples(f, u)
        o.denoise(f, u)
        t_tot = time.perf_counter() - t0                 # keep the second (history present) frame
    o.close()
    return {"value": W * rows / t_tot / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "1 warm-up + 1 timed frame of a %dx%d band (top rows of the %dx%d frame)" % (W, rows, W, H)}


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch

    import synth
    from raytracedgl1_b200 import denoiser as dn
    from raytracedgl1_b200 import layout

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the A-SVGF path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H, N, desc = WORKLOADS[args.workload]
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    # the latency-bound gradient chain (gradient samples + 4 gradient a-trous dispatches) is enqueued on its own stream:
    # the library orders it against the temporal passes itself (asvgf.h, ASVGF_CREATE_EAGER_GRADIENT_ATROUS), so the chain
    # of frame t+1 overlaps the a-trous iterations of frame t
    gstream = torch.cuda.Stream()
    gh = gstream.cuda_stream
    d = dn.Denoiser(W, H, N, device=local_rank, flags=dn.CREATE_EAGER_GRADIENT_ATROUS)

    # A cycled sequence of consecutive synthetic frames (camera motion, moving objects, disocclusions, fireflies;
    # independent stream per rank: its own seed).  Host copies live in pinned memory (the e2e leg uploads them),
    # device copies stay resident in HBM and are bound as the producer images of the frame slot before each step
    # (asvgfBindExternalImage = the G-buffer producer handing over its images; no copy).
    seed = synth.DEFAULT_SEED + rank
    n_frames = max(2, args.frames - args.frames % 2)
    n_host = min(n_frames, 4)                      # frames kept in pinned host memory for the e2e leg
    swappable = {n for n in layout.PRODUCER_IMAGES if n + "_Prev" in layout.IMAGES}
    host, resident, uniforms = [], [], []
    scratch = None
    for f in range(n_frames):
        if f < n_host or scratch is None:
            bufs = {}
            for name in layout.PRODUCER_IMAGES:
                dtype, shape = layout.image_numpy_shape(name, W, H)
                t = torch.empty(int(np.prod(shape)) * dtype.itemsize, dtype=torch.uint8, pin_memory=True)
                bufs[name] = t.numpy().view(dtype).reshape(shape)
                bufs["_keep_" + name] = t
            if f >= n_host:
                scratch = bufs
        else:
            torch.cuda.synchronize()               # the previous frame's copies out of the scratch buffers are done
            bufs = scratch
        synth.generate(W, H, f, seed, out={k: v for k, v in bufs.items() if not k.startswith("_keep_")})
        dev = {}
        for name in layout.PRODUCER_IMAGES:
            _, pitch, _, _, rows = d.image(name, 0)
            row_bytes = bufs[name].nbytes // rows
            dt = torch.zeros((rows, pitch), dtype=torch.uint8, device="cuda")
            dt[:, :row_bytes].copy_(bufs["_keep_" + name].view(rows, row_bytes), non_blocking=True)
            dev[name] = (dt, pitch)
        if f < n_host:
            host.append(bufs)
        uniforms.append(synth.uniform(W, H, f))
        resident.append(dev)
    torch.cuda.synchronize()
    # three slot-sized staging sets for the e2e leg (host -> device copies land here; a ring of three lets the upload
    # of frame i+1 overlap the denoising of frame i, whose "previous" images are frame i-1's set)
    staging = []
    for _ in range(3):
        staging.append({name: (torch.zeros_like(resident[0][name][0]), resident[0][name][1]) for name in layout.PRODUCER_IMAGES})
    h2d_bytes = sum(host[0][n].nbytes for n in layout.PRODUCER_IMAGES)
    out_host = {}
    for name in ("PreFinal", "HistogramInput"):
        dtype, shape = layout.image_numpy_shape(name, W, H)
        t = torch.empty(int(np.prod(shape)) * dtype.itemsize, dtype=torch.uint8, pin_memory=True)
        out_host[name] = (t, t.numpy().view(dtype).reshape(shape))
    d2h_bytes = sum(v[1].nbytes for v in out_host.values())

    def bind(fi, images, prev_images=None):
        """Producer images of frame slot fi := `images` (physical image X for slot 0, X_Prev for slot 1: the
        reference's cur/prev swap, Framebuffers.cpp:35-50); optionally the other slot := `prev_images`."""
        for name in layout.PRODUCER_IMAGES:
            t, pitch = images[name]
            d.bind_external(name + "_Prev" if (fi == 1 and name in swappable) else name, t.data_ptr(), pitch)
            if prev_images is not None and name in swappable:
                t, pitch = prev_images[name]
                d.bind_external(name if fi == 1 else name + "_Prev", t.data_ptr(), pitch)

    def upload(fi, f, on=None):
        bufs = host[f]
        for name in layout.PRODUCER_IMAGES:
            d.upload(name, fi, bufs[name], stream=sh if on is None else on)

    # the launches of one frame in dispatch order: gradient samples, gradient a-trous x4, temporal, anti-firefly,
    # variance, N a-trous iterations
    # (the per-pass leg runs them one by one on one stream; pass 0 then carries the whole gradient chain, the four
    # gradient a-trous ids are no-ops, and the guide plane gets its own slot)
    passes = list(range(0, layout.PASS_ATROUS_0)) + [layout.PASS_GUIDE] + list(range(layout.PASS_ATROUS_0, layout.PASS_ATROUS_0 + N))

    def step_production(i):
        f, fi = i % n_frames, i & 1
        u = uniforms[f]
        bind(fi, resident[f])
        d.gradient_samples(fi, u, stream=gh)      # = PathTracer::CalculateGradientsSamples (image-space part) + gradient a-trous
        d.denoise(fi, u, stream=sh)               # = Denoiser::Denoise

    def step_per_pass(i, events):
        f, fi = i % n_frames, i & 1
        u = uniforms[f]
        bind(fi, resident[f])
        for k, p in enumerate(passes):
            if p == 0:
                d.gradient_samples(fi, u, stream=sh)
            else:
                d.denoise_pass(p, fi, u, stream=sh)
            events[k + 1].record(stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: every frame's inputs already in HBM ----
    for i in range(args.warmup):
        step_production(i)
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = d.launch_count
    sampler.start() if sampler.ok else None
    e_begin.record(stream)
    frame_evs = []                            # an event behind every frame: per-frame times (p50 / p95), not part of `value`
    for i in range(args.steps):
        step_production(args.warmup + i)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        frame_evs.append(ev)
    stream.wait_stream(gstream)               # (the last frame's gradient chain is already behind its temporal pass)
    e_end.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = d.launch_count - launches0
    barrier()
    total_ms = e_begin.elapsed_time(e_end)
    per_frame = sorted(a.elapsed_time(b) for a, b in zip([e_begin] + frame_evs[:-1], frame_evs))
    frame_stats = {"p50_ms": per_frame[len(per_frame) // 2], "p95_ms": per_frame[min(len(per_frame) - 1, int(0.95 * len(per_frame)))],
                   "min_ms": per_frame[0], "max_ms": per_frame[-1]}

    # ---- per-pass breakdown (same launches, an event after each): explains the number above, not part of it ----
    pp_steps = max(3, min(args.steps, 50))
    n_ev = len(passes) + 1
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(pp_steps)]
    for i in range(pp_steps):
        evs[i][0].record(stream)
        step_per_pass(args.warmup + args.steps + i, evs[i])
    torch.cuda.synchronize()
    pass_ms = [sum(evs[i][k].elapsed_time(evs[i][k + 1]) for i in range(pp_steps)) / pp_steps for k in range(len(passes))]

    # ---- the two dispatches that consume the frame right after the denoiser: reported next to
    # the metric, not part of it ----
    post_ms = {}
    fi_last = (args.warmup + args.steps + pp_steps - 1) & 1
    u_last = uniforms[(args.warmup + args.steps + pp_steps - 1) % n_frames]
    import copy
    u_vol = copy.copy(u_last)
    u_vol.volumeEnableType, u_vol.volumeScattering = 1, 1500.0   # VOLUME_ENABLE_SIMPLE: the sample is computed in the kernel

    def scatter(fi, u, stream=0):
        d.scatter_accumulate(fi, u_vol, 16.0, stream=stream)

    # LPM control block of the final prepare: any finite block times the same (the reference-generated one ships in the fixture)
    try:
        lpm_block = np.load(os.path.join(ROOT, "tests", "golden", "asvgf_ref_96x54.npz"))["lpm_default"]
    except Exception:  # noqa: BLE001
        lpm_block = np.zeros(96, np.uint32)
    u_fin = copy.copy(u_last)
    u_fin.luminanceWhitePoint, u_fin.bloomThreshold, u_fin.bloomEmissionMultiplier, u_fin.emissionMaxScreenColor = 10.0, 4.0, 16.0, 1.5

    def prepare_final(fi, u, stream=0):
        d.prepare_final(fi, u_fin, lpm_block, stream=stream)

    for name, call in (("k_checkerboard_resolve", d.resolve_checkerboard), ("k_lum_histogram+k_lum_average", d.calculate_exposure),
                       ("k_scatter_accum", scatter), ("k_prepare_final", prepare_final)):
        call(fi_last, u_last, stream=sh)
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        pe[0].record(stream)
        for _ in range(10):
            call(fi_last, u_last, stream=sh)
        pe[1].record(stream)
        torch.cuda.synchronize()
        post_ms[name] = pe[0].elapsed_time(pe[1]) / 10

    # ---- end to end through the C ABI with host buffers ----
    e2e_steps = max(3, min(args.steps, 20))
    # Pipelined the way a caller with host-resident G-buffers would: the upload of frame i+1 (copy stream) overlaps the
    # denoising of frame i; every step still copies that frame's producer images in and reads its result back.
    up = torch.cuda.Stream()
    up_done = [torch.cuda.Event() for _ in range(e2e_steps + 4)]
    comp_done = [torch.cuda.Event() for _ in range(e2e_steps + 4)]

    def enqueue_upload(i):
        f, fi = i % n_host, i & 1
        if i >= 2:
            up.wait_event(comp_done[i - 2])          # ring set i%3 was frame i-3's, read as "previous" by frame i-2
        bind(fi, staging[i % 3])
        upload(fi, f, on=up.cuda_stream)
        up_done[i].record(up)

    def run_step(i, read_back):
        f, fi = i % n_host, i & 1
        stream.wait_event(up_done[i])
        gstream.wait_event(up_done[i])
        bind(fi, staging[i % 3], staging[(i - 1) % 3])
        d.gradient_samples(fi, uniforms[f], stream=gh)
        d.denoise(fi, uniforms[f], stream=sh)
        comp_done[i].record(stream)
        enqueue_upload(i + 1)                        # next frame's copy starts while this frame is being denoised
        if read_back:
            for name, (t, arr) in out_host.items():
                d.download(name, fi, stream=sh, out=arr.ctypes.data)
            stream.synchronize()                     # the caller consumes the frame's result

    torch.cuda.synchronize()
    enqueue_upload(0)
    run_step(0, True); run_step(1, True)             # warm-up (also fills the history)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(2, 2 + e2e_steps):
        run_step(i, True)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)

    # ---- max over ranks ----
    if dist is not None:
        t = torch.tensor([total_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
        ln = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln)
        launches = int(ln[0])

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        value = world * W * H * args.steps / (total_ms * 1e-3) / 1e6
        bpp = [dn.pass_bytes_per_pixel(p, N) for p in passes]
        chain_bpp = sum(bpp)
        per_pass = []
        for k, p in enumerate(passes):
            name = dn.lib().asvgfPassKernelName(p, N).decode()
            gbs = bpp[k] * W * H / (pass_ms[k] * 1e-3) / 1e9 if pass_ms[k] > 0 else 0.0
            per_pass.append({"pass": p, "kernel": name, "ms": round(pass_ms[k], 5), "bytes_per_px": bpp[k],
                             "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)})
        # dominant kernel = the kernel function with the largest share of the step
        share = {}
        for e in per_pass:
            s = share.setdefault(e["kernel"], {"ms": 0.0, "bytes": 0.0, "launches": 0})
            s["ms"] += e["ms"]; s["bytes"] += e["bytes_per_px"] * W * H; s["launches"] += 1
        dom = max(share.items(), key=lambda kv: kv[1]["ms"])
        dom_gbs = dom[1]["bytes"] / (dom[1]["ms"] * 1e-3) / 1e9
        frame_ms = total_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": frame_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": run_config(args.workload, n_frames),
            "sequence": "%d consecutive synthetic frames (camera motion, moving objects, disocclusions), cycled" % n_frames,
            "algorithmic_bytes_per_px": chain_bpp,
            "frame_ms": {k: round(v, 4) for k, v in frame_stats.items()},
            "clocks": clocks,
            "e2e": {"value": world * W * H * e2e_steps / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom[0], "launches_per_step": dom[1]["launches"],
                         "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                         "traffic": measured_traffic(dom[0], W, H, N),
                         "algorithmic_bytes_per_launch": dom[1]["bytes"] / dom[1]["launches"],
                         "peak_source": peak_src,
                         "chain": {"achieved": chain_bpp * W * H / (frame_ms * 1e-3) / 1e9,
                                   "frac": chain_bpp * W * H / (frame_ms * 1e-3) / 1e9 / peak,
                                   "frac_of_8000_nominal": chain_bpp * W * H / (frame_ms * 1e-3) / 1e9 / 8000.0}},
            "passes": per_pass,
            # the checkerboard resolve of the three image pairs (26 B/px algorithmic: 3 x (4 read + 4 written) + Throughput.a 2;
            # neighbours from cache), the final prepare (28 B/px: Final 4 + 4, ScreenEmission 4, Scattering 8, AcidFog 4, BloomInput 4), the exposure
            # histogram + adaptation (2 B/px) and the scattering accumulation (40 B/px of unique reads + writes: motion 8, depth 2,
            # normal 4, previous depth 2 / normal 4 / scattering 8 / history 2, 10 written): the reference's next dispatches,
            # timed separately (not part of the metric)
            "post": [{"kernel": k, "ms": round(v, 5),
                      "achieved_gbs": round({"k_checkerboard_resolve": 26.0, "k_scatter_accum": 40.0, "k_prepare_final": 28.0}.get(k, 2.0) * W * H / (v * 1e-3) / 1e9, 1)}
                     for k, v in post_ms.items()],
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(W, H, N, None)
        else:
            line["cpu_baseline"] = None
    # ---- the other BASELINE configurations, in the same line (explanatory legs; `value` above stays the headline metric) ----
    d.close()
    del resident, staging, host
    torch.cuda.empty_cache()
    legs = [] if args.legs == "none" else [x for x in args.legs.split(",") if x]
    gloo = None
    if dist is not None:
        gloo = dist.new_group(backend="gloo")         # host-side barrier: the idle ranks must not spin on their GPUs
    if "view_streams" in legs:                         # config 5: 48 x 1080p streams, 48 / N handles per GPU
        try:
            vs = leg_view_streams(rank, world, local_rank, dist)
        except Exception as e:  # noqa: BLE001
            vs = {"error": repr(e)}
        if rank == 0:
            line["view_streams"] = vs
    if "bands" in legs:                                # config 4: one 8K frame as bands over all the GPUs of this run
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=gloo)                   # every rank has released its GPU memory
        if rank == 0:
            try:
                line["bands"] = leg_bands(list(range(world)), "8k-bands", max(10, min(args.steps, 40)), max(3, min(args.warmup, 5)))
            except Exception as e:  # noqa: BLE001
                line["bands"] = {"error": repr(e)}
        if dist is not None:
            dist.barrier(group=gloo)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def leg_bands(devices, workload, steps, warmup, n_frames=4, log=None):
    """BASELINE config 4: ONE frame per step, split into horizontal bands across `devices` (asvgfBandSet*, include/asvgf.h):
    a single process drives every GPU, halos travel as peer copies behind the producing passes.  Returns the `bands`
    object of the JSON line.  Device-timed (events on every band's stream, the longest band counts)."""
    import numpy as np
    import torch

    import synth
    from raytracedgl1_b200 import bands as bands_mod
    from raytracedgl1_b200 import denoiser as dn
    from raytracedgl1_b200 import layout

    W, H, N, desc = WORKLOADS[workload]
    n = len(devices)
    swappable = {name for name in layout.PRODUCER_IMAGES if name + "_Prev" in layout.IMAGES}
    uniforms = [synth.uniform(W, H, f) for f in range(n_frames)]
    host = [synth.generate(W, H, f) for f in range(n_frames)]          # whole frames once; every band takes its rows

    def build(devs, boundaries, exchange=True):
        os.environ["ASVGF_BAND_NO_EXCHANGE"] = "0" if exchange else "1"
        bs = dn.BandSet(W, H, devs, N, boundaries=boundaries)
        resident = []
        for f in range(n_frames):
            per_band = []
            for k, b in enumerate(bs.bands):
                dev = {}
                for name in layout.PRODUCER_IMAGES:
                    _, pitch, _, r0, rows = b["denoiser"].image(name, 0)
                    a = host[f][name][r0:r0 + rows]
                    row_bytes = a.nbytes // rows
                    t = torch.zeros((rows, pitch), dtype=torch.uint8, device="cuda:%d" % b["device"])
                    t[:, :row_bytes].copy_(torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(rows, row_bytes)))
                    dev[name] = (t, pitch)
                per_band.append(dev)
            resident.append(per_band)
        for d in set(devs):
            torch.cuda.synchronize(d)
        return bs, resident

    def run(bs, resident, k_steps, w_steps):
        def step(i):
            f, fi = i % n_frames, i & 1
            for k, b in enumerate(bs.bands):
                for name in layout.PRODUCER_IMAGES:
                    t, pitch = resident[f][k][name]
                    b["denoiser"].bind_external(name + "_Prev" if (fi == 1 and name in swappable) else name, t.data_ptr(), pitch)
            bs.frame(fi, uniforms[f])
        for i in range(w_steps):
            step(i)
        bs.synchronize()
        bs.begin_timing()
        t0 = time.perf_counter()
        for i in range(k_steps):
            step(w_steps + i)
        ms, per = bs.end_timing()
        bs.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        return ms / k_steps, [p / k_steps for p in per], wall_ms / k_steps

    out = {"workload": desc.replace("with per-pass halo exchange over NCCL/NVLink", "halos pushed between neighbours as peer copies behind the producing pass"),
           "width": W, "height": H, "atrous_iterations": N, "n_gpus": len(set(devices)), "bands": n, "steps": steps}
    k_cal = max(4, steps // 4)
    # one band on the first device: the single-GPU time of the same frame
    bs, res = build(devices[:1], None)
    t1, _, _ = run(bs, res, max(4, steps // 2) if n > 1 else steps, warmup)
    bs.close(); del res
    out["ms_per_step_n1"] = t1
    if n == 1:
        out.update({"ms_per_step": t1, "value": W * H / (t1 * 1e-3) / 1e6, "unit": UNIT})
        return out
    # equal bands without exchange: every band's own kernel time -> boundaries of equal cost (rows are not equally expensive:
    # sky rows skip the taps), rounded to multiples of 3
    bs, res = build(devices, None, exchange=False)
    halo = bs.stored_rows(1, "Normal")[0]
    halo = bs.bands[1]["rows"][0] - halo
    _, per_equal, _ = run(bs, res, k_cal, warmup)
    equal_boundaries = [b["rows"][0] for b in bs.bands] + [H]
    bs.close(); del res
    boundaries = bands_mod.rebalanced_rows(equal_boundaries, per_equal, min_rows=halo)
    # the measured configuration, and the same without exchange (what the halo traffic and its waits cost)
    bs, res = build(devices, boundaries, exchange=False)
    t_compute, per_compute, _ = run(bs, res, k_cal, warmup)
    bs.close(); del res
    bs, res = build(devices, boundaries, exchange=True)
    tN, per, wall = run(bs, res, steps, warmup)
    bs.close(); del res
    out.update({"ms_per_step": tN, "value": W * H / (tN * 1e-3) / 1e6, "unit": UNIT, "speedup_vs_n1": t1 / tN,
                "efficiency_vs_n1": t1 / (tN * len(set(devices))), "per_band_ms": [round(x, 4) for x in per],
                "per_band_kernel_ms_equal_bands": [round(x, 4) for x in per_equal],
                "per_band_kernel_ms": [round(x, 4) for x in per_compute], "ms_per_step_without_exchange": t_compute,
                "exchange_exposed_ms": tN - t_compute, "exchange_exposed_frac": (tN - t_compute) / tN,
                "host_wall_ms_per_step": wall, "boundaries": boundaries, "halo_rows": halo,
                "scaling": "strong"})
    return out


def leg_view_streams(rank, world, local_rank, dist, steps_per_stream=64, total_streams=48, n_frames=4):
    """BASELINE config 5: 48 independent 1920x1080 view streams sharded over the GPUs (48 / N handles per GPU, each with its own
    CUDA stream and its own gradient stream), 64 frames each, no inter-GPU communication.  Returns the `view_streams`
    object of the JSON line (rank 0) -- device-timed, max over ranks."""
    import numpy as np
    import torch

    import synth
    from raytracedgl1_b200 import denoiser as dn
    from raytracedgl1_b200 import layout

    W, H, N = 1920, 1080, 4
    per_gpu = total_streams // world
    torch.cuda.set_device(local_rank)
    swappable = {name for name in layout.PRODUCER_IMAGES if name + "_Prev" in layout.IMAGES}
    handles = [dn.Denoiser(W, H, N, device=local_rank, flags=dn.CREATE_EAGER_GRADIENT_ATROUS) for _ in range(per_gpu)]
    streams = [(torch.cuda.Stream(), torch.cuda.Stream()) for _ in range(per_gpu)]
    uniforms = [synth.uniform(W, H, f) for f in range(n_frames)]
    resident = []                                  # the views share one synthetic sequence (read-only inputs)
    for f in range(n_frames):
        g = synth.generate(W, H, f, synth.DEFAULT_SEED + 100 + rank)
        dev = {}
        for name in layout.PRODUCER_IMAGES:
            _, pitch, _, _, rows = handles[0].image(name, 0)
            row_bytes = g[name].nbytes // rows
            t = torch.zeros((rows, pitch), dtype=torch.uint8, device="cuda")
            t[:, :row_bytes].copy_(torch.from_numpy(g[name].view(np.uint8).reshape(rows, row_bytes)))
            dev[name] = (t, pitch)
        resident.append(dev)
    torch.cuda.synchronize()

    def step(i):
        f, fi = i % n_frames, i & 1
        for d, (st, gst) in zip(handles, streams):
            for name in layout.PRODUCER_IMAGES:
                t, pitch = resident[f][name]
                d.bind_external(name + "_Prev" if (fi == 1 and name in swappable) else name, t.data_ptr(), pitch)
            d.gradient_samples(fi, uniforms[f], stream=gst.cuda_stream)
            d.denoise(fi, uniforms[f], stream=st.cuda_stream)

    main = torch.cuda.current_stream()
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st, gst in streams:
        st.wait_event(e0); gst.wait_event(e0)
    for i in range(steps_per_stream):
        step(4 + i)
    for st, gst in streams:
        main.wait_stream(st); main.wait_stream(gst)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    for d in handles:
        d.close()
    del resident
    torch.cuda.empty_cache()
    frames = per_gpu * world * steps_per_stream
    return {"workload": "48 independent 1920x1080 view streams, full chain (4 a-trous iterations), %d frames each, sharded over the GPUs, "
                        "no inter-GPU communication" % steps_per_stream,
            "streams": per_gpu * world, "streams_per_gpu": per_gpu, "frames_per_stream": steps_per_stream, "n_gpus": world,
            "ms_total": ms, "ms_per_frame_per_stream": ms / steps_per_stream / per_gpu, "value": frames * W * H / (ms * 1e-3) / 1e6, "unit": UNIT,
            "scaling": "strong (48 streams in total)"}


def run_bands(args, rank, world, local_rank):
    """--workload 8k-bands / 4k-bands: the band leg alone, as the line's metric.  One process drives --gpus devices."""
    if rank != 0:
        return
    import torch
    n = args.gpus
    if torch.cuda.device_count() < n:
        raise SystemExit("bench.py: --gpus %d but %d visible" % (n, torch.cuda.device_count()))
    b = leg_bands(list(range(n)), args.workload, args.steps, args.warmup, n_frames=max(2, min(args.frames, 4) // 2 * 2))
    W, H, N, desc = WORKLOADS[args.workload]
    line = {"metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": b["workload"], "width": W, "height": H, "atrous_iterations": N,
                                            "parallelism": "one frame as %d horizontal bands, one process, peer copies" % n},
            "bands": b}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frames", type=int, default=32, help="distinct synthetic frames in the cycled sequence (even)")
    ap.add_argument("--legs", default="bands,view_streams",
                    help="extra workloads reported in the same line: bands (BASELINE config 4: the 8K frame as bands over the "
                         "--gpus devices) and view_streams (config 5: 48 x 1080p streams); 'none' to skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the synthetic-input generator and the CPU oracle are OpenMP code and
    # would take minutes per rank: give every rank its share of the cores (set before anything loads libgomp)
    if world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // world))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload.endswith("-bands"):
        run_bands(args, rank, world, local_rank)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus %d needs torchrun (one process per GPU)" % args.gpus)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
