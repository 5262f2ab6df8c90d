#!/usr/bin/env python
"""Feasibility measurement (TIMING ONLY, results are not checked): how much throughput would overlapping the last a-trous
iterations of frame t with the temporal / anti-firefly / variance passes of frame t+1 buy?  Passes are enqueued through
asvgfDenoisePass on two streams with the ordering the real pipeline would need (A1(t) after A0(t); F(t+1) after A2(t);
A0(t+1) after A3(t)); the guide plane is NOT double-buffered here, so the images are wrong -- only the clock matters.

Result on B200 (round 2, 4K, 8-frame cycle): serial 1.524 ms/frame, overlapped 1.665 ms/frame -- co-resident kernels of two frames
are 9 % SLOWER than the strict chain (every pass already fills the machine; sharing the SMs costs L1/L2 locality and buys no
issue slots).  Not adopted: asvgfDenoise keeps one stream per frame chain (only the latency-bound gradient chain overlaps)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth
from raytracedgl1_b200 import denoiser as dn, layout

W, H, N = 3840, 2160, 4
d = dn.Denoiser(W, H, N, flags=dn.CREATE_EAGER_GRADIENT_ATROUS)
n_frames = 8
swappable = {n for n in layout.PRODUCER_IMAGES if n + "_Prev" in layout.IMAGES}
resident, uniforms = [], []
for f in range(n_frames):
    g = synth.generate(W, H, f)
    dev = {}
    for name in layout.PRODUCER_IMAGES:
        _, pitch, _, _, rows = d.image(name, 0)
        rb = g[name].nbytes // rows
        t = torch.zeros((rows, pitch), dtype=torch.uint8, device="cuda")
        t[:, :rb].copy_(torch.from_numpy(g[name].view(np.uint8).reshape(rows, rb)))
        dev[name] = (t, pitch)
    resident.append(dev); uniforms.append(synth.uniform(W, H, f))
torch.cuda.synchronize()
s_main, s_tail, s_grad = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

def bind(i):
    f, fi = i % n_frames, i & 1
    for name in layout.PRODUCER_IMAGES:
        t, pitch = resident[f][name]
        d.bind_external(name + "_Prev" if (fi == 1 and name in swappable) else name, t.data_ptr(), pitch)
    return f, fi

def serial(i):
    f, fi = bind(i)
    d.gradient_samples(fi, uniforms[f], stream=s_grad.cuda_stream)
    d.denoise(fi, uniforms[f], stream=s_main.cuda_stream)

ev = {}
def overlapped(i):
    f, fi = bind(i)
    u = uniforms[f]
    d.gradient_samples(fi, u, stream=s_grad.cuda_stream)
    m, t = s_main.cuda_stream, s_tail.cuda_stream
    d.denoise_pass(layout.PASS_GRADIENT_ATROUS_ALL, fi, u, stream=m)
    d.denoise_pass(layout.PASS_TEMPORAL, fi, u, stream=m)
    if "a2" in ev: s_main.wait_event(ev["a2"])
    d.denoise_pass(layout.PASS_ANTIFIREFLY, fi, u, stream=m)
    d.denoise_pass(layout.PASS_VARIANCE, fi, u, stream=m)
    if "a3" in ev: s_main.wait_event(ev["a3"])
    d.denoise_pass(layout.PASS_ATROUS_0, fi, u, stream=m)
    e0 = torch.cuda.Event(); e0.record(s_main); s_tail.wait_event(e0)
    d.denoise_pass(layout.PASS_ATROUS_0 + 1, fi, u, stream=t)
    d.denoise_pass(layout.PASS_ATROUS_0 + 2, fi, u, stream=t)
    ev["a2"] = torch.cuda.Event(); ev["a2"].record(s_tail)
    d.denoise_pass(layout.PASS_ATROUS_0 + 3, fi, u, stream=t)
    ev["a3"] = torch.cuda.Event(); ev["a3"].record(s_tail)

def run(step, k=80, w=16):
    for i in range(w): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    b.record(s_main)
    for i in range(w, w + k): step(i)
    s_main.wait_stream(s_tail); e.record(s_main)
    torch.cuda.synchronize()
    return b.elapsed_time(e) / k, (time.perf_counter() - t0) * 1e3 / k

for name, fn in (("serial", serial), ("overlapped", overlapped), ("serial", serial), ("overlapped", overlapped)):
    ms, wall = run(fn)
    print("%-10s %.4f ms/frame (device), %.4f ms/frame (host wall)" % (name, ms, wall), flush=True)
d.close()
