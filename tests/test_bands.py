"""Band decomposition of one frame across ranks (raytracedgl1_b200/bands.py): plan arithmetic, the pass/exchange
schedule, and the halo exchange itself over a 2-rank gloo group on CPU tensors.  The GPU leg (needs 2 GPUs) checks that
a frame denoised as two bands is bit-identical to the single-GPU frame."""
import os
import socket
import sys

import numpy as np
import pytest

from raytracedgl1_b200 import bands, layout


def test_band_rows_are_multiples_of_three_and_cover_the_frame():
    for H in (2160, 4320, 1080, 720, 181):
        for world in (1, 2, 4, 8):
            b = bands.band_rows(H, world)
            assert b[0] == 0 and b[-1] == H and len(b) == world + 1
            assert all(x % 3 == 0 for x in b[:-1])
            assert all(b[i + 1] > b[i] for i in range(world))
    with pytest.raises(ValueError):
        bands.band_rows(6, 8)


def test_plan_rows():
    p = bands.BandPlan(4320, 8, 3, halo_rows=48)
    assert p.own == (1620, 2160) and p.stored == (1572, 2208)
    assert p.gown == (540, 720) and p.gstored == (524, 736)
    assert p.up == 2 and p.down == 4
    top = bands.BandPlan(4320, 8, 0, halo_rows=48)
    assert top.stored[0] == 0 and top.up is None
    last = bands.BandPlan(2161 - 1, 4, 3, halo_rows=48)
    assert last.own[1] == 2160 and last.down is None and last.gown[1] == (2160 + 1) // 3
    with pytest.raises(ValueError):
        bands.BandPlan(4320, 8, 0, halo_rows=50)


def test_schedule_matches_the_dispatch_order_and_reaches():
    s = bands.frame_schedule(4)
    passes = [x[1] for x in s if x[0] == "pass"]
    assert passes == list(range(0, 12))                       # Denoiser.cpp:87-254 after CalculateGradientsSamples
    # every exchange directly precedes the pass that reads the halo, with that pass's reach
    ex = {s[i + 1][1]: s[i] for i in range(len(s) - 1) if s[i][0] == "exchange" and s[i + 1][0] == "pass"}
    assert ex[layout.PASS_GRADIENT_ATROUS_0 + 3][1:] == (["DISPongGradient"], 8, True)
    assert ex[layout.PASS_ANTIFIREFLY][2] == 2 and ex[layout.PASS_VARIANCE][2] == 1
    assert ex[layout.PASS_ATROUS_0 + 3][1:] == (["DiffPongColorAndVariance", "SpecPongColor", "IndirPong"], 8, False)
    assert ex[layout.PASS_ATROUS_0 + 1][1][0] == "DiffColorHistory"
    s5 = bands.frame_schedule(5)
    assert [x for x in s5 if x[0] == "exchange"][-1][2] == 16
    assert bands.physical_name("Normal", 1) == "Normal_Prev" and bands.physical_name("Normal_Prev", 1) == "Normal"
    assert bands.physical_name("Motion", 1) == "Motion" and bands.physical_name("IndirAccum_Prev", 0) == "IndirAccum_Prev"


def test_cost_weighted_band_rows():
    """Bands of equal summed cost instead of equal height (measured load balancing of bench.py --band-balance measured)."""
    H = 4320
    assert bands.band_rows(H, 4, np.ones(H)) == bands.band_rows(H, 4)
    cost = np.r_[np.full(H // 2, 3.0), np.full(H // 2, 3.9)]          # the 8K synthetic scene: bottom half 30 % dearer
    b = bands.band_rows(H, 2, cost)
    assert b[0] == 0 and b[2] == H and b[1] % 3 == 0 and b[1] > H // 2
    assert abs(cost[:b[1]].sum() - cost[b[1]:].sum()) <= 3 * cost.max() * 1.01
    assert bands.rebalanced_rows([0, 2160, 4320], [3.0, 3.9]) == b
    r = bands.rebalanced_rows([0, 1080, 2160, 3240, 4320], [0.1, 0.1, 5.0, 5.0], min_rows=66)
    assert all(r[k + 1] - r[k] >= 66 and r[k] % 3 == 0 for k in range(4)) and r[-1] == H
    p = bands.BandPlan(H, 2, 1, halo_rows=66, boundaries=b)
    assert p.own == (b[1], H) and p.stored == (b[1] - 66, H)
    with pytest.raises(ValueError):
        bands.BandPlan(H, 2, 0, halo_rows=66, boundaries=[0, 2000, H])   # 2000 is not a multiple of 3
    with pytest.raises(ValueError):
        bands.band_rows(H, 2, np.ones(H - 1))


def test_apron_extents_and_schedule():
    """ASVGF_CREATE_BAND_APRON: the rows each pass computes beyond the band are the summed reaches of the passes after
    it, and the apron schedule keeps only two exchanges: previous-frame state before, gradient samples after pass 0."""
    from raytracedgl1_b200 import denoiser as dn
    rows = {p: dn.band_apron_rows(4, p) for p in range(12)}
    assert [rows[p] for p in (11, 10, 9, 8)] == [0, 8, 12, 14]             # a-trous 3..0: +8, +4, +2 (CmSVGFAtrous.comp:61)
    assert (rows[layout.PASS_VARIANCE], rows[layout.PASS_ANTIFIREFLY], rows[layout.PASS_TEMPORAL]) == (15, 16, 18)
    assert [rows[p] for p in (4, 3, 2, 1)] == [6, 14, 18, 20]              # gradient a-trous 3..0, gradient rows (18/3, +8, +4, +2)
    assert rows[layout.PASS_GRADIENT_SAMPLES] == 21
    assert dn.band_apron_rows(5, layout.PASS_TEMPORAL) == 34 and dn.band_apron_rows(1, 0) == -1
    rows_of = lambda p: dn.band_apron_rows(4, p)  # noqa: E731
    s = bands.frame_schedule_apron(4, rows_of)
    assert [x[1] for x in s if x[0] == "pass"] == list(range(12))
    ex = [i for i, x in enumerate(s) if x[0] == "exchange"]
    assert ex == [0, 2] and s[2][1] == [(["DISPingGradient"], 21, True)]
    assert bands.apron_halo_rows(4, rows_of) == 66 and bands.apron_halo_rows(5, lambda p: dn.band_apron_rows(5, p)) % 3 == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, H, pitch, halo, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        full = rng.integers(0, 256, size=(H, pitch), dtype=np.uint8)            # the frame every rank agrees on
        gfull = rng.integers(0, 256, size=((H + 1) // 3, pitch), dtype=np.uint8)
        plan = bands.BandPlan(H, world, rank, halo)
        ok = True
        for third, ref in ((False, full), (True, gfull)):
            (o0, o1), (s0, s1) = plan.rows(third)
            for depth in (1, 2, 8, halo // 3 if third else halo):
                t = torch.zeros((s1 - s0, pitch), dtype=torch.uint8)
                t[o0 - s0:o1 - s0] = torch.from_numpy(ref[o0:o1].copy())      # only the owned rows are known
                bands.exchange_halos([t, t.clone()][:1], depth, plan, third)
                lo, hi = max(s0, o0 - depth), min(s1, o1 + depth)
                ok &= bool(np.array_equal(t.numpy()[lo - s0:hi - s0], ref[lo:hi]))
                # rows beyond the exchanged depth stay untouched (zero)
                if lo > s0:
                    ok &= bool((t.numpy()[:lo - s0] == 0).all())
        # the apron schedule's frame-start exchange: full-resolution and gradient-resolution images in ONE batch,
        # with unequal (cost-balanced) bands
        b = [0] + [min(H, 3 * ((k * H // world + 20 * (k % 2)) // 3)) for k in range(1, world)] + [H]
        plan2 = bands.BandPlan(H, world, rank, halo, boundaries=b)
        specs, refs = [], []
        for third, ref, depth in ((False, full, halo), (True, gfull, 7)):
            (o0, o1), (s0, s1) = plan2.rows(third)
            t = torch.zeros((s1 - s0, pitch), dtype=torch.uint8)
            t[o0 - s0:o1 - s0] = torch.from_numpy(ref[o0:o1].copy())
            specs.append(([t], depth, third))
            refs.append((t, ref, depth, third))
        bands.exchange_halos_multi(specs, plan2)
        for t, ref, depth, third in refs:
            (o0, o1), (s0, s1) = plan2.rows(third)
            lo, hi = max(s0, o0 - depth), min(s1, o1 + depth)
            ok &= bool(np.array_equal(t.numpy()[lo - s0:hi - s0], ref[lo:hi]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, 360, 64, 48, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _nccl_worker(rank, world, port, W, H, frames, apron, boundaries, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        band = bands.BandedDenoiser(W, H, 4, rank=rank, world=world, device=rank, apron=apron, boundaries=boundaries)
        whole = bands.BandedDenoiser(W, H, 4, rank=0, world=1, device=rank) if rank == 0 else None
        ok, detail = True, []
        for f in range(frames):
            fi = f & 1
            g = synth.generate(W, H, f)
            u = synth.uniform(W, H, f)
            for name, arr in g.items():
                band.upload(name, fi, arr)
                if whole:
                    whole.upload(name, fi, arr)
            band.frame(fi, u)
            if whole:
                whole.frame(fi, u)
            torch.cuda.synchronize()
            for name in ("PreFinal", "HistogramInput", "DiffColorHistory", "AccumHistoryLength", "SpecAccumColor", "IndirAccum",
                         "DISPingGradient", "GradientPrevPix", "AtrousFilteredVariance"):
                mine = torch.from_numpy(band.download_own(name, fi).copy()).cuda().view(torch.uint8).flatten()
                parts = [torch.empty_like(mine) for _ in range(world)] if False else None
                sizes = torch.tensor([mine.numel()], device="cuda")
                all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
                dist.all_gather(all_sizes, sizes)
                mx = int(max(int(s) for s in all_sizes))
                pad = torch.zeros(mx, dtype=torch.uint8, device="cuda")
                pad[:mine.numel()] = mine
                parts = [torch.zeros_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad)
                if rank == 0:
                    joined = np.concatenate([p[:int(s)].cpu().numpy() for p, s in zip(parts, all_sizes)])
                    ref = whole.download_own(name, fi).view(np.uint8).flatten()
                    same = bool(np.array_equal(joined, ref))
                    ok &= same
                    if not same:
                        detail.append((f, name, int((joined != ref).sum())))
        q.put((rank, ok, detail))
    finally:
        dist.destroy_process_group()


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# NCCL refuses two ranks on one device, so the two-process cases only exist on a box with two GPUs (gpurun --gpus 2).  On one GPU
# the same schedule -- apron passes, push points, unequal bands -- is exercised by tests/test_bandset_gpu.py (several bands of a band
# set on cuda:0) and, for the torch.distributed plumbing, by the world_size-2 gloo tests above.
if _gpu_count() >= 2:
    @pytest.mark.gpu
    @pytest.mark.parametrize("apron,boundaries", [(True, None), (False, None), (True, [0, 150, 360]), (False, [0, 231, 360])])
    def test_two_bands_bit_identical_to_single_gpu(apron, boundaries):
        import torch.multiprocessing as mp
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, 640, 360, 4, apron, boundaries, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=600) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert all(r[1] for r in res), res


@pytest.mark.gpu
def test_banded_denoiser_single_rank_equals_plain_denoiser():
    """world = 1: the per-pass schedule over externally bound (torch) images must give exactly what
    asvgfGradientSamples + asvgfDenoise give on library-owned images."""
    import synth
    from raytracedgl1_b200 import denoiser as dn
    W, H = 192, 108
    b = bands.BandedDenoiser(W, H, 4, rank=0, world=1, device=0)
    d = dn.Denoiser(W, H, 4)
    try:
        for f in range(3):
            fi = f & 1
            g = synth.generate(W, H, f)
            u = synth.uniform(W, H, f)
            for name, arr in g.items():
                b.upload(name, fi, arr)
                d.upload(name, fi, arr)
            b.frame(fi, u)
            d.gradient_samples(fi, u)
            d.denoise(fi, u)
            d.synchronize()
            for name in ("PreFinal", "HistogramInput", "DiffColorHistory", "GradientPrevPix", "DISPingGradient", "SpecAccumColor"):
                assert np.array_equal(b.download_own(name, fi), d.download(name, fi)), (f, name)
    finally:
        b.close(); d.close()
