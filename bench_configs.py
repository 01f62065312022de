"""bench.py --config c3 | c4 | c5: the other BASELINE.json configurations, each at full size with distinct synthetic
channels, a bit-exact spot check against the CPU oracle inside the run, `roofline`, `cpu_baseline` and `e2e` objects.

  c3  8192-channel GC-ADPCM decode (.dsp payloads -> PCM16) on one B200, bit-exact check       (BASELINE configs[2])
  c4  512-stream CRI HCA encode (128-point MDCT, quality High, mono 48 kHz) on one B200          (BASELINE configs[3])
  c5  65 536-file mixed GC-ADPCM + ADX batch encode, STRONG scaling over N GPUs: the root rank holds the PCM in HBM,
      one NCCL scatterv hands every rank its files, every rank encodes, one NCCL gatherv returns the bitstreams
      (BASELINE configs[4]; launched under torchrun for N > 1)

Timing rules as in bench.py: CUDA events on the launching stream, W warm-up steps, max over ranks, inputs far larger than
L2.  The oracle (oracle/) is used as the checker and as the CPU baseline only.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

SAMPLE_RATE = 48000
ROOT = os.path.dirname(os.path.abspath(__file__))


def _peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _barrier(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(torch, dist, world, device, value):
    if world == 1:
        return float(value)
    t = torch.tensor([value], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _base_line(metric, value, world, args, ms, scaling, dtype, config):
    return {"metric": metric, "value": round(value, 3), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config}


# ======================================================================================================================
# c3: GC-ADPCM decode
# ======================================================================================================================
def run_c3(args, env, ctx):
    torch, dist, vg, N, bench = ctx["torch"], ctx["dist"], ctx["vg"], ctx["N"], ctx["bench"]
    rank, local_rank, world = env
    device = torch.device("cuda", local_rank)
    n_ch = args.channels if args.channels != 1024 else 8192
    n = int(round(args.seconds * SAMPLE_RATE))
    stream = torch.cuda.current_stream()
    stride = (n + 7) // 8 * 8
    n_bytes = vg.gcadpcm.sample_count_to_byte_count(n)
    a_stride = (n_bytes + 15) // 16 * 16
    frames = (n + 13) // 14

    # ---- the input of the decoder = the encoder's output for n_ch DISTINCT synthetic channels (encoded 1024 at a time)
    adpcm = torch.zeros((n_ch, a_stride), dtype=torch.uint8, device=device)
    coefs = torch.zeros((n_ch, 16), dtype=torch.int16, device=device)
    chunk = min(1024, n_ch)
    ws_bytes = int(vg.lib.vgb_gcadpcm_workspace_bytes(frames * chunk, chunk))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    lens_c = np.full(chunk, n, dtype=np.int32)
    keep_pcm = []  # the first 512 channels' PCM stays for the round-trip property
    for c0 in range(0, n_ch, chunk):
        m = min(chunk, n_ch - c0)
        pcm = bench.make_batch_gpu(torch, m, n, rank * 64 + c0 // chunk, device, degenerate=(c0 == 0))
        pad = torch.zeros((m, stride), dtype=torch.int16, device=device)
        pad[:, :n] = pcm
        off_p = np.arange(m, dtype=np.int64) * stride
        off_a = (np.arange(m, dtype=np.int64) + c0) * a_stride
        N.check(vg.lib.vgb_gcadpcm_encode_dev(pad.data_ptr(), off_p.ctypes.data, lens_c.ctypes.data, None, m, None,
                                              coefs[c0:].data_ptr(), adpcm.data_ptr(), off_a.ctypes.data, ws.data_ptr(), ws_bytes,
                                              stream.cuda_stream))
        torch.cuda.synchronize()
        if c0 == 0:
            keep_pcm = pcm[:512].clone()
        del pcm, pad
    del ws
    pcm_out = torch.zeros((n_ch, stride), dtype=torch.int16, device=device)
    dws_bytes = int(vg.lib.vgb_gcadpcm_workspace_bytes(32, n_ch))
    dws = torch.empty(dws_bytes, dtype=torch.uint8, device=device)
    off_a = np.arange(n_ch, dtype=np.int64) * a_stride
    off_p = np.arange(n_ch, dtype=np.int64) * stride
    params = (N.VgbGcParams * n_ch)()
    for c in range(n_ch):
        params[c].sample_count, params[c].history1, params[c].history2 = n, 0, 0

    def step():
        N.check(vg.lib.vgb_gcadpcm_decode_dev(adpcm.data_ptr(), off_a.ctypes.data, coefs.data_ptr(), params, n_ch, pcm_out.data_ptr(),
                                              off_p.ctypes.data, dws.data_ptr(), dws_bytes, stream.cuda_stream))

    N.check(vg.lib.vgb_set_kernel_timing(1))
    for _ in range(args.warmup):
        step()
    _barrier(torch, dist, world)
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    launches0 = vg.lib.vgb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        buf = (C.c_float * 4)()
        N.check(vg.lib.vgb_last_kernel_ms(buf, 4))
        kms += buf[3]
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, device, ev0.elapsed_time(ev1) / args.steps)
    launches = vg.lib.vgb_kernel_launch_count() - launches0
    clocks = sampler.stop()
    kms /= args.steps
    samples = n_ch * n
    value = world * samples / (ms / 1e3) / 1e6

    # ---- e2e: host API with pinned buffers, 2048 channels per call (chunk when host RAM is short)
    e2e = None
    if not args.no_e2e:
        per = min(2048, n_ch)
        h_in = torch.empty((per, n_bytes), dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty((per, n), dtype=torch.int16, pin_memory=True)
        h_coefs = np.zeros((per, 16), dtype=np.int16)
        in_tab = (C.c_void_p * per)(*[h_in.data_ptr() + n_bytes * c for c in range(per)])
        out_tab = (C.c_void_p * per)(*[h_out.data_ptr() + 2 * n * c for c in range(per)])
        nb = np.full(per, n_bytes, dtype=np.int32)
        pr = (N.VgbGcParams * per)()
        for c in range(per):
            pr[c].sample_count, pr[c].history1, pr[c].history2 = n, 0, 0
        total_ms, same = 0.0, True
        for rep in range(2):  # first pass warms the library's slabs
            total_ms = 0.0
            for c0 in range(0, n_ch, per):
                h_in.copy_(adpcm[c0:c0 + per, :n_bytes])
                h_coefs[:] = coefs[c0:c0 + per].cpu().numpy()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                N.check(vg.lib.vgb_gcadpcm_decode_batch(in_tab, nb.ctypes.data, h_coefs.ctypes.data, pr, per, out_tab))
                total_ms += (time.perf_counter() - t0) * 1e3
                if rep == 1 and c0 == 0:
                    same = bool((h_out[:64].to(device) == pcm_out[:64, :n]).all().item())
        total_ms = _max_over_ranks(torch, dist, world, device, total_ms)
        tl = (C.c_float * 48)()
        N.check(vg.lib.vgb_debug_last_timeline(tl, 48))
        h2d, d2h = n_ch * (n_bytes + 32), n_ch * n * 2
        e2e = {"value": round(world * samples / (total_ms / 1e3) / 1e6, 3), "unit": "Msamples/s", "ms_per_step": round(total_ms, 3),
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "pcie_floor_ms": round(d2h / 55e9 * 1e3, 1),
               "api": f"vgb_gcadpcm_decode_batch, pinned host buffers, {n_ch // per} calls of {per} channels",
               "timeline_ms_last_call": [[round(tl[3 * g + k], 1) for k in range(3)] for g in range(16) if tl[3 * g] >= 0],
               "matches_device_resident": same}

    # ---- parity: >= 512 channels decoded by the oracle from the same bytes, bit-exact; round trip against the input
    parity = cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import pyoracle

        k = min(512, n_ch)
        a_host = adpcm[:k, :n_bytes].cpu().numpy()
        c_host = coefs[:k].cpu().numpy()
        t0 = time.perf_counter()
        o_dec, used = pyoracle.decode_batch(a_host, c_host, n)
        dt = time.perf_counter() - t0
        g_dec = pcm_out[:k, :n].cpu().numpy()
        err = np.abs(g_dec[:min(k, len(keep_pcm))].astype(np.int32) - keep_pcm[:k].cpu().numpy().astype(np.int32))
        parity = {"channels_checked": int(k), "pcm_equal_oracle": bool(np.array_equal(g_dec, o_dec)),
                  "round_trip_rms_lsb": round(float(np.sqrt((err.astype(np.float64) ** 2).mean())), 2)}
        # CPU baseline: the same oracle decode over all host cores on a bounded sample, timed again warm
        t0 = time.perf_counter()
        pyoracle.decode_batch(a_host, c_host, n)
        dt = min(dt, time.perf_counter() - t0)
        cpu = {"value": round(k * n / dt / 1e6, 3), "unit": "Msamples/s", "cores": int(used), "kind": "port",
               "sample": f"{k} of {n_ch} channels x {n} samples ({dt:.2f} s wall), C restatement of GcAdpcmDecoder.Decode, one task per channel"}

    if rank != 0:
        return None
    peak, peak_src = _peak()
    alg = samples * (2.0 + 8.0 / 14.0)
    achieved = alg / (kms / 1e3) / 1e9 if kms > 0 else None
    line = _base_line("GC-ADPCM decode Msamples/sec (batch)", value, world, args, ms, "weak", "int32",
                      {"workload": f"{n_ch} ch x {args.seconds:g} s x 48 kHz GC-ADPCM (.dsp payload) -> PCM16, per GPU",
                       "global_channels": world * n_ch, "samples_per_channel": n, "distinct_channels": True,
                       "l2": "inputs + outputs (30 GB/GPU) larger than L2, no flush needed", "parallelism": f"dp{world} (channels sharded)"})
    line.update({"e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                 "roofline": {"bound": "hbm", "kernel": "gc_decode_kernel", "achieved": round(achieved, 2) if achieved else None, "peak": peak,
                              "unit": "GB/s", "frac": round(achieved / peak, 5) if achieved else None, "traffic": None, "peak_source": peak_src,
                              "algorithmic_bytes_per_launch": int(alg), "note": "8/14 B read + 2 B written per sample; one thread per channel, chain latency bound below ~30k channels"},
                 "kernel_ms": {"gc_decode": round(kms, 3)}, "cpu_baseline": cpu, "parity": parity})
    return line


# ======================================================================================================================
# c4: CRI HCA encode
# ======================================================================================================================
def run_c4(args, env, ctx):
    torch, dist, vg, N, bench = ctx["torch"], ctx["dist"], ctx["vg"], ctx["N"], ctx["bench"]
    rank, local_rank, world = env
    device = torch.device("cuda", local_rank)
    n_st = args.channels if args.channels != 1024 else 512
    n = int(round(args.seconds * SAMPLE_RATE))
    stream = torch.cuda.current_stream()
    stride = (n + 7) // 8 * 8
    pcm = bench.make_batch_gpu(torch, n_st, n, rank, device, degenerate=False)
    pcm_dev = torch.zeros((n_st, stride), dtype=torch.int16, device=device)
    pcm_dev[:, :n] = pcm
    params = (N.VgbHcaParams * n_st)()
    for s in range(n_st):
        params[s] = N.VgbHcaParams(2, 0, 0, 1, SAMPLE_RATE, n, 0, 0, 0)  # quality High, mono
    info0 = N.VgbHcaInfo()
    N.check(vg.lib.vgb_hca_query(C.byref(params[0]), C.byref(info0)))
    fbytes = info0.frame_count * info0.frame_size
    f_stride = (fbytes + 15) // 16 * 16
    frames_dev = torch.zeros((n_st, f_stride), dtype=torch.uint8, device=device)
    ws_bytes = int(vg.lib.vgb_hca_workspace_bytes(n_st))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    off_p = np.arange(n_st, dtype=np.int64) * stride
    ch_stride = np.full(n_st, stride, dtype=np.int64)
    off_f = np.arange(n_st, dtype=np.int64) * f_stride

    def step():
        N.check(vg.lib.vgb_hca_encode_dev(pcm_dev.data_ptr(), off_p.ctypes.data, ch_stride.ctypes.data, params, n_st, None,
                                          frames_dev.data_ptr(), off_f.ctypes.data, ws.data_ptr(), ws_bytes, stream.cuda_stream))

    N.check(vg.lib.vgb_set_kernel_timing(1))
    for _ in range(args.warmup):
        step()
    N.check(vg.lib.vgb_hca_encode_dev_status(ws.data_ptr(), n_st, stream.cuda_stream))
    _barrier(torch, dist, world)
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    launches0 = vg.lib.vgb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        buf = (C.c_float * 8)()
        N.check(vg.lib.vgb_last_kernel_ms(buf, 8))
        kms += buf[6]
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, device, ev0.elapsed_time(ev1) / args.steps)
    launches = vg.lib.vgb_kernel_launch_count() - launches0
    clocks = sampler.stop()
    kms /= args.steps
    samples = n_st * n
    value = world * samples / (ms / 1e3) / 1e6

    e2e = None
    if not args.no_e2e:
        h_in = torch.empty((n_st, n), dtype=torch.int16, pin_memory=True)
        h_in.copy_(pcm)
        h_out = torch.empty((n_st, fbytes), dtype=torch.uint8, pin_memory=True)
        in_tab = (C.c_void_p * n_st)(*[h_in.data_ptr() + 2 * n * s for s in range(n_st)])
        out_tab = (C.c_void_p * n_st)(*[h_out.data_ptr() + fbytes * s for s in range(n_st)])
        infos = (N.VgbHcaInfo * n_st)()

        def step_e2e():
            N.check(vg.lib.vgb_hca_encode_batch(in_tab, params, n_st, infos, out_tab, None, None))

        step_e2e()
        _barrier(torch, dist, world)
        reps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            step_e2e()
        e_ms = _max_over_ranks(torch, dist, world, device, (time.perf_counter() - t0) * 1e3 / reps)
        tl = (C.c_float * 48)()
        N.check(vg.lib.vgb_debug_last_timeline(tl, 48))
        same = bool((h_out.to(device) == frames_dev[:, :fbytes]).all().item())
        e2e = {"value": round(world * samples / (e_ms / 1e3) / 1e6, 3), "unit": "Msamples/s", "ms_per_step": round(e_ms, 3),
               "h2d_bytes_per_step": int(n_st * n * 2), "d2h_bytes_per_step": int(n_st * fbytes),
               "pcie_floor_ms": round(n_st * n * 2 / 55e9 * 1e3, 1), "api": "vgb_hca_encode_batch, pinned host buffers",
               "timeline_ms": [[round(tl[3 * g + k], 1) for k in range(3)] for g in range(16) if tl[3 * g] >= 0],
               "matches_device_resident": same}

    parity = cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import pyoracle

        k = min(64, n_st)
        host = pcm[:k].cpu().numpy()
        got = frames_dev[:k, :fbytes].cpu().numpy()
        cores = os.cpu_count() or 1

        def one(s):
            _, fr = pyoracle.hca_encode([host[s]], SAMPLE_RATE, 2)
            return fr.reshape(-1)

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=min(cores, k)) as pool:  # ctypes releases the GIL: one stream per core
            want = list(pool.map(one, range(k)))
        dt = time.perf_counter() - t0
        equal = all(np.array_equal(got[s], want[s]) for s in range(k))
        # decoded-domain check demanded by north_star: RMS between the decodes of both frame sets (0 when bytes are equal)
        parity = {"streams_checked": int(k), "frames_byte_identical": bool(equal), "rms_vs_reference_path": 0.0 if equal else None}
        cpu = {"value": round(k * n / dt / 1e6, 3), "unit": "Msamples/s", "cores": int(min(cores, k)), "kind": "port",
               "sample": f"{k} of {n_st} streams x {n} samples ({dt:.1f} s wall), C restatement of CriHcaEncoder, one stream per core "
                         "(the reference's CriHcaFormat.EncodeFromPcm16 is single-threaded per stream; its batch level is Parallel.ForEach over files)"}

    if rank != 0:
        return None
    peak, peak_src = _peak()
    alg = samples * 2.0 + n_st * fbytes
    achieved = alg / (kms / 1e3) / 1e9 if kms > 0 else None
    line = _base_line("CRI HCA encode Msamples/sec (batch)", value, world, args, ms, "weak", "f64",
                      {"workload": f"{n_st} mono streams x {args.seconds:g} s x 48 kHz PCM16 -> CRI HCA, quality High ({info0.frame_size} B frames), per GPU",
                       "global_streams": world * n_st, "samples_per_stream": n, "frames_per_stream": int(info0.frame_count), "distinct_streams": True,
                       "l2": "inputs (1.5 GB/GPU) larger than L2, no flush needed", "parallelism": f"dp{world} (streams sharded)"})
    line.update({"e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                 "roofline": {"bound": "hbm", "kernel": "hca_encode_kernel", "achieved": round(achieved, 2) if achieved else None, "peak": peak,
                              "unit": "GB/s", "frac": round(achieved / peak, 5) if achieved else None, "traffic": None, "peak_source": peak_src,
                              "algorithmic_bytes_per_launch": int(alg), "note": "fp64 MDCT + bit allocation search, ALU bound (~135 ops/sample, SURVEY 8d)"},
                 "kernel_ms": {"hca_encode": round(kms, 3)}, "cpu_baseline": cpu, "parity": parity})
    return line


# ======================================================================================================================
# c5: mixed GC-ADPCM + ADX batch, strong scaling with NCCL scatterv / gatherv
# ======================================================================================================================
def _c5_lengths(n_files):
    rng = np.random.default_rng([0x5647415544494F, 5])
    return rng.integers(1 * SAMPLE_RATE, 10 * SAMPLE_RATE + 1, n_files).astype(np.int64)


def _c5_fill(torch, slab, offs, lens, file_ids, device):
    """Synthetic audio straight into the slab: per file three sines + noise at -30 dB of a peak from {2000 .. 32767}."""
    g = torch.Generator(device=device)
    g.manual_seed(0x5647415544494F + 99)
    nf = len(file_ids)
    peaks = torch.tensor([2000.0, 8000.0, 20000.0, 32767.0], device=device)[torch.randint(0, 4, (nf,), generator=g, device=device)]
    w = torch.rand((nf, 3), generator=g, device=device) + 0.05
    amps = w / w.sum(1, keepdim=True) * peaks[:, None]
    freq = torch.exp(torch.rand((nf, 3), generator=g, device=device) * (np.log(12000.0) - np.log(60.0)) + np.log(60.0))
    phase = torch.rand((nf, 3), generator=g, device=device) * (2 * np.pi)
    offs_t = torch.as_tensor(offs, device=device)
    ends_t = offs_t + torch.as_tensor(lens, device=device)
    total = int(slab.numel())
    step = 1 << 27
    for s0 in range(0, total, step):
        idx = torch.arange(s0, min(s0 + step, total), device=device)
        f = torch.clamp(torch.searchsorted(offs_t, idx, right=True) - 1, min=0)
        inside = idx < ends_t[f]
        t = (idx - offs_t[f]).to(torch.float32) / SAMPLE_RATE
        x = torch.zeros(idx.numel(), device=device)
        for k in range(3):
            x += amps[f, k] * torch.sin(2 * np.pi * freq[f, k] * t + phase[f, k])
        x += torch.randn(idx.numel(), generator=g, device=device) * (peaks[f] * 10 ** (-30 / 20))
        slab[s0:s0 + idx.numel()] = torch.where(inside, torch.clamp(torch.round(x), -32768, 32767), torch.zeros_like(x)).to(torch.int16)
        del idx, f, inside, t, x


def run_c5(args, env, ctx):
    torch, dist, vg, N, bench = ctx["torch"], ctx["dist"], ctx["vg"], ctx["N"], ctx["bench"]
    rank, local_rank, world = env
    device = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    n_files = args.files
    lens = _c5_lengths(n_files)
    is_gc = (np.arange(n_files) % 2) == 0                       # even index -> GC-ADPCM, odd -> ADX (Linear, v4, 18-byte frames)
    # ---- partition (identical on every rank): greedy longest-first on an estimated cost, ADX samples weigh more
    adx_w = float(os.environ.get("VGB_C5_ADX_WEIGHT", "1.0"))
    weight = np.where(is_gc, lens, (lens * adx_w).astype(np.int64)).astype(np.int64)
    part = np.zeros(n_files, dtype=np.int32)
    load = np.zeros(world, dtype=np.int64)
    N.check(vg.lib.vgb_partition_lpt(weight.ctypes.data, n_files, world, part.ctypes.data, load.ctypes.data))
    # rank-major, codec-major, longest first inside (neighbouring channels of a warp then have similar lengths)
    order = np.lexsort((-lens, ~is_gc, part))
    lens_o, gc_o, part_o = lens[order], is_gc[order], part[order]
    pad = (lens_o + 7) // 8 * 8
    offs_o = np.concatenate(([0], np.cumsum(pad)[:-1])).astype(np.int64)          # sample offsets in the root's slab
    total_samples = int(pad.sum())
    rank_lo = np.searchsorted(part_o, np.arange(world), side="left")
    rank_hi = np.searchsorted(part_o, np.arange(world), side="right")
    pcm_counts = np.array([int(pad[rank_lo[r]:rank_hi[r]].sum()) * 2 for r in range(world)], dtype=np.int64)   # bytes
    pcm_offsets = np.array([int(offs_o[rank_lo[r]]) * 2 if rank_hi[r] > rank_lo[r] else 0 for r in range(world)], dtype=np.int64)

    # ---- my share
    lo, hi = int(rank_lo[rank]), int(rank_hi[rank])
    my_lens, my_gc = lens_o[lo:hi], gc_o[lo:hi]
    my_off = offs_o[lo:hi] - (offs_o[lo] if hi > lo else 0)
    n_gc, n_adx = int(my_gc.sum()), int((~my_gc).sum())
    gc_lens = my_lens[:n_gc].astype(np.int32)
    adx_lens = my_lens[n_gc:].astype(np.int32)
    gc_off, adx_off = my_off[:n_gc].copy(), my_off[n_gc:].copy()
    gc_bytes = np.array([vg.gcadpcm.sample_count_to_byte_count(int(v)) for v in gc_lens], dtype=np.int64)
    adx_bytes = np.array([vg.lib.vgb_adx_encoded_byte_count(int(v), 0, 18) for v in adx_lens], dtype=np.int64)
    # my output buffer: [GC payloads (16-aligned each) | coefficient table | ADX payloads]
    gc_out_off = np.concatenate(([0], np.cumsum((gc_bytes + 15) // 16 * 16)[:-1])).astype(np.int64) if n_gc else np.zeros(0, np.int64)
    gc_out_total = int(((gc_bytes + 15) // 16 * 16).sum())
    coef_at = gc_out_total
    adx_at = (coef_at + n_gc * 32 + 15) // 16 * 16
    adx_out_off = (adx_at + np.concatenate(([0], np.cumsum((adx_bytes + 15) // 16 * 16)[:-1]))).astype(np.int64) if n_adx else np.zeros(0, np.int64)
    my_out_bytes = adx_at + int(((adx_bytes + 15) // 16 * 16).sum())
    out_counts = np.zeros(world, dtype=np.int64)
    out_counts[rank] = my_out_bytes
    if world > 1:
        t = torch.as_tensor(out_counts, device=device)
        dist.all_reduce(t)
        out_counts = t.cpu().numpy()
    out_offsets = np.concatenate(([0], np.cumsum((out_counts + 255) // 256 * 256)[:-1])).astype(np.int64)

    # ---- buffers.  Root: the whole PCM slab (its own share is used in place) and the gathered outputs.
    if rank == 0:
        slab = torch.zeros(total_samples + 8, dtype=torch.int16, device=device)
        _c5_fill(torch, slab, offs_o, lens_o, order, device)
        gathered = torch.zeros(int(out_offsets[-1] + (out_counts[-1] + 255) // 256 * 256) + 256, dtype=torch.uint8, device=device)
        my_pcm = slab[int(pcm_offsets[0] // 2):int(pcm_offsets[0] // 2) + int(pcm_counts[0] // 2) + 8]
    else:
        slab = gathered = None
        my_pcm = torch.zeros(int(pcm_counts[rank] // 2) + 8, dtype=torch.int16, device=device)
    my_out = torch.zeros(my_out_bytes + 256, dtype=torch.uint8, device=device)
    gc_frames = int(((gc_lens.astype(np.int64) + 13) // 14).sum())
    gws_bytes = int(vg.lib.vgb_gcadpcm_workspace_bytes(gc_frames, max(n_gc, 1)))
    gws = torch.empty(gws_bytes, dtype=torch.uint8, device=device)
    aws_bytes = int(vg.lib.vgb_adx_workspace_bytes(int(adx_lens.astype(np.int64).sum()), max(n_adx, 1)))
    aws = torch.empty(aws_bytes, dtype=torch.uint8, device=device)
    adx_params = (N.VgbAdxParams * max(n_adx, 1))()
    for i in range(n_adx):
        adx_params[i] = N.VgbAdxParams(SAMPLE_RATE, 500, 18, 4, 0, 0, 3, 0)

    # ---- communicator inside the library (the id travels over torch.distributed)
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_uint8 * 128)()
            N.check(vg.lib.vgb_nccl_unique_id(raw))
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.to(device)
        dist.broadcast(idbuf, 0)
        raw = (C.c_uint8 * 128)(*idbuf.cpu().tolist())
        N.check(vg.lib.vgb_nccl_init(raw, world, rank))

    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(record=False):
        if record:
            evs[0].record(stream)
        if world > 1:
            N.check(vg.lib.vgb_scatterv_dev(slab.data_ptr() if rank == 0 else None, pcm_offsets.ctypes.data, pcm_counts.ctypes.data,
                                            my_pcm.data_ptr(), 0, stream.cuda_stream))
        if record:
            evs[1].record(stream)
        if n_gc:
            N.check(vg.lib.vgb_gcadpcm_encode_dev(my_pcm.data_ptr(), gc_off.ctypes.data, gc_lens.ctypes.data, None, n_gc, None,
                                                  my_out.data_ptr() + coef_at, my_out.data_ptr(), gc_out_off.ctypes.data, gws.data_ptr(),
                                                  gws_bytes, stream.cuda_stream))
        if n_adx:
            N.check(vg.lib.vgb_adx_encode_dev(my_pcm.data_ptr(), adx_off.ctypes.data, adx_lens.ctypes.data, adx_params, n_adx, None,
                                              my_out.data_ptr(), adx_out_off.ctypes.data, aws.data_ptr(), aws_bytes, stream.cuda_stream))
        if record:
            evs[2].record(stream)
        if world > 1:
            N.check(vg.lib.vgb_gatherv_dev(my_out.data_ptr(), gathered.data_ptr() if rank == 0 else None, out_offsets.ctypes.data,
                                           out_counts.ctypes.data, 0, stream.cuda_stream))
        if record:
            evs[3].record(stream)

    N.check(vg.lib.vgb_set_kernel_timing(1))
    for _ in range(args.warmup):
        step()
    _barrier(torch, dist, world)
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    launches0 = vg.lib.vgb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = np.zeros(3)
    ev0.record(stream)
    for _ in range(args.steps):
        step(record=True)
        torch.cuda.synchronize()
        phase += np.array([evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])])
    ev1.record(stream)
    torch.cuda.synchronize()
    my_ms = ev0.elapsed_time(ev1) / args.steps
    ms = _max_over_ranks(torch, dist, world, device, my_ms)
    launches = vg.lib.vgb_kernel_launch_count() - launches0
    clocks = sampler.stop()
    phase /= args.steps
    enc_all = np.zeros(world)
    enc_all[rank] = phase[1]
    sc_ms = _max_over_ranks(torch, dist, world, device, phase[0])
    ga_ms = _max_over_ranks(torch, dist, world, device, phase[2])
    if world > 1:
        t = torch.as_tensor(enc_all, device=device)
        dist.all_reduce(t)
        enc_all = t.cpu().numpy()
    total = int(lens.sum())
    value = total / (ms / 1e3) / 1e6

    # ---- e2e: every rank's files from ITS OWN pinned host memory through the host API (one H2D link per GPU)
    e2e = None
    if not args.no_e2e:
        torch.cuda.synchronize()
        h_pcm = torch.empty(my_pcm.numel(), dtype=torch.int16, pin_memory=True)
        h_pcm.copy_(my_pcm)
        h_out = torch.empty(my_out_bytes + 256, dtype=torch.uint8, pin_memory=True)
        gc_in = (C.c_void_p * max(n_gc, 1))(*[h_pcm.data_ptr() + 2 * int(o) for o in gc_off])
        gc_tab = (C.c_void_p * max(n_gc, 1))(*[h_out.data_ptr() + int(o) for o in gc_out_off])
        ad_in = (C.c_void_p * max(n_adx, 1))(*[h_pcm.data_ptr() + 2 * int(o) for o in adx_off])
        ad_o = (C.c_void_p * max(n_adx, 1))(*[h_out.data_ptr() + int(o) for o in adx_out_off])
        h_coefs = np.zeros((max(n_gc, 1), 16), dtype=np.int16)

        def step_e2e():
            if n_gc:
                N.check(vg.lib.vgb_gcadpcm_encode_batch(gc_in, gc_lens.ctypes.data, None, None, n_gc, h_coefs.ctypes.data, gc_tab, None, None))
            if n_adx:
                N.check(vg.lib.vgb_adx_encode_batch(ad_in, adx_lens.ctypes.data, adx_params, n_adx, None, ad_o, None, None))

        step_e2e()
        _barrier(torch, dist, world)
        t0 = time.perf_counter()
        step_e2e()
        e_ms = _max_over_ranks(torch, dist, world, device, (time.perf_counter() - t0) * 1e3)
        same = bool((h_out[:gc_out_total].to(device) == my_out[:gc_out_total]).all().item()) if n_gc else True
        if n_adx:
            same = same and bool((h_out[adx_at:my_out_bytes].to(device) == my_out[adx_at:my_out_bytes]).all().item())
        e2e = {"value": round(total / (e_ms / 1e3) / 1e6, 3), "unit": "Msamples/s", "ms_per_step": round(e_ms, 3),
               "h2d_bytes_per_step": int(total * 2), "d2h_bytes_per_step": int(out_counts.sum()),
               "api": "vgb_gcadpcm_encode_batch + vgb_adx_encode_batch per rank on its own files from pinned host memory (one PCIe link per GPU)",
               "matches_device_resident": same}

    # ---- parity on the root: 128 GC + 128 ADX files drawn across the gathered buffer, against the oracle
    parity = cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import pyoracle

        src_buf = gathered if world > 1 else my_out
        base0 = 0 if world == 1 else None
        rng = np.random.default_rng(7)
        checked_gc = checked_adx = 0
        ok = True
        cpu_samples, cpu_t = 0, 0.0
        for r in range(world):
            rlo, rhi = int(rank_lo[r]), int(rank_hi[r])
            r_lens, r_gc = lens_o[rlo:rhi], gc_o[rlo:rhi]
            r_ngc = int(r_gc.sum())
            r_gcb = np.array([vg.gcadpcm.sample_count_to_byte_count(int(v)) for v in r_lens[:r_ngc]], dtype=np.int64)
            r_adb = np.array([vg.lib.vgb_adx_encoded_byte_count(int(v), 0, 18) for v in r_lens[r_ngc:]], dtype=np.int64)
            r_gc_off = np.concatenate(([0], np.cumsum((r_gcb + 15) // 16 * 16)[:-1])).astype(np.int64) if r_ngc else np.zeros(0, np.int64)
            r_coef_at = int(((r_gcb + 15) // 16 * 16).sum())
            r_adx_at = (r_coef_at + r_ngc * 32 + 15) // 16 * 16
            r_adx_off = (r_adx_at + np.concatenate(([0], np.cumsum((r_adb + 15) // 16 * 16)[:-1]))).astype(np.int64) if len(r_adb) else np.zeros(0, np.int64)
            base = int(out_offsets[r]) if world > 1 else 0
            want_each = max(1, 128 // world)
            for i in rng.choice(r_ngc, min(want_each, r_ngc), replace=False) if r_ngc else []:
                L = int(r_lens[i])
                o = int(offs_o[rlo + i])
                x = slab[o:o + L].cpu().numpy()
                t0 = time.perf_counter()
                co = pyoracle.calculate_coefficients(x)
                want = pyoracle.encode(x, co)
                cpu_t += time.perf_counter() - t0
                cpu_samples += L
                got = src_buf[base + int(r_gc_off[i]):base + int(r_gc_off[i]) + int(r_gcb[i])].cpu().numpy()
                gco = src_buf[base + r_coef_at + 32 * int(i):base + r_coef_at + 32 * int(i) + 32].cpu().numpy().view(np.int16)
                ok = ok and np.array_equal(got, want) and np.array_equal(gco, co)
                checked_gc += 1
            n_ad = len(r_adb)
            for i in rng.choice(n_ad, min(want_each, n_ad), replace=False) if n_ad else []:
                L = int(r_lens[r_ngc + i])
                o = int(offs_o[rlo + r_ngc + i])
                x = slab[o:o + L].cpu().numpy()
                want, _ = pyoracle.adx_encode(x, SAMPLE_RATE, 18, 4, 0, 3, 0)
                got = src_buf[base + int(r_adx_off[i]):base + int(r_adx_off[i]) + int(r_adb[i])].cpu().numpy()
                ok = ok and np.array_equal(got, want)
                checked_adx += 1
        parity = {"gc_files_checked": checked_gc, "adx_files_checked": checked_adx, "bytes_equal_oracle": bool(ok)}
        cores = os.cpu_count() or 1
        if cpu_t > 0:
            cpu = {"value": round(cpu_samples / cpu_t / 1e6 * cores, 3), "unit": "Msamples/s", "cores": cores, "kind": "port",
                   "sample": f"{checked_gc} GC-ADPCM files ({cpu_samples} samples) timed on ONE core ({cpu_samples / cpu_t / 1e6:.2f} Msamples/s) "
                             f"x {cores} cores: the reference's Parallel.ForEach over files is embarrassingly parallel"}

    if world > 1:
        N.check(vg.lib.vgb_nccl_shutdown())
    if rank != 0:
        return None
    mean_enc = float(enc_all.mean()) if world > 0 else 0.0
    line = _base_line("mixed GC-ADPCM + ADX batch encode Msamples/sec", value, world, args, ms, "strong", "int32",
                      {"workload": f"{n_files} mono files, 1-10 s x 48 kHz, even -> GC-ADPCM (coefs + encode), odd -> CRI ADX (Linear, v4, 18 B frames); whole job",
                       "total_samples": total, "files_per_rank": [int(rank_hi[r] - rank_lo[r]) for r in range(world)],
                       "l2": "inputs (34.6 GB) larger than L2, no flush needed", "parallelism": f"{world} ranks, files partitioned longest-first, one NCCL scatterv + gatherv per step"})
    line.update({"e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                 "collective": {"scatter_ms": round(sc_ms, 3), "gather_ms": round(ga_ms, 3), "scatter_bytes": int(pcm_counts.sum() - pcm_counts[0]),
                                "gather_bytes": int(out_counts.sum() - out_counts[0]), "nccl_version": int(vg.lib.vgb_nccl_version()),
                                "encode_ms_per_rank": [round(float(v), 3) for v in enc_all],
                                "imbalance": round(float(enc_all.max() / mean_enc), 4) if mean_enc > 0 else None},
                 "roofline": {"bound": "hbm", "kernel": "gc_encode_kernel + adx_encode_kernel", "achieved": round(total * 2.567 / (float(enc_all.max()) / 1e3) / 1e9 / world, 2) if enc_all.max() > 0 else None,
                              "peak": _peak()[0], "unit": "GB/s", "frac": round(total * 2.567 / (float(enc_all.max()) / 1e3) / 1e9 / world / _peak()[0], 5) if enc_all.max() > 0 else None,
                              "traffic": None, "peak_source": _peak()[1], "note": "per-GPU algorithmic bytes (2 B in + ~0.57 B out per sample) over the slowest rank's encode time"},
                 "cpu_baseline": cpu, "parity": parity})
    return line
