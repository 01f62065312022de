#!/usr/bin/env python
"""tools/secondary_bench.py — the non-headline paths of SURVEY.md §8 measured on one GPU.

For each of GC-ADPCM decode (BASELINE config C3 shape), CRI ADX encode/decode and CRI HCA encode/decode (C4 shape) it
reports, through the host C ABI (pageable numpy buffers):
  wall_ms / Msamples_per_s   the synchronous call, H2D and D2H included
  kernel_ms                  CUDA events around the kernel(s) on the library's stream (vgb_set_kernel_timing)
  roofline                   algorithmic bytes of the kernel / kernel_ms against MEASURED_PEAKS.json's HBM copy peak
Shapes are scaled by --scale (default 0.25 of C3/C4) to keep pageable host buffers and run time moderate.
Usage: python tools/secondary_bench.py [--scale 0.25] [--only gcdec,adx,hca]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vgaudio_b200 as vg  # noqa: E402
from vgaudio_b200 import _native as N  # noqa: E402
from vgaudio_b200 import synth  # noqa: E402

SLOT = {"gc_decode": 3, "adx_encode": 4, "adx_decode": 5, "hca_encode": 6, "hca_decode": 7, "interleave": 8, "deinterleave": 9}


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 7700.0, "fallback (B200_PROFILING.md nominal)"


def timed(fn, slot, reps=3):
    fn()
    best, kms = None, None
    buf = (C.c_float * 10)()
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        N.check(vg.lib.vgb_last_kernel_ms(buf, 10))
        if best is None or dt < best:
            best, kms = dt, float(buf[slot])
    return best, kms


def entry(samples, dt, kms, alg_bytes, extra):
    peak, src = peak_gbs()
    out = dict(extra)
    out.update({"Msamples_per_s": round(samples / dt / 1e6, 1), "wall_ms": round(dt * 1e3, 1), "kernel_ms": round(kms, 3),
                "kernel_Msamples_per_s": round(samples / kms / 1e3, 1) if kms > 0 else None,
                "roofline": {"bound": "hbm", "achieved": round(alg_bytes / kms / 1e6, 1) if kms > 0 else None, "peak": peak,
                             "unit": "GB/s", "frac": round(alg_bytes / kms / 1e6 / peak, 4) if kms > 0 else None,
                             "algorithmic_bytes": int(alg_bytes), "peak_source": src}})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--only", default="gcdec,adx,hca,interleave")
    a = ap.parse_args()
    only = set(a.only.split(","))
    N.check(vg.lib.vgb_set_kernel_timing(1))
    out = {}
    n = int(48000 * a.seconds)
    base = synth.batch(16, n, degenerate=False, first_index=10)

    if "gcdec" in only:  # GC-ADPCM decode, C3 shape: 8192 channels x 30 s (scaled)
        n_ch = max(64, int(8192 * a.scale))
        coefs16, adpcm16 = vg.gcadpcm.encode_batch(base)
        idx = np.arange(n_ch) % 16
        ad = np.stack(adpcm16)[idx]
        coefs = np.asarray(coefs16)[idx]
        cfgs = [vg.gcadpcm.GcAdpcmParameters(n)] * n_ch
        dt, kms = timed(lambda: vg.gcadpcm.decode_batch(ad, coefs, cfgs), SLOT["gc_decode"])
        out["gcadpcm_decode"] = entry(n_ch * n, dt, kms, n_ch * (n * 2 + ad.shape[1]), {"channels": n_ch})

    if "adx" in only:  # CRI ADX encode / decode
        n_ch = max(64, int(4096 * a.scale))
        pcm = base[np.arange(n_ch) % 16]
        cfg = vg.criadx.CriAdxParameters()
        dt, kms = timed(lambda: vg.criadx.encode_batch(pcm, cfg), SLOT["adx_encode"])
        adx, hist = vg.criadx.encode_batch(pcm, cfg)
        adx = np.stack(adx)
        out["adx_encode"] = entry(pcm.size, dt, kms, pcm.size * 2 + adx.size, {"channels": n_ch})
        dcfg = [vg.criadx.CriAdxParameters(history=int(h)) for h in hist]
        dt, kms = timed(lambda: vg.criadx.decode_batch(adx, n, dcfg), SLOT["adx_decode"])
        out["adx_decode"] = entry(pcm.size, dt, kms, pcm.size * 2 + adx.size, {"channels": n_ch})

    if "hca" in only:  # CRI HCA, C4 shape: 512 mono streams x 30 s, Quality=High (scaled)
        n_st = max(16, int(512 * a.scale))
        streams = [[base[s % 16]] for s in range(n_st)]
        dt, kms = timed(lambda: vg.crihca.encode_batch(streams, 48000), SLOT["hca_encode"])
        infos, frames = vg.crihca.encode_batch(streams, 48000)
        fbytes = sum(f.size for f in frames)
        out["hca_encode_mono_high"] = entry(n_st * n, dt, kms, n_st * n * 2 + fbytes, {"streams": n_st})
        dt, kms = timed(lambda: vg.crihca.decode_batch(infos, frames), SLOT["hca_decode"])
        out["hca_decode_mono_high"] = entry(n_st * n, dt, kms, n_st * n * 2 + fbytes, {"streams": n_st})
        streams2 = [[base[s % 16], base[(s + 1) % 16]] for s in range(n_st // 2)]
        dt, kms = timed(lambda: vg.crihca.encode_batch(streams2, 48000), SLOT["hca_encode"])
        infos2, frames2 = vg.crihca.encode_batch(streams2, 48000)
        fbytes2 = sum(f.size for f in frames2)
        out["hca_encode_stereo_high"] = entry(n_st // 2 * 2 * n, dt, kms, n_st // 2 * 2 * n * 2 + fbytes2, {"streams": n_st // 2})
    if "interleave" in only:  # device-resident block interleave of .dsp-like payloads (0x2000-byte blocks, 2 channels)
        import torch
        items, count, interleave = max(64, int(8192 * a.scale)), 2, 0x2000
        in_size = gc_bytes = ((n + 13) // 14) * 8
        out_size = -(-in_size // interleave) * interleave  # the writers pad the last block
        src = torch.randint(0, 256, (items, count, in_size), dtype=torch.uint8, device="cuda")
        dst = torch.zeros((items, count * out_size), dtype=torch.uint8, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream

        def run_i():
            N.check(vg.lib.vgb_interleave_dev(src.data_ptr(), in_size, count * in_size, dst.data_ptr(), count * out_size, items, count,
                                              in_size, interleave, out_size, stream))
            torch.cuda.synchronize()
        dt, kms = timed(run_i, SLOT["interleave"])
        e = entry(items * count * in_size, dt, kms, items * count * (in_size + out_size), {"items": items, "channels": count, "interleave": interleave})
        e["unit_note"] = "Msamples_per_s fields count BYTES here"
        out["interleave_dev"] = e
        back = torch.zeros((items, count, in_size), dtype=torch.uint8, device="cuda")

        def run_d():
            N.check(vg.lib.vgb_deinterleave_dev(dst.data_ptr(), count * out_size, back.data_ptr(), in_size, count * in_size, items, count,
                                                out_size, interleave, in_size, stream))
            torch.cuda.synchronize()
        dt, kms = timed(run_d, SLOT["deinterleave"])
        e = entry(items * count * in_size, dt, kms, items * count * (in_size + out_size), {"items": items, "channels": count, "interleave": interleave})
        e["unit_note"] = "Msamples_per_s fields count BYTES here"
        e["round_trip_equal"] = bool(torch.equal(back, src))
        out["deinterleave_dev"] = e
        del gc_bytes
    print(json.dumps(out))


if __name__ == "__main__":
    main()
