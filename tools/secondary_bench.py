#!/usr/bin/env python
"""tools/secondary_bench.py — wall-clock throughput of the non-headline paths through the host C ABI (pageable numpy
buffers, so H2D/D2H are part of every number): GC-ADPCM decode (config C3 shape, scaled), CRI ADX encode/decode and
CRI HCA encode (config C4 shape, scaled).  Usage: python tools/secondary_bench.py [--scale 0.25]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vgaudio_b200 as vg  # noqa: E402
from vgaudio_b200 import synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    a = ap.parse_args()
    out = {}
    rng = np.random.default_rng(0)
    base = synth.batch(16, 48000 * 30, degenerate=False, first_index=10)

    # GC-ADPCM decode, C3 shape: 8192 channels x 30 s (scaled)
    n_ch = max(64, int(8192 * a.scale))
    pcm = base[np.arange(n_ch) % 16]
    coefs, adpcm = vg.gcadpcm.encode_batch(pcm)
    ad = np.stack(adpcm)
    cfgs = [vg.gcadpcm.GcAdpcmParameters(pcm.shape[1])] * n_ch
    dt = timed(lambda: vg.gcadpcm.decode_batch(ad, coefs, cfgs))
    out["gcadpcm_decode"] = {"channels": n_ch, "Msamples_per_s": round(pcm.size / dt / 1e6, 1), "ms": round(dt * 1e3, 1)}

    # CRI ADX encode / decode
    n_ch = max(64, int(4096 * a.scale))
    pcm = base[np.arange(n_ch) % 16]
    cfg = vg.criadx.CriAdxParameters()
    dt = timed(lambda: vg.criadx.encode_batch(pcm, cfg))
    out["adx_encode"] = {"channels": n_ch, "Msamples_per_s": round(pcm.size / dt / 1e6, 1), "ms": round(dt * 1e3, 1)}
    adx, hist = vg.criadx.encode_batch(pcm, cfg)
    dcfg = [vg.criadx.CriAdxParameters(history=int(h)) for h in hist]
    dt = timed(lambda: vg.criadx.decode_batch(np.stack(adx), pcm.shape[1], dcfg))
    out["adx_decode"] = {"channels": n_ch, "Msamples_per_s": round(pcm.size / dt / 1e6, 1), "ms": round(dt * 1e3, 1)}

    # CRI HCA encode, C4 shape: 512 mono streams x 30 s, Quality=High (scaled)
    n_st = max(16, int(512 * a.scale))
    streams = [[base[s % 16]] for s in range(n_st)]
    dt = timed(lambda: vg.crihca.encode_batch(streams, 48000))
    out["hca_encode_mono_high"] = {"streams": n_st, "Msamples_per_s": round(n_st * base.shape[1] / dt / 1e6, 1), "ms": round(dt * 1e3, 1)}
    streams2 = [[base[s % 16], base[(s + 1) % 16]] for s in range(n_st // 2)]
    dt = timed(lambda: vg.crihca.encode_batch(streams2, 48000))
    out["hca_encode_stereo_high"] = {"streams": n_st // 2, "Msample_channels_per_s": round(n_st // 2 * 2 * base.shape[1] / dt / 1e6, 1), "ms": round(dt * 1e3, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
