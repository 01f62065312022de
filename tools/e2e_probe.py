#!/usr/bin/env python
"""Timeline of the host GC-ADPCM encode call (vgb_gcadpcm_encode_batch, pinned buffers) on the C2 batch for several
pipeline settings (VGB_ENCODE_GROUPS x VGB_GC_SEGMENTS).  Tuning tool, not a bench value.

  python tools/e2e_probe.py [channels] [seconds] [groups,groups,...] [segments,segments,...]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vgaudio_b200 as vg  # noqa: E402
from vgaudio_b200 import _native as N  # noqa: E402


def main():
    n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
    groups = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 4, 8, 16]
    segs = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 3, 8]
    n = int(seconds * 48000)
    dev = torch.device("cuda", 0)
    N.check(vg.lib.vgb_init(0, 0))
    pcm = bench.make_batch_gpu(torch, n_ch, n, 0, dev)
    n_bytes = vg.gcadpcm.sample_count_to_byte_count(n)
    pcm_host_t = torch.empty((n_ch, n), dtype=torch.int16, pin_memory=True)
    pcm_host_t.copy_(pcm)
    del pcm
    adpcm_host_t = torch.empty((n_ch, n_bytes), dtype=torch.uint8, pin_memory=True)
    coefs_host = np.zeros((n_ch, 16), dtype=np.int16)
    lens = np.full(n_ch, n, dtype=np.int32)
    in_tab = (C.c_void_p * n_ch)(*[pcm_host_t.data_ptr() + 2 * n * c for c in range(n_ch)])
    out_tab = (C.c_void_p * n_ch)(*[adpcm_host_t.data_ptr() + n_bytes * c for c in range(n_ch)])

    def call():
        N.check(vg.lib.vgb_gcadpcm_encode_batch(in_tab, lens.ctypes.data, None, None, n_ch, coefs_host.ctypes.data,
                                                out_tab, None, None))

    ref = None
    for g in groups:
        for s in segs:
            for k, v in (("VGB_ENCODE_GROUPS", g), ("VGB_GC_SEGMENTS", s)):
                if v > 0:
                    os.environ[k] = str(v)
                else:
                    os.environ.pop(k, None)
            call()
            torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                t0 = time.perf_counter()
                call()
                ms.append((time.perf_counter() - t0) * 1e3)
            tl = (C.c_float * 48)()
            N.check(vg.lib.vgb_debug_last_timeline(tl, 48))
            tc = (C.c_float * 16)()
            N.check(vg.lib.vgb_debug_last_coefs_done(tc, 16))
            st = (C.c_uint64 * 19)()
            N.check(vg.lib.vgb_gcadpcm_debug_splice_stats(st, 19))
            if ref is None:
                ref = adpcm_host_t.clone()
            row = {"groups": g, "segments": s, "ms": [round(m, 1) for m in ms],
                   "h2d": [round(tl[3 * i], 1) for i in range(16) if tl[3 * i] >= 0],
                   "coefs": [round(tc[i], 1) for i in range(16) if tl[3 * i] >= 0],
                   "kern": [round(tl[3 * i + 1], 1) for i in range(16) if tl[3 * i] >= 0],
                   "d2h": [round(tl[3 * i + 2], 1) for i in range(16) if tl[3 * i] >= 0],
                   "last_group_segments": int(st[0]), "last_group_runon": int(st[1]), "last_group_cascade": int(st[2]),
                   "longest_runon": int(st[4]), "same": bool((adpcm_host_t == ref).all().item())}
            print(json.dumps(row), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
