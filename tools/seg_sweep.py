#!/usr/bin/env python
"""Sweep of the time-parallel encoder's segment count on the C2 batch (device resident): per-kernel ms for each
VGB_GC_SEGMENTS value, bytes compared with the serial loop (segments = 1).  Tuning tool, not a bench value."""
import ctypes as C
import json
import os
import sys

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
    segs = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3, 4, 6, 7, 10, 13, 17, 24, 31, 48, 64, 0]
    n = int(seconds * 48000)
    dev = torch.device("cuda", 0)
    N.check(vg.lib.vgb_init(0, 0))
    pcm = bench.make_batch_gpu(torch, n_ch, n, 0, dev)
    stride = (n + 7) // 8 * 8
    pcm_dev = torch.zeros((n_ch, stride), dtype=torch.int16, device=dev)
    pcm_dev[:, :n] = pcm
    n_bytes = vg.gcadpcm.sample_count_to_byte_count(n)
    a_stride = (n_bytes + 15) // 16 * 16
    adpcm = torch.zeros((n_ch, a_stride), dtype=torch.uint8, device=dev)
    coefs = torch.zeros((n_ch, 16), dtype=torch.int16, device=dev)
    frames = (n + 13) // 14
    ws_bytes = int(vg.lib.vgb_gcadpcm_workspace_bytes(frames * n_ch, n_ch))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    pcm_off = np.arange(n_ch, dtype=np.int64) * stride
    ad_off = np.arange(n_ch, dtype=np.int64) * a_stride
    lens = np.full(n_ch, n, dtype=np.int32)
    stream = torch.cuda.current_stream()
    N.check(vg.lib.vgb_set_kernel_timing(1))
    ref = None
    rows = []
    for s in segs:
        if s > 0:
            os.environ["VGB_GC_SEGMENTS"] = str(s)
        else:
            os.environ.pop("VGB_GC_SEGMENTS", None)
        ms = []
        for it in range(3):
            adpcm.zero_()
            N.check(vg.lib.vgb_gcadpcm_encode_dev(pcm_dev.data_ptr(), pcm_off.ctypes.data, lens.ctypes.data, None, n_ch, None,
                                                  coefs.data_ptr(), adpcm.data_ptr(), ad_off.ctypes.data, ws.data_ptr(), ws_bytes,
                                                  stream.cuda_stream))
            buf = (C.c_float * 4)()
            N.check(vg.lib.vgb_last_kernel_ms(buf, 4))
            ms.append(list(buf))
        torch.cuda.synchronize()
        st = (C.c_uint64 * 19)()
        N.check(vg.lib.vgb_gcadpcm_debug_splice_stats(st, 19))
        if ref is None:
            ref = adpcm.clone()
        same = bool((adpcm == ref).all().item())
        row = {"segments_forced": s, "segments": int(st[0]), "gc_encode_ms": round(min(m[2] for m in ms), 3),
               "coef_frames_ms": round(min(m[0] for m in ms), 3), "coef_refine_ms": round(min(m[1] for m in ms), 3),
               "runon_frames": int(st[1]), "cascade_frames": int(st[2]), "cascade_boundaries": int(st[3]),
               "fallback_frac": round((int(st[1]) + int(st[2])) / (n_ch * frames), 6), "longest_runon": int(st[4]),
               "runon_log2_hist": [int(st[5 + b]) for b in range(14)], "same_bytes_as_first": same}
        rows.append(row)
        print(json.dumps(row), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
