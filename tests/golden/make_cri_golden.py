"""Regenerates tests/golden/cri_golden.json: frozen ORACLE outputs of the CRI ADX and CRI HCA codecs.

Like gcadpcm_golden.json these are not reference outputs (the C# reference cannot run in the build container and holds
no golden bitstreams for these codecs): they freeze the oracle + synthetic generator so that an accidental change to
either shows up as a diff, and they travel to the GPU box where the CUDA path is compared against them without running
the oracle.  Run from the repo root:  python tests/golden/make_cri_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from vgaudio_b200 import synth  # noqa: E402

ADX_CASES = [  # (name, synth index, samples, type, version, frame_size)
    ("linear_v4", 4, 32 * 300, 3, 4, 18), ("linear_v3", 5, 32 * 300 + 7, 3, 3, 18), ("fixed_v4", 6, 5000, 2, 4, 18),
    ("exponential_v4", 7, 9999, 4, 4, 18), ("linear_v4_fs34", 40, 64 * 100 + 3, 3, 4, 34),
]
HCA_CASES = [  # (name, first synth index, channels, samples, quality, loop)
    ("mono_high", 100, 1, 20000, 2, None), ("stereo_high", 102, 2, 20000, 2, None), ("stereo_lowest", 104, 2, 12345, 5, None),
    ("mono_highest", 106, 1, 5000, 1, None), ("six_ch_high", 108, 6, 8000, 2, None),
    ("mono_high_loop", 114, 1, 30000, 2, (5000, 25000)), ("stereo_low_loop", 116, 2, 30000, 4, (1024, 20000)),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    pyoracle.build()
    out = {"generator": "tests/golden/make_cri_golden.py", "source": "oracle/criadx.c, oracle/crihca.c (CPU restatements)",
           "adx": [], "hca": []}
    for name, index, n, typ, version, frame_size in ADX_CASES:
        pcm = synth.channel(index, n, degenerate=False)
        adpcm, hist = pyoracle.adx_encode(pcm, 48000, frame_size, version, 0, typ, 2)
        dec = pyoracle.adx_decode(adpcm, n, 48000, 500, frame_size, version, hist, 0, typ)
        out["adx"].append({"name": name, "index": index, "n": n, "type": typ, "version": version, "frame_size": frame_size,
                           "pcm_sha256": sha(pcm), "history": int(hist), "adpcm_sha256": sha(adpcm),
                           "adpcm_head_hex": np.asarray(adpcm)[:36].tobytes().hex(), "decoded_sha256": sha(dec)})
    for name, first, nch, n, quality, loop in HCA_CASES:
        chans = [synth.channel(first + c, n, degenerate=False) for c in range(nch)]
        info, frames = pyoracle.hca_encode(chans, 48000, quality, loop=loop)
        dec = pyoracle.hca_decode(info, frames)
        out["hca"].append({"name": name, "first_index": first, "channels": nch, "n": n, "quality": quality, "loop": loop,
                           "pcm_sha256": sha(np.stack(chans)), "info": info.as_dict(), "frames_sha256": sha(frames),
                           "frame0_hex": frames[0].tobytes().hex()[:96], "decoded_sha256": sha(dec)})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cri_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
