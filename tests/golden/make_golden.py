"""Regenerates tests/golden/gcadpcm_golden.json from the CPU oracle.

The reference (C#) cannot be executed in the build container (no .NET), so these vectors are ORACLE outputs, not
reference outputs: they freeze the oracle (and the synthetic generator) so that an accidental change to either shows
up as a diff, and they travel to the GPU box where the CUDA path is compared against them.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from vgaudio_b200 import synth  # noqa: E402

CASES = [
    ("zero", 0, 4200, True), ("square", 1, 4200, True), ("ref_sine", 2, 48000, True), ("ramp", 3, 1000, True),
    ("mix4", 4, 48000, True), ("mix5", 5, 48001, True), ("mix6", 6, 14 * 1000 + 5, True), ("mix7", 7, 9999, True),
    ("mix12_short", 12, 13, True), ("mix13_1", 13, 1, True), ("mix40", 40, 96000, True),
]


def main():
    pyoracle.build()
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/gcadpcm.c (CPU restatement)", "cases": []}
    for name, index, n, degenerate in CASES:
        pcm = synth.channel(index, n, degenerate=degenerate)
        coefs = pyoracle.calculate_coefficients(pcm)
        adpcm = pyoracle.encode(pcm, coefs)
        out["cases"].append({
            "name": name, "index": index, "n": n, "degenerate": degenerate,
            "pcm_sha256": hashlib.sha256(pcm.tobytes()).hexdigest(),
            "coefs": coefs.tolist(),
            "adpcm_sha256": hashlib.sha256(adpcm.tobytes()).hexdigest(),
            "adpcm_head_hex": adpcm[:32].tobytes().hex(),
        })
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gcadpcm_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
