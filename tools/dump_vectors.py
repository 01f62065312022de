#!/usr/bin/env python
"""tools/dump_vectors.py — writes the seeded inputs and THIS REPOSITORY'S outputs (CUDA path through the C ABI) for a set of
GC-ADPCM / CRI ADX / CRI HCA cases to a directory, so that a machine with .NET (and no GPU) can close the "parity
unpinned" gap: `ParityHarness vectors <dir>` (bindings/csharp/ParityHarness.cs) re-encodes every input with the unmodified
managed VGAudio and diffs byte for byte.  Needs a CUDA device.  With --oracle the CPU oracle writes the outputs instead
(no GPU needed; only meaningful because the GPU path is tested bit-exact against the oracle).

  python tools/dump_vectors.py out_dir [--oracle]

manifest.tsv:  codec <TAB> name <TAB> key=value,... <TAB> input files (comma separated, raw int16 LE per channel) <TAB> output file
  gcadpcm         output = 16 coefficients (int16 LE) followed by the ADPCM bytes of GcAdpcmEncoder.Encode
  gcadpcm_decode  input = such a blob, output = the decoded PCM16
  criadx          output = CriAdxCodec.Encode bytes
  crihca          output = the frames of CriHcaFormat.EncodeFromPcm16, concatenated
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vgaudio_b200 import synth  # noqa: E402  (data generation only)


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        return 2
    out = sys.argv[1]
    use_oracle = "--oracle" in sys.argv[2:]
    os.makedirs(out, exist_ok=True)
    if use_oracle:
        from oracle import pyoracle as o
    else:
        import vgaudio_b200 as vg

        vg._native.check(vg.lib.vgb_init(0, 0))
    lines = ["# codec\tname\tparams\tinputs\toutput"]

    def put(name, arr):
        np.ascontiguousarray(arr).tofile(os.path.join(out, name))
        return name

    # ---- GC-ADPCM: the synthetic set (degenerate channels included), ragged lengths, the reference's own sines
    gc_cases = [(i, n) for i, n in [(0, 4000), (1, 4000), (2, 48000), (3, 20000), (4, 48000), (5, 14 * 3000 + 5), (11, 30000), (12, 96000), (20, 13), (21, 1)]]
    for idx, n in gc_cases:
        pcm = synth.channel(idx, n)
        fin = put(f"gc_{idx}_{n}.pcm", pcm)
        if use_oracle:
            coefs = o.calculate_coefficients(pcm)
            adpcm = o.encode(pcm, coefs)
            dec = o.decode(adpcm, coefs, n)
        else:
            coefs_b, adpcm_b = vg.gcadpcm.encode_batch([pcm])
            coefs, adpcm = coefs_b[0], adpcm_b[0]
            dec = vg.gcadpcm.decode(adpcm, coefs, vg.gcadpcm.GcAdpcmParameters(n))
        blob = np.concatenate([np.asarray(coefs, dtype="<i2").view(np.uint8), np.asarray(adpcm, dtype=np.uint8)])
        fb = put(f"gc_{idx}_{n}.dsp", blob)
        lines.append(f"gcadpcm\tch{idx}_n{n}\tsample_count={n}\t{fin}\t{fb}")
        lines.append(f"gcadpcm_decode\tch{idx}_n{n}\tsample_count={n}\t{fb}\t{put(f'gc_{idx}_{n}.dec', np.asarray(dec, dtype='<i2'))}")

    # ---- CRI ADX: types x versions x padding x frame sizes
    k = 0
    for typ in (2, 3, 4):
        for version in (3, 4):
            for padding, frame_size in ((0, 18), (45, 18), (0, 34)):
                n = 9000 + 37 * k
                pcm = synth.channel(30 + k, n)
                fin = put(f"adx_{k}.pcm", pcm)
                filt = k % 4
                if use_oracle:
                    enc, _ = o.adx_encode(pcm, 48000, frame_size, version, padding, typ, filt)
                else:
                    cfg = vg.criadx.CriAdxParameters(sample_rate=48000, frame_size=frame_size, version=version, padding=padding, type=typ, filter=filt)
                    enc = vg.criadx.encode_batch([pcm], [cfg])[0][0]
                lines.append(f"criadx\tt{typ}_v{version}_p{padding}_f{frame_size}\tsample_rate=48000,frame_size={frame_size},version={version},"
                             f"padding={padding},type={typ},filter={filt}\t{fin}\t{put(f'adx_{k}.adx', np.asarray(enc, dtype=np.uint8))}")
                k += 1

    # ---- CRI HCA: qualities x channel counts, two looping cases
    k = 0
    for quality in (1, 2, 3, 4, 5):
        for nch in (1, 2, 3, 6, 8):
            n = 20000 + 333 * k
            chans = [synth.channel(60 + k * 8 + c, n, degenerate=False) for c in range(nch)]
            fins = [put(f"hca_{k}_{c}.pcm", ch) for c, ch in enumerate(chans)]
            loop = (4000, 18000) if (nch == 2 and quality in (2, 4)) else None
            if use_oracle:
                _, frames = o.hca_encode(chans, 48000, quality, loop=loop)
            else:
                cfg = vg.crihca.CriHcaParameters(quality=quality, looping=loop is not None, loop_start=loop[0] if loop else 0, loop_end=loop[1] if loop else 0)
                _, frames = vg.crihca.encode(chans, 48000, cfg)
            params = (f"quality={quality},bitrate=0,limit_bitrate=0,sample_rate=48000,looping={1 if loop else 0},"
                      f"loop_start={loop[0] if loop else 0},loop_end={loop[1] if loop else 0}")
            lines.append(f"crihca\tq{quality}_ch{nch}{'_loop' if loop else ''}\t{params}\t{','.join(fins)}\t{put(f'hca_{k}.frames', np.asarray(frames, dtype=np.uint8))}")
            k += 1

    with open(os.path.join(out, "manifest.tsv"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(f"{len(lines) - 1} cases -> {out} ({'oracle' if use_oracle else 'CUDA path'})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
