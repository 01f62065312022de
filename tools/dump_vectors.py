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
  wave_to_dsp / wave_to_adx / wave_to_hca   input = a WAVE file, output = the finished container file (the batch converter's
                  output; the managed side runs WaveReader + DspWriter / AdxWriter / HcaWriter with the params' key options)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vgaudio_b200 import synth  # noqa: E402  (data generation only)


def _wave16(chans, rate, loop):
    """A 16-bit PCM WAVE image (RIFF / fmt [extensible above two channels] / smpl when looping / data) - input data only."""
    ch, n = len(chans), len(chans[0])
    data = np.stack(chans, axis=1).astype("<i2").tobytes()
    u16, u32, i32 = (lambda *v: np.array(v, dtype="<u2").tobytes()), (lambda *v: np.array(v, dtype="<u4").tobytes()), (lambda *v: np.array(v, dtype="<i4").tobytes())
    if ch > 2:
        mask = {4: 0x33, 5: 0x133, 6: 0x633, 7: 0x1f3, 8: 0x6f3}.get(ch, (1 << ch) - 1)
        guid = bytes([0x01, 0, 0, 0, 0, 0, 0x10, 0, 0x80, 0, 0, 0xAA, 0, 0x38, 0x9B, 0x71])
        fmt = u16(0xFFFE, ch) + u32(rate, rate * 2 * ch) + u16(2 * ch, 16) + u16(22, 16) + u32(mask) + guid
    else:
        fmt = u16(1, ch) + u32(rate, rate * 2 * ch) + u16(2 * ch, 16)
    smpl = b""
    if loop:
        body = [0] * 15
        body[7] = 1
        body[11], body[12] = loop
        smpl = b"smpl" + u32(0x3c) + i32(*body)
    body = b"WAVE" + b"fmt " + u32(len(fmt)) + fmt + smpl + b"data" + u32(len(data)) + data
    return np.frombuffer(b"RIFF" + u32(len(body)) + body, dtype=np.uint8)


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

    # ---- containers: WAVE file in -> finished .dsp / .adx / .hca file out, and .dsp -> .wav.  The managed
    # side runs WaveReader -> DspWriter / AdxWriter / HcaWriter (GetFile) with the options in the params column.
    if use_oracle:
        ow = o
    else:
        from vgaudio_b200 import containers as ct
    k = 0
    for nch, n, loop in ((1, 30000, None), (2, 20011, (1001, 19000)), (6, 9000, None), (1, 50000, (0, 50000))):
        chans = [synth.channel(150 + k * 8 + c, n, degenerate=False) for c in range(nch)]
        wav = _wave16(chans, 32000, loop)
        fin = put(f"ctn_{k}.wav", wav)
        for kind, ext, params in (("dsp", "dsp", ""), ("adx", "adx", "keystring=karaage"), ("adx_plain", "adx", ""), ("hca", "hca", "quality=2,keycode=12345")):
            if use_oracle:
                st, info = ow.wave_parse(wav)
                rows = ow.wave_read(wav, info)
                if kind == "dsp":
                    co = np.stack([ow.calculate_coefficients(r) for r in rows])
                    ad = [ow.encode(r, c) for r, c in zip(rows, co)]
                    ctxs = np.stack([np.array(ow.gc_loop_context(a, ow.decode(a, c, n), loop[0]), dtype=np.int16) for a, c in zip(ad, co)]) if loop else None
                    blob = ow.dsp_write(ad, co, 32000, n, loop, ctxs)
                elif kind.startswith("adx"):
                    align = (-loop[0]) % (64 if nch == 1 else 32) if loop else 0
                    enc = [ow.adx_encode(r, 32000, 18, 4, align, 3, 0) for r in rows]
                    key = ow.adx_key(key_string="karaage") if kind == "adx" else None
                    blob = ow.adx_write([e[0] for e in enc], [e[1] for e in enc], 32000, n, loop, align, 18, 4, 3, 500, 8 if key else 0, key)
                else:
                    hi, fr = ow.hca_encode(rows, 32000, quality=2, loop=loop)
                    blob = ow.hca_write(hi, fr, ow.hca_key_tables(56, 12345)[1], 56)
            else:
                if kind == "dsp":
                    opt = ct.convert_options(ct.CONTAINER_DSP)
                elif kind == "adx":
                    kk = ct.adx_key(key_string="karaage")
                    opt = ct.convert_options(ct.CONTAINER_ADX, adx_has_key=1, adx_key_seed=kk.seed, adx_key_mult=kk.mult, adx_key_inc=kk.inc, adx_encryption_type=8)
                elif kind == "adx_plain":
                    opt = ct.convert_options(ct.CONTAINER_ADX)
                else:
                    opt = ct.convert_options(ct.CONTAINER_HCA, hca_quality=2, hca_key_type=56, hca_key_code=12345)
                blob = ct.convert_wave_batch([wav], opt)[0][0]
            lines.append(f"wave_to_{kind.split('_')[0]}\tctn{k}_{kind}\t{params}\t{fin}\t{put(f'ctn_{k}_{kind}.{ext}', blob)}")
        k += 1

    with open(os.path.join(out, "manifest.tsv"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(f"{len(lines) - 1} cases -> {out} ({'oracle' if use_oracle else 'CUDA path'})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
