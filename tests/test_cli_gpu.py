"""vgaudio_batch (vgaudio_b200/cli): a directory of WAVE files in, a directory of encoded files out - the CLI's batch job
(src/VGAudio.Cli/Batch.cs:11-51) - checked file by file against the oracle's reader -> encoder -> writer chain."""
import os
import subprocess

import numpy as np
import pytest

from vgaudio_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "vgaudio_b200", "cli", "vgaudio_batch")


def test_batch_directory_to_dsp_and_hca(tmp_path, oracle):
    src = tmp_path / "in"
    (src / "sub").mkdir(parents=True)
    specs = {"a.wav": (1, 30000, None), "b.WAV": (2, 12345, (100, 12000)), "sub/c.wav": (3, 5000, None)}
    pcm = {}
    for k, (name, (ch, n, loop)) in enumerate(specs.items()):
        rows = [synth.channel(200 + 3 * k + c, n) for c in range(ch)]
        pcm[name] = (rows, n, loop)
        (src / name).write_bytes(oracle.wave_write16(rows, 32000, loop).tobytes())
    (src / "broken.wav").write_bytes(b"RIFF\x04\x00\x00\x00JUNK")
    (src / "ignored.txt").write_bytes(b"not audio")
    out = tmp_path / "dsp"
    r = subprocess.run([CLI, "-i", str(src), "-o", str(out), "--out-format", "dsp", "-r"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 3, r.stderr            # one file failed, the others were written (Batch.cs:39-43)
    assert "Error converting broken.wav" in r.stderr and "3 files converted, 1 failed" in r.stdout
    for name, (rows, n, loop) in pcm.items():
        coefs = np.stack([oracle.calculate_coefficients(p) for p in rows])
        adpcm = [oracle.encode(p, c) for p, c in zip(rows, coefs)]
        ctx = np.stack([np.array(oracle.gc_loop_context(a, oracle.decode(a, c, n), loop[0]), dtype=np.int16) for a, c in zip(adpcm, coefs)]) if loop else None
        want = oracle.dsp_write(adpcm, coefs, 32000, n, loop, ctx)
        got = (out / name).with_suffix(".dsp").read_bytes()
        assert got == want.tobytes(), name
    out2 = tmp_path / "hca"
    r = subprocess.run([CLI, "-i", str(src), "-o", str(out2), "--out-format", "hca", "--hcaquality", "Middle", "--keycode", "12345"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 3 and "2 files converted, 1 failed" in r.stdout, (r.stdout, r.stderr)   # not recursive: a.wav, b.WAV
    table = oracle.hca_key_tables(56, 12345)[1]
    for name in ("a.wav", "b.WAV"):
        rows, n, loop = pcm[name]
        info, frames = oracle.hca_encode(rows, 32000, quality=3, loop=loop)
        assert (out2 / name).with_suffix(".hca").read_bytes() == oracle.hca_write(info, frames, table, 56).tobytes(), name


def test_batch_directory_dsp_to_wav(tmp_path, oracle):
    src = tmp_path / "in"
    src.mkdir()
    rows = [synth.channel(260 + c, 20000) for c in range(2)]
    coefs = np.stack([oracle.calculate_coefficients(p) for p in rows])
    adpcm = [oracle.encode(p, c) for p, c in zip(rows, coefs)]
    (src / "x.dsp").write_bytes(oracle.dsp_write(adpcm, coefs, 32000, 20000).tobytes())
    out = tmp_path / "wav"
    r = subprocess.run([CLI, "-i", str(src), "-o", str(out), "--out-format", "wav"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "1 files converted, 0 failed" in r.stdout, (r.stdout, r.stderr)
    want = oracle.wave_write16([oracle.decode(a, c, 20000) for a, c in zip(adpcm, coefs)], 32000, None)
    assert (out / "x.wav").read_bytes() == want.tobytes()
