"""The kit that closes "parity unpinned" on a .NET machine (bindings/csharp/ParityHarness.cs `vectors` mode): the dump
tool must write a complete, self-consistent vector set, and the set written by the CUDA path must be byte-identical to
the one written by the oracle (so either can be shipped)."""
import filecmp
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dump(tmp, *flags):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dump_vectors.py"), str(tmp), *flags], check=True, cwd=ROOT)
    rows = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(tmp, "manifest.tsv")) if not ln.startswith("#")]
    return rows


def test_dump_vectors_manifest_is_complete(tmp_path):
    rows = _dump(tmp_path, "--oracle")
    assert {r[0] for r in rows} == {"gcadpcm", "gcadpcm_decode", "criadx", "crihca", "wave_to_dsp", "wave_to_adx", "wave_to_hca"}
    assert len(rows) >= 75
    for codec, name, params, inputs, output in rows:
        for f in inputs.split(",") + [output]:
            assert os.path.getsize(os.path.join(tmp_path, f)) >= 0, (codec, name, f)
        assert all("=" in kv for kv in params.split(",") if kv)
    # the C# side exists and names every codec of the manifest
    harness = open(os.path.join(ROOT, "bindings", "csharp", "ParityHarness.cs")).read()
    for codec in ("gcadpcm", "gcadpcm_decode", "criadx", "crihca", "wave_to_dsp", "wave_to_adx", "wave_to_hca"):
        assert f'"{codec}"' in harness
    tool = open(os.path.join(ROOT, "bindings", "csharp", "DspToolB200.cs")).read()
    for member in ("EncodeChannel", "DspCorrelateCoefs", "DspEncodeFrame", "DecodeChannel", "DecodeAdpcm"):  # IDspTool.cs:5-12
        assert member in tool


@pytest.mark.gpu
def test_cuda_vectors_equal_oracle_vectors(tmp_path):
    a, b = tmp_path / "cuda", tmp_path / "oracle"
    rows_a = _dump(a)
    rows_b = _dump(b, "--oracle")
    assert rows_a == rows_b
    for _, name, _, _, output in rows_a:
        assert filecmp.cmp(os.path.join(a, output), os.path.join(b, output), shallow=False), name
