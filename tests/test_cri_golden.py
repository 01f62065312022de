"""Committed golden vectors of the CRI codecs (tests/golden/cri_golden.json, made by tests/golden/make_cri_golden.py).

They are frozen ORACLE outputs (see the generator's docstring): the CPU tests detect drift of the oracle or of the
synthetic generator, the GPU tests compare the CUDA path with the same hashes without running the oracle on the box."""
import hashlib
import json
import os

import numpy as np
import pytest

from vgaudio_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cri_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", GOLD["adx"], ids=lambda c: c["name"])
def test_oracle_adx_matches_golden(oracle, case):
    pcm = synth.channel(case["index"], case["n"], degenerate=False)
    assert sha(pcm) == case["pcm_sha256"], "synthetic generator drifted"
    adpcm, hist = oracle.adx_encode(pcm, 48000, case["frame_size"], case["version"], 0, case["type"], 2)
    assert int(hist) == case["history"] and sha(adpcm) == case["adpcm_sha256"]
    dec = oracle.adx_decode(adpcm, case["n"], 48000, 500, case["frame_size"], case["version"], hist, 0, case["type"])
    assert sha(dec) == case["decoded_sha256"]


@pytest.mark.parametrize("case", GOLD["hca"], ids=lambda c: c["name"])
def test_oracle_hca_matches_golden(oracle, case):
    chans = [synth.channel(case["first_index"] + c, case["n"], degenerate=False) for c in range(case["channels"])]
    assert sha(np.stack(chans)) == case["pcm_sha256"], "synthetic generator drifted"
    info, frames = oracle.hca_encode(chans, 48000, case["quality"], loop=tuple(case["loop"]) if case["loop"] else None)
    assert info.as_dict() == case["info"]
    assert sha(frames) == case["frames_sha256"]
    assert sha(oracle.hca_decode(info, frames)) == case["decoded_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["adx"], ids=lambda c: c["name"])
def test_gpu_adx_matches_golden(vg, case):
    pcm = synth.channel(case["index"], case["n"], degenerate=False)
    cfg = vg.criadx.CriAdxParameters(frame_size=case["frame_size"], version=case["version"], type=case["type"], filter=2)
    adpcm = vg.criadx.encode(pcm, cfg)
    assert sha(adpcm) == case["adpcm_sha256"], adpcm[:36].tobytes().hex() + " vs " + case["adpcm_head_hex"]
    if case["version"] == 4:
        assert cfg.history == case["history"]
    dcfg = vg.criadx.CriAdxParameters(frame_size=case["frame_size"], version=case["version"], type=case["type"],
                                      history=case["history"])
    assert sha(vg.criadx.decode(adpcm, case["n"], dcfg)) == case["decoded_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["hca"], ids=lambda c: c["name"])
def test_gpu_hca_matches_golden(vg, case):
    chans = [synth.channel(case["first_index"] + c, case["n"], degenerate=False) for c in range(case["channels"])]
    loop = case["loop"]
    cfg = vg.crihca.CriHcaParameters(quality=case["quality"], looping=bool(loop), loop_start=loop[0] if loop else 0,
                                     loop_end=loop[1] if loop else 0)
    info, frames = vg.crihca.encode(chans, 48000, cfg)
    assert info.as_dict() == case["info"]
    assert sha(frames) == case["frames_sha256"], frames[0].tobytes().hex()[:96] + " vs " + case["frame0_hex"]
    assert sha(np.stack(vg.crihca.decode(info, frames))) == case["decoded_sha256"]
