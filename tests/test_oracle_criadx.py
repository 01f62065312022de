"""CPU tests of the CRI ADX oracle (oracle/criadx.c).

PARITY UNPINNED: the reference has no ADX test and no golden ADX data and cannot be executed in
this environment, so these tests hold the oracle to the codec's own invariants: the high-pass coefficient values
SURVEY.md derived independently, encoder reconstruction == decoder, header/frame layout, padding behaviour."""
import numpy as np
import pytest

from vgaudio_b200 import synth


def test_highpass_coefficients(oracle):
    # 48 kHz -> (7400, -3342); 44.1 kHz -> (7334, -3283), computed there with Python doubles
    assert oracle.adx_coefficients(500, 48000).tolist() == [7400, -3342]
    assert oracle.adx_coefficients(500, 44100).tolist() == [7334, -3283]


@pytest.mark.parametrize("typ", [2, 3, 4])
@pytest.mark.parametrize("version", [3, 4])
def test_encoder_reconstruction_equals_decoder(oracle, typ, version):
    pcm = synth.channel(6, 32 * 200)
    adpcm, hist = oracle.adx_encode(pcm, 48000, 18, version, 0, typ, 2)
    dec = oracle.adx_decode(adpcm, len(pcm), 48000, 500, 18, version, hist, 0, typ)
    coefs = oracle.adx_coefficients(500, 48000) if typ != 2 else np.array([0x1CC0, -3328], dtype=np.int16)
    window = np.zeros(34, dtype=np.int16)
    if version == 4:
        window[0] = window[1] = pcm[0]
    for f in range(200):
        window[2:] = pcm[32 * f: 32 * f + 32]
        frame = oracle.adx_encode_frame(window, coefs, 32, typ, version)
        if typ == 2:
            frame[0] |= 2 << 5
        assert frame.tobytes() == adpcm[18 * f: 18 * f + 18].tobytes(), f
        assert np.array_equal(window[2:], dec[32 * f: 32 * f + 32]), f
        window[0], window[1] = window[32], window[33]


def test_frame_layout_and_sizes(oracle):
    pcm = synth.reference_sine(1000, 440, 48000)
    adpcm, hist = oracle.adx_encode(pcm)
    assert len(adpcm) == ((1000 + 31) // 32) * 18
    assert hist == int(pcm[0])                      # CriAdxCodec.cs:73 (version 4, no padding)
    scales = (adpcm[0::18].astype(int) << 8 | adpcm[1::18]) & 0x1FFF
    assert (scales <= 0x0FFF).all()
    assert (adpcm[0::18] >> 5 == 0).all()           # filter bits only for the Fixed type
    dec = oracle.adx_decode(adpcm, 1000, history=hist)
    # the last (zero-padded) frame ends in a step to silence and needs a coarse scale: leave it out
    assert np.abs(dec[64:960].astype(int) - pcm[64:960]).max() < 600


@pytest.mark.parametrize("padding", [1, 13, 31, 32, 33, 40, 64, 100])
def test_padding_skips_whole_frames_and_zero_fills(oracle, padding):
    pcm = synth.channel(8, 500)
    adpcm, hist = oracle.adx_encode(pcm, padding=padding)
    assert hist == 0                                 # history is only seeded when padding == 0
    assert len(adpcm) == ((500 + padding + 31) // 32) * 18
    skipped = padding // 32
    assert not adpcm[: skipped * 18].any()           # fully padded frames are never written (`continue`)
    dec = oracle.adx_decode(adpcm, 500, padding=padding)
    assert len(dec) == 500


def test_exponential_scale_is_a_power_of_two_exponent(oracle):
    pcm = synth.channel(9, 3200)
    adpcm, _ = oracle.adx_encode(pcm, type=4)
    scales = (adpcm[0::18].astype(int) << 8 | adpcm[1::18]) & 0x1FFF
    assert (scales <= 12).all()
