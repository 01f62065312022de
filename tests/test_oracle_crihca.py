"""CPU tests of the CRI HCA oracle (oracle/crihca.c).

Pins taken from the reference's own tests (via tests/golden/hca_tables.json, extracted by make_hca_tables.py):
  * every table the codec uses is the reference's test literal itself (exact);  the runtime formulas of
    CriHcaTables.cs are re-evaluated here and compared with those literals so a wrong index/formula would show;
  * MDCT shuffle tables exact, sin/cos to 14 decimals (src/VGAudio.Tests/Utilities/MdctTests.cs:18-59).
PARITY UNPINNED for frame bytes: the reference never runs its HCA encoder/decoder/packer/CRC in a test, so those are
held to self-consistency (decode(encode(x)) tracks x, bit budget respected, CRC check value, MDCT/IMDCT TDAC)."""
import json
import math
import os

import numpy as np
import pytest

from vgaudio_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hca_tables.json")))


def test_generated_tables_follow_the_reference_formulas():
    g = GOLD["generated"]
    # CriHcaTables.cs:55-66 (libm pow may differ in the last ulp from .NET's: compare to 1e-15 relative; the oracle
    # itself uses the literals, not these formulas)
    deq = [math.sqrt(128) * math.pow(math.pow(2, 53.0 / 128), x - 63) for x in range(64)]
    assert np.allclose(g["DequantizerScalingTable"], deq, rtol=1e-14, atol=0)
    assert np.allclose(g["QuantizerScalingTable"], [1 / v for v in deq], rtol=1e-14, atol=0)
    rmax = [x if x < 8 else (1 << (x - 4)) - 1 for x in range(16)]
    assert g["ResolutionMaxValue"] == rmax
    assert g["QuantizerInverseStepSize"] == [v + 0.5 for v in rmax]
    assert g["QuantizerStepSize"] == [0 if x == 0 else 1 / (rmax[x] + 0.5) for x in range(16)]
    assert g["IntensityRatioTable"] == [(28 - x * 2) / 14.0 for x in range(15)]
    assert g["IntensityRatioBoundsTable"] == [(27 - x * 2) / 14.0 for x in range(14)]
    conv = [math.pow(math.pow(2, 53.0 / 128), x - 64) if 1 < x < 127 else 0 for x in range(128)]
    assert np.allclose(g["ScaleConversionTable"], conv, rtol=1e-14, atol=0)


def test_mdct_tables_match_prebuilt(oracle):  # MdctTests.cs:18-59 vs PreBuiltMdctTables.cs
    m = GOLD["mdct"]
    for bits in range(8):
        s, c, sh = oracle.hca_mdct_tables(bits)
        assert sh.tolist() == m["ShuffleTables"][bits]
        assert np.allclose(s, m["SinTables"][bits], rtol=0, atol=1e-14)
        assert np.allclose(c, m["CosTables"][bits], rtol=0, atol=1e-14)


def test_crc16_check_value(oracle):
    assert oracle.crc16(b"123456789") == 0xFEE8   # CRC-16 poly 0x8005, init 0, no reflection
    assert oracle.crc16(b"") == 0


def test_mdct_matches_direct_dct4_and_imdct_reconstructs(oracle):
    rng = np.random.default_rng(3)
    blocks = rng.uniform(-1, 1, (12, 128))
    spec = oracle.hca_mdct(blocks)
    # direct O(N^2) evaluation (Mdct.Dct4Slow :214-227) of the windowed/folded input
    w = np.array(GOLD["unpacked"]["MdctWindow"])
    prev = np.zeros(128)
    scale = math.sqrt(2.0 / 128)
    for k in range(12):
        x = blocks[k]
        d = np.zeros(128)
        for i in range(64):
            d[i] = w[64 - i - 1] * -x[64 + i] - w[64 + i] * x[64 - i - 1]
            d[64 + i] = w[i] * prev[i] - w[127 - i] * prev[127 - i]
        n = np.arange(128)
        direct = np.array([np.sum(np.cos(math.pi / 128 * (kk + 0.5) * (n + 0.5)) * d) * scale for kk in range(128)])
        assert np.allclose(spec[k], direct, atol=1e-12)
        prev = x
    # time-domain alias cancellation: IMDCT of the MDCT gives the input back one block late
    rec = oracle.hca_imdct(spec)
    assert np.allclose(rec[1:], blocks[:-1], atol=1e-6)   # the window itself is float32 data: PR holds to ~1e-7


@pytest.mark.parametrize("nch,quality", [(1, 2), (2, 2), (1, 1), (2, 5), (1, 5), (2, 3), (6, 2)])
def test_stream_parameters_and_round_trip(oracle, nch, quality):
    n = 48000
    chans = [synth.reference_sine(n, f, 48000) // 2 for f in (261.63, 329.63, 392, 523.25, 659.25, 783.99)[:nch]]
    info, frames = oracle.hca_encode(chans, quality=quality)
    assert frames.shape == (info.frame_count, info.frame_size)
    assert info.frame_count == (n + 128 + 1023) // 1024
    pcm_bitrate = 48000 * nch * 16
    ratio = {1: 4, 2: 6, 3: 8, 4: 10 if nch == 1 else 12, 5: 12 if nch == 1 else 16}[quality]
    assert info.frame_size == (pcm_bitrate // ratio) * 1024 // 48000 // 8      # CriHcaEncoder.cs:288-328
    assert (frames[:, 0] == 0xFF).all() and (frames[:, 1] == 0xFF).all()        # sync word
    for f in frames:                                                          # checksum of every frame
        assert oracle.crc16(f.tobytes()) == 0                                  # CRC over data+crc is zero
    dec = oracle.hca_decode(info, frames)
    for c in range(nch):
        err = dec[c, 2048:-2048].astype(float) - chans[c][2048:-2048]
        limit = 60 if quality <= 3 else 2500   # sines at high/middle quality are nearly transparent
        assert np.sqrt((err ** 2).mean()) < limit, (c, np.sqrt((err ** 2).mean()))


def test_c4_configuration_numbers(oracle):
    """48 kHz High -> mono 128000 bit/s, 341-byte frames; stereo 682; 1407 frames per 30 s."""
    p = oracle.HcaParams(2, 0, 0, 1, 48000, 1440000, 0, 0, 0)
    info = oracle.hca_init(p)
    assert (info.bitrate, info.frame_size, info.frame_count, info.total_band_count, info.base_band_count) == (128000, 341, 1407, 128, 128)
    assert (info.inserted_samples, info.appended_samples, info.hfr_group_count, info.stereo_band_count) == (128, 640, 0, 0)
    p2 = oracle.HcaParams(2, 0, 0, 2, 48000, 1440000, 0, 0, 0)
    assert oracle.hca_init(p2).frame_size == 682


def test_silence_and_short_inputs(oracle):
    for n in (1, 127, 128, 1023, 1024, 1025, 5000):
        info, frames = oracle.hca_encode([np.zeros(n, dtype=np.int16)])
        assert frames.shape[0] == (n + 128 + 1023) // 1024
        dec = oracle.hca_decode(info, frames)
        assert not dec.any()
