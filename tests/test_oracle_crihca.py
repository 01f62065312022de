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


# ---- looping (CriHcaEncoder.Initialize :89-99, CalculateLoopInfo :383-398, CalculateHeaderSize :400-418 and the
# streaming front end :126-272).  The reference has no test for it: the restatement is held to the format's invariants.

@pytest.mark.parametrize("loop", [(5000, 25000), (0, 30000), (1024, 20000), (29000, 30000), (29900, 30000), (100, 500),
                                  (2048, 2049), (12345, 23456)])
def test_looping_stream_invariants(oracle, loop):
    n = 30000
    x = synth.reference_sine(n, 440, 48000)
    info, frames = oracle.hca_encode([x], 48000, 2, loop=loop)
    d = info.as_dict()
    assert d["looping"] == 1 and d["sample_count"] == min(loop[1], n)
    # HcaInfo.LoopStartSample / LoopEndSample (HcaInfo.cs:35-36) give the requested loop points back
    assert d["loop_start_frame"] * 1024 + d["pre_loop_samples"] - d["inserted_samples"] == loop[0]
    assert (d["loop_end_frame"] + 1) * 1024 - d["post_loop_samples"] - d["inserted_samples"] == loop[1]
    # the loop start lands one subframe (the codec delay) into a frame, and that frame on a 2048-byte file boundary
    assert d["pre_loop_samples"] == 128
    assert (d["header_size"] + d["frame_size"] * d["loop_start_frame"]) % 2048 == 0
    assert d["frame_count"] * 1024 == d["inserted_samples"] + d["appended_samples"] + min(-(-d["sample_count"] // 128) * 128, n) + 256
    # the stream decodes to the input up to the loop end ...
    y = oracle.hca_decode(info, frames)[0].astype(np.float64)
    sc = d["sample_count"]
    assert np.sqrt(((y[:sc] - x[:sc]) ** 2).mean()) < 150
    # ... and what follows the loop end inside the last frames is the audio of the loop start (seamless wrap-around):
    # decode with a sample count that exposes the appended audio
    import copy
    wide = copy.copy(info)
    extra = min(128, d["frame_count"] * 1024 - d["inserted_samples"] - sc)
    wide.sample_count = sc + extra
    z = oracle.hca_decode(wide, frames)[0].astype(np.float64)
    if extra >= 64 and loop[0] + extra <= n:
        wrap = np.sqrt(((z[sc:sc + extra] - x[loop[0]:loop[0] + extra]) ** 2).mean())
        assert wrap < 2000  # the phase jump at the loop end costs some coding noise
        if sc + extra <= n and abs((loop[1] - loop[0]) % 109 - 54) < 40:  # 440 Hz at 48 kHz: period ~109 samples
            assert wrap < 0.5 * np.sqrt(((z[sc:sc + extra] - x[sc:sc + extra]) ** 2).mean())


def test_non_looping_streaming_front_end_equals_plain_windows(oracle):
    """For a non-looping stream the streaming front end reduces to: frame k = k-th 1024-sample window, then zeros.
    The MDCT tap (plain windows) followed by the rest of the frame pipeline is what the golden frames were made from
    before the front end was restated; the frames must not have moved."""
    x = synth.channel(3, 5000)
    info, frames = oracle.hca_encode([x], 48000, 2)
    assert info.frame_count == (5000 + 128 + 1023) // 1024 and info.inserted_samples == 128
    y = oracle.hca_decode(info, frames)[0].astype(np.float64)
    assert np.sqrt(((y - x) ** 2).mean()) < 0.25 * np.sqrt((x.astype(np.float64) ** 2).mean())


def test_ath_curve_stream_round_trips(oracle):
    """HcaInfo.UseAthCurve (HcaInfo.cs:38): with the curve, resolutions come from athCurve[band] + noise level on both
    sides (CriHcaPacking.cs:79-95).  ScaleAthCurve (CriHcaFrame.cs:60-84) resamples the 41856 Hz table: at that rate the
    index advances by 5.1 entries per band, above ~43 kHz * 654/... the tail is 0xff.  The helper stream must decode
    close to its input with the flag and differently without it."""
    from vgaudio_b200 import synth  # data generation only

    pcm = [synth.channel(41, 9000, degenerate=False)]
    info, frames = oracle.hca_encode(pcm, 48000, 2, ath=True)
    assert info.use_ath_curve == 1
    assert oracle.hca_unpack_ok(info, frames)
    dec = oracle.hca_decode(info, frames)
    ref = np.asarray(pcm[0], dtype=np.float64)
    assert np.sqrt(((dec[0] - ref) ** 2).mean()) < 0.35 * np.sqrt((ref ** 2).mean())
    plain_info, plain_frames = oracle.hca_encode(pcm, 48000, 2)
    assert plain_info.use_ath_curve == 0
    assert not np.array_equal(plain_frames, frames)
