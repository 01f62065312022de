"""GPU parity tests of the time-parallel GC-ADPCM encoder (speculate -> verify -> splice, gc_encode.cu) against the CPU
oracle's plain serial loop (GcAdpcmEncoder.cs:30-43).  Bit-exact.  VGB_GC_SEGMENTS forces the segment count, so the same
inputs run as 1 segment (the serial loop), a few, and the maximum."""
import ctypes as C
import os

import numpy as np
import pytest

from vgaudio_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture
def segments():
    """force(n) sets the segment count (None: the library's own choice); the shortest segment is lowered to 256 frames
    for the whole test so that short inputs are really cut (the default of 4096 is sized for the run-on tail)."""
    saved = {k: os.environ.get(k) for k in ("VGB_GC_SEGMENTS", "VGB_GC_MIN_SEG_FRAMES")}
    os.environ["VGB_GC_MIN_SEG_FRAMES"] = "256"

    def force(n):
        if n is None:
            os.environ.pop("VGB_GC_SEGMENTS", None)
        else:
            os.environ["VGB_GC_SEGMENTS"] = str(n)

    yield force
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def splice_stats(vg):
    from vgaudio_b200 import _native as N

    out = (C.c_uint64 * 4)()
    N.check(vg.lib.vgb_gcadpcm_debug_splice_stats(out, 4))
    return {"segments": int(out[0]), "runon_frames": int(out[1]), "cascade_frames": int(out[2]), "cascade_boundaries": int(out[3])}


def white_full_scale(n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-32768, 32768, n).astype(np.int16)


def nyquist_square(n, period=2):
    return np.where((np.arange(n) // (period // 2)) % 2 == 0, 32767, -32768).astype(np.int16)


def check_against_oracle(vg, oracle, chans, coefs=None, configs=None):
    got_coefs, adpcm = vg.gcadpcm.encode_batch(chans, coefs, configs)
    for c, pcm in enumerate(chans):
        co = oracle.calculate_coefficients(pcm) if coefs is None else coefs[c]
        assert np.array_equal(got_coefs[c], co), c
        if configs is None:
            want = oracle.encode(pcm, co)
        else:
            p = configs[c]
            want = oracle.encode(pcm, co, p.sample_count, p.history1, p.history2)
        first = np.flatnonzero(np.frombuffer(adpcm[c].tobytes(), np.uint8) != want)
        assert first.size == 0, f"channel {c}: first differing byte {first[0]} (frame {first[0] // 8}) of {len(want)}"


@pytest.mark.parametrize("seg", [1, 2, 5, 16, 64, None])
def test_segmented_encode_is_bit_exact(vg, oracle, segments, seg):
    segments(seg)
    n = 14 * 4100 + 9  # 4101 frames: up to 16 segments of >= 256 frames
    chans = [synth.channel(i, n) for i in range(12)]
    check_against_oracle(vg, oracle, chans)
    st = splice_stats(vg)
    if seg is not None:
        assert st["segments"] == seg
    if seg == 1:
        assert st["runon_frames"] == 0 and st["cascade_frames"] == 0


def test_channels_that_never_relock_fall_back_to_the_serial_loop(vg, oracle, segments):
    """White full-scale noise, Nyquist squares and a hostile-coefficient channel do not (reliably) re-lock inside a
    segment: the cascade must carry the true chain across segment ends, and the bytes stay exact."""
    segments(8)
    n = 14 * 2600 + 3
    chans = [white_full_scale(n, 1), nyquist_square(n, 2), nyquist_square(n, 8), white_full_scale(n, 2),
             synth.channel(7, n), (white_full_scale(n, 3) // 2).astype(np.int16), synth.channel(1, n), white_full_scale(n, 4)]
    check_against_oracle(vg, oracle, chans)
    rng = np.random.default_rng(17)
    # resonant / unstable predictors: chains from different histories need not meet at all
    coefs = np.stack([rng.choice(np.array([4095, -2047, 4000, -1900, 3800, -2000, 2048, -1024], dtype=np.int16), 16) for _ in chans])
    check_against_oracle(vg, oracle, chans, coefs=coefs)
    st = splice_stats(vg)
    assert st["segments"] == 8
    assert st["runon_frames"] > 0


def test_cascade_runs_when_runons_do_not_splice(vg, oracle, segments):
    """Unstable predictors (|pole| > 1): a wrong history never decays, every boundary is left to the cascade."""
    segments(6)
    n = 14 * 1700
    chans = [white_full_scale(n, 10 + i) for i in range(4)]
    coefs = np.tile(np.array([4300, -2300] * 8, dtype=np.int16), (4, 1))
    check_against_oracle(vg, oracle, chans, coefs=coefs)
    st = splice_stats(vg)
    total = 4 * 1700
    assert 0 < st["runon_frames"] + st["cascade_frames"] <= 2 * total


def test_ragged_batch_with_history_and_sample_counts(vg, oracle, segments):
    segments(7)
    lens = [14 * 3000, 14 * 3000 + 13, 5, 14 * 255, 14 * 256, 14 * 257 + 1, 14 * 1792, 14 * 1793 + 6, 14 * 5000 + 2, 0, 14 * 2049]
    chans = [synth.channel(30 + i, max(L, 1))[:L] for i, L in enumerate(lens)]
    P = vg.gcadpcm.GcAdpcmParameters
    configs = [P(-1, 0, 0), P(14 * 2999 + 5, 100, -100), P(-1, 7, 8), P(-1, -32768, 32767), P(14 * 200, 1, 2), P(-1, 0, 0),
               P(14 * 1792, 3000, 2999), P(-1, -5, -6), P(14 * 4097 + 1, 0, 0), P(-1, 0, 0), P(-1, 12345, -12345)]
    rng = np.random.default_rng(23)
    coefs = rng.integers(-3000, 3000, (len(chans), 16)).astype(np.int16)
    check_against_oracle(vg, oracle, chans, coefs=coefs, configs=configs)


def test_device_resident_slices_carry_history(vg, oracle, segments):
    """The default heuristic on a long single channel (many segments, one warp row) and the decoder round trip."""
    segments(None)
    n = 14 * 20000 + 11
    for idx in (4, 12):  # the loud channels that need the longest run-ons
        pcm = synth.channel(idx, n)
        coefs = oracle.calculate_coefficients(pcm)
        got = vg.gcadpcm.encode(pcm, coefs)
        assert got.tobytes() == oracle.encode(pcm, coefs).tobytes(), idx
        st = splice_stats(vg)
        assert st["segments"] > 1
        dec = vg.gcadpcm.decode(got, coefs, vg.gcadpcm.GcAdpcmParameters(n))
        assert np.array_equal(dec, oracle.decode(got, coefs, n))
