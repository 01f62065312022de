"""CPU tests of the container layer: the oracle restatement (oracle/containers.c) against the pins
the reference's own tests hold - build -> parse round trips (src/VGAudio.Tests/Containers/WaveTests.cs:9-55, DspTests.cs:9-19,
BuildParseTests.cs:9-16) - and the product's host-only entry points (parsers, sizes, key schedules) against the oracle.
No GPU work here."""
import ctypes as C
import struct

import numpy as np
import pytest

from vgaudio_b200 import synth


def _sine_channels(n_ch, n, rate=48000):
    # GenerateAudio.GeneratePcmSineWave: channel c is a sine of the c-th test frequency (GenerateAudio.cs:14-33)
    freqs = [261.63, 329.63, 392.0, 523.25, 659.25, 783.99, 1046.5, 130.81]
    return [synth.reference_sine(n, freqs[c % 8], rate) for c in range(n_ch)]


def _wave8(channels, rate=22050):
    """An 8-bit PCM WAVE image (WaveWriter with WaveCodec.Pcm8Bit writes the same layout, WaveWriter.cs:73-110)."""
    ch, n = len(channels), len(channels[0])
    data = np.stack(channels, axis=1).astype(np.uint8).tobytes()
    fmt = struct.pack("<HHIIHH", 1, ch, rate, rate * ch, ch, 8)
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", len(data)) + data
    return np.frombuffer(b"RIFF" + struct.pack("<I", len(body)) + body, dtype=np.uint8), [((c.astype(np.int32) - 0x80) << 8).astype(np.int16) for c in channels]


@pytest.mark.parametrize("n_ch", [1, 2, 8])
@pytest.mark.parametrize("loop", [None, (0, 1000), (123, 40000)])
def test_wave_build_parse_equal(oracle, n_ch, loop):
    # WaveTests.WavePcm16BuildAndParseEqual / WavePcm16LoopedBuildAndParseEqual
    pcm = _sine_channels(n_ch, 40000)
    w = oracle.wave_write16(pcm, 48000, loop)
    st, info = oracle.wave_parse(w)
    assert st == 0
    assert (info.channel_count, info.sample_rate, info.bits_per_sample, info.sample_count) == (n_ch, 48000, 16, 40000)
    assert bool(info.looping) == (loop is not None)
    if loop:
        assert (info.loop_start, info.loop_end) == loop
    got = oracle.wave_read(w, info)
    assert all(np.array_equal(a, b) for a, b in zip(got, pcm))


def test_wave_8bit_reads_through_pcm8_codec(oracle):
    w, want = _wave8([np.arange(300) % 256, (np.arange(300) * 7) % 256])
    st, info = oracle.wave_parse(w)
    assert st == 0 and info.bits_per_sample == 8 and info.sample_count == 300
    got = oracle.wave_read(w, info)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


def _mutations(oracle):
    good = oracle.wave_write16(_sine_channels(2, 500), 44100, (5, 400)).copy()
    out = {"good": good}
    m = good.copy(); m[0:4] = np.frombuffer(b"RIFX", np.uint8); out["not_riff"] = m
    m = good.copy(); m[8:12] = np.frombuffer(b"WAVX", np.uint8); out["not_wave"] = m
    m = good.copy(); m[12:16] = np.frombuffer(b"fmx ", np.uint8); out["no_fmt"] = m
    m = good.copy(); m[20:22] = (3, 0); out["not_pcm"] = m
    m = good.copy(); m[34:36] = (24, 0); out["bits"] = m
    m = good.copy(); m[32:34] = (3, 0); out["block_align"] = m
    m = good.copy(); m[22:24] = (0, 0); out["zero_channels"] = m
    out["truncated_header"] = good[:30].copy()
    out["truncated_data"] = good[: good.size - 101].copy()
    m = good.copy()
    i = bytes(m).find(b"data"); m[i:i + 4] = np.frombuffer(b"dat_", np.uint8); out["no_data"] = m
    j = bytes(good).find(b"smpl")
    m = good.copy(); m[j + 8 + 36 + 12: j + 8 + 36 + 16] = np.frombuffer(struct.pack("<i", 100000), np.uint8); out["loop_past_end"] = m
    m = good.copy(); m[j + 8 + 36 + 12: j + 8 + 36 + 16] = np.frombuffer(struct.pack("<i", 2), np.uint8); out["loop_end_before_start"] = m
    return out


def test_wave_parse_errors_and_product_parser_matches_oracle(oracle, vg):
    """ValidateWaveFile's checks (WaveReader.cs:71-95) and the product's host parser against the oracle, field by field."""
    from vgaudio_b200 import _native as N

    muts = _mutations(oracle)
    want_fail = {"not_riff", "not_wave", "no_fmt", "not_pcm", "bits", "block_align", "zero_channels", "truncated_header", "no_data",
                 "loop_past_end"}
    for name, img in muts.items():
        st, info = oracle.wave_parse(img)
        assert (st != 0) == (name in want_fail), (name, st)
        pinfo = N.VgbWaveInfo()
        pst = vg.lib.vgb_wave_parse(img.ctypes.data, img.size, C.byref(pinfo))
        assert (pst != 0) == (st != 0), (name, pst, vg.lib.vgb_last_error())
        if st == 0:
            for f, _ in N.VgbWaveInfo._fields_:
                assert getattr(pinfo, f) == getattr(info, f), (name, f)
        else:
            assert pst == N.VGB_E_DATA
    st, info = oracle.wave_parse(muts["truncated_data"])
    assert st == 0 and info.sample_count == (muts["good"].size - 101 - info.data_offset) // 4   # ReadBytes returns what is left
    st, info = oracle.wave_parse(muts["loop_end_before_start"])
    assert st == 0 and not info.looping and info.loop_end == 0                                    # Looping = LoopEnd > LoopStart


@pytest.mark.parametrize("n_ch", [1, 2, 5])
@pytest.mark.parametrize("n", [14 * 300, 14 * 1024 + 5, 20000])
@pytest.mark.parametrize("loop", [None, (100, 3000)])
def test_dsp_build_parse_equal(oracle, vg, n_ch, n, loop):
    """DspTests.DspBuildAndParseEqual: write -> read returns the same channels, coefficients, contexts and loop points; the
    product's host-only vgb_dsp_parse / vgb_dsp_file_size agree with the oracle."""
    from vgaudio_b200 import _native as N

    pcm = _sine_channels(n_ch, n)
    coefs = np.stack([oracle.calculate_coefficients(p) for p in pcm])
    adpcm = [oracle.encode(p, c) for p, c in zip(pcm, coefs)]
    ctx = None
    if loop:
        ctx = np.stack([np.array(oracle.gc_loop_context(a, oracle.decode(a, c, n), loop[0]), dtype=np.int16) for a, c in zip(adpcm, coefs)])
    f = oracle.dsp_write(adpcm, coefs, 32000, n, loop, ctx, trim_file=False)
    st, info = oracle.dsp_parse(f)
    assert st == 0
    assert (info.sample_count, info.sample_rate, info.channel_count, bool(info.looping)) == (n, 32000, n_ch, loop is not None)
    if loop:
        assert (info.loop_start, info.loop_end) == loop
    rows = oracle.dsp_read_data(f, info)
    for c in range(n_ch):
        assert rows[c].tobytes() == adpcm[c].tobytes()
        assert list(info.coefs[c]) == coefs[c].tolist()
        assert info.start_ctx[c][0] == adpcm[c][0]
        if loop:
            assert list(info.loop_ctx[c]) == ctx[c].tolist()
    pinfo = N.VgbDspInfo()
    assert vg.lib.vgb_dsp_parse(f.ctypes.data, f.size, C.byref(pinfo)) == 0
    for name in ("sample_count", "nibble_count", "sample_rate", "looping", "format", "start_address", "end_address", "current_address",
                 "channel_count", "frames_per_interleave", "loop_start", "loop_end"):
        assert getattr(pinfo, name) == getattr(info, name), name
    for c in range(n_ch):
        assert list(pinfo.coefs[c]) == list(info.coefs[c]) and list(pinfo.loop_context[c]) == list(info.loop_ctx[c])
        assert list(pinfo.start_context[c]) == list(info.start_ctx[c]) and pinfo.gain[c] == info.gain[c]
    d = N.VgbDspDesc(n_ch, 32000, n, int(loop is not None), loop[0] if loop else 0, loop[1] if loop else 0, 0, 0, 1)
    assert vg.lib.vgb_dsp_file_size(C.byref(d)) == f.size


def test_dsp_parse_rejects_what_the_reference_rejects(oracle, vg):
    from vgaudio_b200 import _native as N

    pcm = _sine_channels(1, 1400)
    co = oracle.calculate_coefficients(pcm[0])
    f = oracle.dsp_write([oracle.encode(pcm[0], co)], co[None], 32000, 1400)
    for mutate in (lambda m: m.__setitem__(slice(4, 8), 0), lambda m: m.__setitem__(15, 1)):
        m = f.copy(); mutate(m)
        assert oracle.dsp_parse(m)[0] != 0
        assert vg.lib.vgb_dsp_parse(m.ctypes.data, m.size, C.byref(N.VgbDspInfo())) == N.VGB_E_DATA
    short = f[:0x60 + 10].copy()
    assert oracle.dsp_parse(short)[0] != 0 and vg.lib.vgb_dsp_parse(short.ctypes.data, short.size, C.byref(N.VgbDspInfo())) == N.VGB_E_DATA


def test_adx_keys_and_crypt(oracle, vg):
    from vgaudio_b200 import _native as N

    for code in (1, 2, 0x1234567890AB, (1 << 42) - 1):
        k = N.VgbAdxKey()
        assert vg.lib.vgb_adx_key_from_code(code, C.byref(k)) == 0
        assert (k.seed, k.mult, k.inc) == oracle.adx_key(key_code=code)
        assert k.mult & 1 and k.inc & 1 and k.seed < 0x8000
    for s in ("", "a", "karaage", "2394509830"):
        k = N.VgbAdxKey()
        assert vg.lib.vgb_adx_key_from_string(s.encode(), C.byref(k)) == 0
        assert (k.seed, k.mult, k.inc) == oracle.adx_key(key_string=s)
        assert all(0x4000 <= v < 0x8000 for v in (k.seed, k.mult, k.inc))   # entries of the prime table
    pcm = _sine_channels(2, 5000)
    audio = [oracle.adx_encode(p)[0] for p in pcm]
    audio[1][18 * 3: 18 * 4] = 0                                             # an empty frame stays untouched (FrameNotEmpty)
    key = oracle.adx_key(key_string="karaage")
    enc = oracle.adx_crypt(audio, key, 8, 18)
    assert any(not np.array_equal(a, b) for a, b in zip(enc, audio))
    assert not enc[1][18 * 3: 18 * 4].any()
    dec = oracle.adx_crypt(enc, key, 8, 18)                                  # type 8 is an XOR stream: its own inverse
    assert all(np.array_equal(a, b) for a, b in zip(dec, audio))
    enc9 = oracle.adx_crypt(audio, key, 9, 18)
    assert all((e[0::18] & 0xE0 == 0).all() for e in enc9)                   # type 9 masks the scale's top bits (:34)


def test_hca_key_tables_and_crypt(oracle, vg):
    for kt, code in ((0, 0), (1, 0), (56, 1), (56, 0xCC55463930DBE1AB), (56, 12345678901234567)):
        dec, enc = oracle.hca_key_tables(kt, code)
        assert sorted(dec.tolist()) == list(range(256)) and np.array_equal(dec[enc], np.arange(256))
        assert dec[0] == 0 and dec[255] == 255                               # 0x00 and 0xff map to themselves (ShuffleTable)
        pd, pe = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
        assert vg.lib.vgb_hca_key_tables(kt, code, pd.ctypes.data, pe.ctypes.data) == 0
        assert np.array_equal(pd, dec) and np.array_equal(pe, enc)
    info, frames = oracle.hca_encode(_sine_channels(2, 6000), 48000)[:2]
    fs = info.frame_size
    dec, enc = oracle.hca_key_tables(56, 0xCC55463930DBE1AB)
    e = oracle.hca_crypt_frames(frames, fs, enc)
    for k in range(info.frame_count):
        assert oracle.crc16(bytes(e[k * fs:(k + 1) * fs])) == 0              # a frame with its CRC appended checks to zero
    assert np.array_equal(oracle.hca_crypt_frames(e, fs, dec), np.asarray(frames).ravel())


def test_adx_and_hca_file_geometry_matches_oracle(oracle, vg):
    """vgb_adx_file_size (host only) against the oracle writer's size over versions / loops / channel counts, and the HCA
    header the oracle writes: chunk ids, sizes and the header CRC."""
    from vgaudio_b200 import _native as N

    for ch in (1, 2, 6):
        for version in (3, 4):
            for loop in (None, (1000, 9000), (31, 33)):
                for trim in (True, False):
                    n = 10000
                    spf = 32
                    align = 0
                    if loop:
                        mult = spf * 2 if ch == 1 else spf
                        align = (-loop[0]) % mult
                    audio = [np.zeros(oracle.lib().vgo_adx_encoded_byte_count(n, align, 18), np.uint8) for _ in range(ch)]
                    f = oracle.adx_write(audio, [0] * ch, 48000, n, loop, align, 18, version, trim_file=trim)
                    d = N.VgbAdxDesc(ch, 48000, n, int(loop is not None), loop[0] if loop else 0, loop[1] if loop else 0, align, 18, version, 3, 500, 0,
                                     int(not trim))
                    assert vg.lib.vgb_adx_file_size(C.byref(d)) == f.size, (ch, version, loop, trim)
                    assert f[0] == 0x80 and f[1] == 0 and bytes(f[int(f[2]) * 256 + int(f[3]) - 2:][:6]) == b"(c)CRI"
                    if loop:
                        assert (f.size % 0x800) == 0                          # looping files end on a sector boundary (:34)
    info, frames = oracle.hca_encode(_sine_channels(1, 5000), 44100)[:2]
    f = oracle.hca_write(info, frames)
    assert bytes(f[:4]) == b"HCA\0" and bytes(f[8:12]) == b"fmt\0" and bytes(f[24:28]) == b"comp"
    assert oracle.crc16(bytes(f[:info.header_size])) == 0
    assert f.size == info.header_size + info.frame_size * info.frame_count
    dec, enc = oracle.hca_key_tables(56, 99)
    fk = oracle.hca_write(info, frames, enc, 56, comment="hi", volume=0.5)
    assert bytes(fk[:4]) == bytes([0xC8, 0xC3, 0xC1, 0]) and oracle.crc16(bytes(fk[:info.header_size])) == 0
    assert b"\xe3\xef\xed\xed\x00hi\x00" in bytes(fk[:info.header_size])     # masked "comm\0" + UTF8Z comment


def test_parsers_agree_on_mutated_and_truncated_files(oracle, vg):
    """Differential fuzz of the two host parsers (product C++ vs oracle C) on WAVE and DSP images with random header bytes
    and random truncation: same accept / reject decision, same fields when accepted, and no out-of-bounds read (the
    images are exact-size numpy buffers; the oracle build under ASan is the memory check)."""
    from vgaudio_b200 import _native as N

    rng = np.random.default_rng(20260924)
    waves = [oracle.wave_write16(_sine_channels(ch, 300), 32000, loop) for ch, loop in ((1, None), (2, (3, 250)), (3, None), (8, (0, 300)))]
    waves.append(_wave8([np.arange(100) % 256, np.arange(100) % 256])[0])
    n_ok = 0
    for case in range(3000):
        img = waves[case % len(waves)].copy()
        head = min(img.size, 140)
        for _ in range(int(rng.integers(1, 4))):
            img[int(rng.integers(0, head))] = int(rng.integers(0, 256))
        if rng.random() < 0.3:
            img = img[: int(rng.integers(0, img.size + 1))].copy()
        st, info = oracle.wave_parse(img)
        pinfo = N.VgbWaveInfo()
        pst = vg.lib.vgb_wave_parse(img.ctypes.data if img.size else None, img.size, C.byref(pinfo))
        assert (st == 0) == (pst == 0), (case, st, pst, vg.lib.vgb_last_error())
        if st == 0:
            n_ok += 1
            for f, _ in N.VgbWaveInfo._fields_:
                assert getattr(pinfo, f) == getattr(info, f), (case, f)
    assert 300 < n_ok < 2900   # the mutations produce both outcomes

    pcm = _sine_channels(2, 3000)
    coefs = np.stack([oracle.calculate_coefficients(p) for p in pcm])
    adpcm = [oracle.encode(p, c) for p, c in zip(pcm, coefs)]
    dsps = [oracle.dsp_write(adpcm, coefs, 32000, 3000), oracle.dsp_write(adpcm[:1], coefs[:1], 32000, 3000, (14, 2800), np.zeros((1, 3), np.int16))]
    n_ok = 0
    for case in range(2000):
        img = dsps[case % 2].copy()
        for _ in range(int(rng.integers(1, 3))):
            img[int(rng.integers(0, 0xC0))] = int(rng.integers(0, 256))
        if rng.random() < 0.3:
            img = img[: int(rng.integers(0, img.size + 1))].copy()
        st, info = oracle.dsp_parse(img)
        pinfo = N.VgbDspInfo()
        pst = vg.lib.vgb_dsp_parse(img.ctypes.data if img.size else None, img.size, C.byref(pinfo)) if img.size else -2
        if img.size == 0:
            assert st != 0
            continue
        assert (st == 0) == (pst == 0), (case, st, pst, vg.lib.vgb_last_error())
        if st == 0:
            n_ok += 1
            for name in ("sample_count", "nibble_count", "sample_rate", "looping", "channel_count", "frames_per_interleave", "loop_start", "loop_end"):
                assert getattr(pinfo, name) == getattr(info, name), (case, name)
    assert n_ok > 100


def test_converter_sizing_pass_is_host_only_and_matches_oracle_sizes(oracle, vg):
    """vgb_convert_wave_batch / vgb_convert_dsp_to_wave_batch with files_out == NULL: parse, validate and size every output on
    the host (no device needed); sizes equal the oracle writers' file sizes, a bad file gets a status and size 0."""
    from vgaudio_b200 import _native as N
    from vgaudio_b200 import containers as ct

    specs = [(1, 48000, None, 48000), (2, 30001, (5, 30000), 44100), (3, 777, None, 22050), (1, 14 * 5000 + 3, (1000, 60000), 32000), (6, 100, None, 8000)]
    files, meta = [], []
    for ch, n, loop, rate in specs:
        pcm = _sine_channels(ch, n, rate)
        files.append(oracle.wave_write16(pcm, rate, loop))
        meta.append((pcm, n, loop, rate))
    files.append(np.frombuffer(b"RIFF\x10\x00\x00\x00WAVEjunkjunkjunk", dtype=np.uint8))
    n = len(files)
    ftab = (C.c_void_p * n)(*[f.ctypes.data for f in files])
    lens = (C.c_int64 * n)(*[f.size for f in files])
    for out_type in (ct.CONTAINER_DSP, ct.CONTAINER_ADX, ct.CONTAINER_HCA):
        sizes, status = (C.c_int64 * n)(), (C.c_int32 * n)()
        opt = ct.convert_options(out_type, hca_quality=2)
        assert vg.lib.vgb_convert_wave_batch(ftab, lens, n, C.byref(opt), sizes, None, status, None, None) == 0
        assert status[n - 1] != 0 and sizes[n - 1] == 0
        for k, (pcm, ns, loop, rate) in enumerate(meta):
            ch = len(pcm)
            assert status[k] == 0, (out_type, k, vg.lib.vgb_last_error())
            if out_type == ct.CONTAINER_DSP:
                zeros = [np.zeros(oracle.sample_count_to_byte_count(ns), np.uint8) for _ in range(ch)]
                want = oracle.dsp_write(zeros, np.zeros((ch, 16), np.int16), rate, ns, loop, np.zeros((ch, 3), np.int16) if loop else None).size
            elif out_type == ct.CONTAINER_ADX:
                align = (-loop[0]) % (64 if ch == 1 else 32) if loop else 0
                audio = [np.zeros(oracle.lib().vgo_adx_encoded_byte_count(ns, align, 18), np.uint8) for _ in range(ch)]
                want = oracle.adx_write(audio, [0] * ch, rate, ns, loop, align).size
            else:
                info = oracle.hca_init(oracle.hca_params(pcm, rate, 2, 0, False, loop))
                want = info.header_size + info.frame_size * info.frame_count
            assert sizes[k] == want, (out_type, k)
    # the decode direction: .dsp images -> WAVE sizes
    dsps = []
    for pcm, ns, loop, rate in meta[:3]:
        ch = len(pcm)
        zeros = [np.zeros(oracle.sample_count_to_byte_count(ns), np.uint8) for _ in range(ch)]
        dsps.append(oracle.dsp_write(zeros, np.zeros((ch, 16), np.int16), rate, ns, loop, np.zeros((ch, 3), np.int16) if loop else None, trim_file=False))
    m = len(dsps)
    dtab = (C.c_void_p * m)(*[f.ctypes.data for f in dsps])
    dl = (C.c_int64 * m)(*[f.size for f in dsps])
    sizes, status = (C.c_int64 * m)(), (C.c_int32 * m)()
    assert vg.lib.vgb_convert_dsp_to_wave_batch(dtab, dl, m, sizes, None, status) == 0
    for k, (pcm, ns, loop, rate) in enumerate(meta[:3]):
        assert status[k] == 0 and sizes[k] == oracle.wave_write16(pcm, rate, loop).size, k
