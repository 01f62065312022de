"""GPU parity tests of the container layer through the C ABI: every file the CUDA path writes is
BYTE-IDENTICAL to the one the CPU oracle (oracle/containers.c over the oracle codecs) writes from the same input, and the
readers return the oracle's arrays."""
import struct

import numpy as np
import pytest

from vgaudio_b200 import synth

pytestmark = pytest.mark.gpu


def _pcm(n_ch, n, first=4):
    return [synth.channel(first + c, max(n, 1))[:n] for c in range(n_ch)]


def _wave8(channels, rate=22050):
    ch = len(channels)
    data = np.stack(channels, axis=1).astype(np.uint8).tobytes()
    fmt = struct.pack("<HHIIHH", 1, ch, rate, rate * ch, ch, 8)
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", len(data)) + data
    return np.frombuffer(b"RIFF" + struct.pack("<I", len(body)) + body, dtype=np.uint8)


# ---- WAVE front end ---------------------------------------------------------------------------------------------------
def test_wave_read_batch_matches_oracle(vg, oracle):
    from vgaudio_b200 import containers as ct

    files = []
    for n_ch, n, loop in [(1, 1, None), (1, 48000, None), (2, 30001, (5, 30000)), (3, 777, None), (8, 12345, None),
                          (2, 8192, None), (5, 8200, None), (64, 300, None)]:
        files.append(oracle.wave_write16(_pcm(n_ch, n), 44100, loop))
    rng = np.random.default_rng(3)
    files.append(_wave8([rng.integers(0, 256, 5001) for _ in range(3)]))
    files.append(_wave8([rng.integers(0, 256, 17)]))
    # an empty data chunk is not seen by RiffParser's loop (`Position + 8 < endOffset`, RiffParser.cs:50): both sides reject it
    empty = oracle.wave_write16(_pcm(6, 0), 44100, None)
    assert oracle.wave_parse(empty)[0] != 0
    with pytest.raises(vg.VgbError):
        ct.wave_parse(empty)
    got = ct.wave_read_batch(files)
    for f, (info, rows) in zip(files, got):
        st, oi = oracle.wave_parse(f)
        assert st == 0 and info.sample_count == oi.sample_count and info.channel_count == oi.channel_count
        want = oracle.wave_read(f, oi)
        assert len(rows) == len(want)
        for a, b in zip(rows, want):
            assert np.array_equal(a, b)


# ---- DSP --------------------------------------------------------------------------------------------------------------
def _encode_gc(oracle, pcm):
    coefs = np.stack([oracle.calculate_coefficients(p) for p in pcm])
    return coefs, [oracle.encode(p, c) for p, c in zip(pcm, coefs)]


def _loop_ctx(oracle, adpcm, coefs, n, loop_start):
    return np.stack([np.array(oracle.gc_loop_context(a, oracle.decode(a, c, n), loop_start), dtype=np.int16) for a, c in zip(adpcm, coefs)])


DSP_CASES = [
    # channels, samples, loop, samples_per_interleave, loop alignment, trim
    (1, 14 * 100, None, 0, 0, True), (1, 14 * 100 + 3, None, 0, 0, True), (1, 1, None, 0, 0, True),
    (2, 14 * 2000 + 9, None, 0, 0, True), (2, 14 * 2000 + 9, None, 14 * 64, 0, True), (3, 50000, None, 14, 0, True),
    (2, 40000, (1000, 30000), 0, 0, True), (2, 40000, (1000, 30000), 0, 0, False), (1, 40000, (1003, 39000), 0, 14, False),
    (4, 14 * 1024, (0, 14 * 1024), 14 * 256, 0, True), (6, 20011, (17, 20011), 0, 0, True), (2, 14 * 0x3800 // 14 * 2 + 5, None, 0, 0, True),
]


def test_dsp_write_batch_matches_oracle(vg, oracle):
    from vgaudio_b200 import containers as ct

    files, want = [], []
    for k, (ch, n, loop, spi, lpa, trim) in enumerate(DSP_CASES):
        pcm = _pcm(ch, n, first=4 + k)
        coefs, adpcm = _encode_gc(oracle, pcm)
        ctx = _loop_ctx(oracle, adpcm, coefs, n, loop[0]) if loop else None
        gain = np.arange(ch, dtype=np.int16) * 3
        hist = np.arange(2 * ch, dtype=np.int16).reshape(ch, 2) - 4
        files.append(ct.DspFile(adpcm, coefs, 32000 + k, n, loop is not None, loop[0] if loop else 0, loop[1] if loop else 0, ctx, gain, hist,
                                spi, lpa, trim))
        want.append(oracle.dsp_write(adpcm, coefs, 32000 + k, n, loop, ctx, gain, hist, spi or 0x3800, lpa or 1, trim))
    got = ct.dsp_write_batch(files)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g.size == w.size, DSP_CASES[k]
        assert g.tobytes() == w.tobytes(), (DSP_CASES[k], int(np.flatnonzero(g != w)[0]))


def test_dsp_write_rejects_bad_descriptions(vg, oracle):
    from vgaudio_b200 import containers as ct

    pcm = _pcm(1, 2000)
    coefs, adpcm = _encode_gc(oracle, pcm)
    with pytest.raises(vg.VgbError):  # DspConfiguration.SamplesPerInterleave must be divisible by 14
        ct.dsp_write_batch([ct.DspFile(adpcm, coefs, 32000, 2000, samples_per_interleave=100)])
    with pytest.raises(vg.VgbError):  # loop end past the audio
        ct.dsp_write_batch([ct.DspFile(adpcm, coefs, 32000, 2000, True, 10, 3000, np.zeros((1, 3), np.int16))])


def test_dsp_read_batch_matches_oracle(vg, oracle):
    from vgaudio_b200 import containers as ct

    files = []
    for k, (ch, n, loop, spi, lpa, trim) in enumerate(DSP_CASES):
        pcm = _pcm(ch, n, first=40 + k)
        coefs, adpcm = _encode_gc(oracle, pcm)
        ctx = _loop_ctx(oracle, adpcm, coefs, n, loop[0]) if loop else None
        files.append(oracle.dsp_write(adpcm, coefs, 32000, n, loop, ctx, None, None, spi or 0x3800, lpa or 1, trim))
    got = ct.dsp_read_batch(files)
    for k, (f, (info, rows)) in enumerate(zip(files, got)):
        st, oi = oracle.dsp_parse(f)
        assert st == 0
        want = oracle.dsp_read_data(f, oi)
        assert len(rows) == len(want)
        for a, b in zip(rows, want):
            assert a.tobytes() == b.tobytes(), DSP_CASES[k]


# ---- CRI ADX ----------------------------------------------------------------------------------------------------------
ADX_CASES = [
    # channels, samples, loop, frame_size, version, type, key, encryption type, trim
    (1, 5000, None, 18, 4, 3, None, 0, True), (2, 5000, None, 18, 3, 3, None, 0, True), (1, 32 * 100, None, 18, 4, 4, None, 0, True),
    (2, 9001, (1000, 8000), 18, 4, 3, None, 0, True), (2, 9001, (1000, 8000), 18, 3, 3, None, 0, False), (1, 9001, (33, 9001), 18, 4, 3, None, 0, True),
    (3, 4000, None, 34, 4, 2, None, 0, True), (2, 6000, None, 18, 4, 3, "karaage", 8, True), (2, 6000, (64, 5000), 18, 4, 3, 0x1234567890, 9, True),
    (6, 777, None, 10, 4, 3, "x", 8, True),
]


def _adx_case(oracle, case, k):
    ch, n, loop, fs, version, typ, key, enc_type, trim = case
    spf = (fs - 2) * 2
    align = 0
    if loop:
        mult = spf * 2 if ch == 1 else spf
        align = (-loop[0]) % mult
    pcm = _pcm(ch, n, first=70 + k)
    enc = [oracle.adx_encode(p, 48000, fs, version, align, typ, 1) for p in pcm]
    okey = None
    if key is not None:
        okey = oracle.adx_key(key_string=key) if isinstance(key, str) else oracle.adx_key(key_code=key)
    want = oracle.adx_write([e[0] for e in enc], [e[1] for e in enc], 48000, n, loop, align, fs, version, typ, 500, enc_type, okey, trim)
    return enc, align, okey, want


def test_adx_write_batch_matches_oracle(vg, oracle):
    from vgaudio_b200 import _native as N
    from vgaudio_b200 import containers as ct

    for keyed in (False, True):  # one key per call: group the cases by it
        groups = {}
        for k, case in enumerate(ADX_CASES):
            if (case[6] is not None) != keyed:
                continue
            groups.setdefault(case[6], []).append((k, case))
        for key, cases in groups.items():
            files, want = [], []
            pk = None
            for k, case in cases:
                ch, n, loop, fs, version, typ, _, enc_type, trim = case
                enc, align, okey, w = _adx_case(oracle, case, k)
                files.append(ct.AdxFile([e[0] for e in enc], [e[1] for e in enc], 48000, n, loop is not None, loop[0] if loop else 0,
                                        loop[1] if loop else 0, align, fs, version, typ, 500, enc_type, trim))
                want.append(w)
                if okey is not None:
                    pk = N.VgbAdxKey(*okey)
            got = ct.adx_write_batch(files, pk)
            for (k, case), g, w in zip(cases, got, want):
                assert g.size == w.size, case
                assert g.tobytes() == w.tobytes(), (case, int(np.flatnonzero(g != w)[0]))


def test_adx_crypt_batch_matches_oracle(vg, oracle):
    from vgaudio_b200 import containers as ct

    pcm = _pcm(3, 20000, first=90)
    audio = [oracle.adx_encode(p)[0] for p in pcm]
    audio[2][18 * 7: 18 * 9] = 0
    for enc_type, key in ((8, ct.adx_key(key_string="karaage")), (9, ct.adx_key(key_code=123456789012))):
        okey = (key.seed, key.mult, key.inc)
        got = ct.adx_crypt(audio, key, enc_type, 18)
        want = oracle.adx_crypt(audio, okey, enc_type, 18)
        for a, b in zip(got, want):
            assert a.tobytes() == b.tobytes()
        if enc_type == 8:
            back = ct.adx_crypt(got, key, 8, 18)
            assert all(a.tobytes() == b.tobytes() for a, b in zip(back, audio))


# ---- CRI HCA ----------------------------------------------------------------------------------------------------------
def test_hca_write_and_crypt_match_oracle(vg, oracle):
    from vgaudio_b200 import _native as N
    from vgaudio_b200 import containers as ct

    streams = [oracle.hca_encode(_pcm(1, 9000, 110), 48000), oracle.hca_encode(_pcm(2, 20000, 112), 44100, quality=3),
               oracle.hca_encode(_pcm(2, 30000, 114), 48000, loop=(2000, 25000)), oracle.hca_encode(_pcm(4, 1500, 116), 32000)]

    def pinfo(oi):
        p = N.VgbHcaInfo()
        for name, _ in N.VgbHcaInfo._fields_:
            setattr(p, name, getattr(oi, name))
        return p

    infos = [pinfo(s[0]) for s in streams]
    frames = [s[1] for s in streams]
    for key_type, key_code in ((-1, 0), (0, 0), (1, 0), (56, 0xCC55463930DBE1AB), (56, 7)):
        table = oracle.hca_key_tables(key_type, key_code)[1] if key_type >= 0 else None
        comments = [None, "a comment", "  ", "x"]
        volumes = [1.0, 0.5, 1.0, 2.0]
        got = ct.hca_write_batch(infos, frames, key_type, key_code, comments, volumes)
        for s, g, cm, vol in zip(streams, got, comments, volumes):
            want = oracle.hca_write(s[0], s[1], table, max(key_type, 0), cm, vol)
            assert g.tobytes() == want.tobytes(), (key_type, cm, int(np.flatnonzero(g != want)[0]))
    # Crypt alone: encrypt == oracle, decrypt(encrypt) == identity
    same = [s for s in streams if s[0].frame_size == streams[0][0].frame_size]
    fs = same[0][0].frame_size
    enc = ct.hca_crypt_batch([s[1] for s in same], fs, 56, 12345)
    dec_t, enc_t = oracle.hca_key_tables(56, 12345)
    for s, e in zip(same, enc):
        assert e.tobytes() == oracle.hca_crypt_frames(s[1], fs, enc_t).tobytes()
    back = ct.hca_crypt_batch(enc, fs, 56, 12345, decrypt=True)
    for s, b in zip(same, back):
        assert b.tobytes() == np.asarray(s[1]).tobytes()


# ---- batch conversion: WAVE in, encoded file out ----------------------------------------------------------------------------
def _batch_inputs(oracle):
    specs = [(1, 48000, None, 48000), (2, 30001, None, 44100), (1, 14 * 5000 + 3, (1000, 60000), 32000), (2, 20000, (2000, 20000), 48000),
             (3, 9000, None, 48000), (1, 1, None, 48000), (1, 100000, None, 48000), (2, 777, None, 22050), (6, 5000, None, 48000),
             (1, 50000, (0, 50000), 48000), (2, 65536, None, 48000), (1, 30000, (1001, 29000), 48000), (2, 9000, (7, 8000), 44100)]
    files, meta = [], []
    for k, (ch, n, loop, rate) in enumerate(specs):
        pcm = _pcm(ch, n, first=130 + 2 * k)
        files.append(oracle.wave_write16(pcm, rate, loop))
        meta.append((pcm, n, loop, rate))
    files.insert(3, np.frombuffer(b"RIFFxxxxJUNKnot a wave file at all", dtype=np.uint8))   # a bad file does not stop the batch
    meta.insert(3, None)
    return files, meta


@pytest.mark.parametrize("group_bytes", [0, 150000])
def test_convert_wave_to_dsp_matches_oracle(vg, oracle, group_bytes):
    from vgaudio_b200 import containers as ct

    files, meta = _batch_inputs(oracle)
    seen = []
    outs, status = ct.convert_wave_batch(files, ct.convert_options(ct.CONTAINER_DSP, group_bytes=group_bytes), progress=seen.append)
    assert sum(seen) == len(files) - 1
    for k, m in enumerate(meta):
        if m is None:
            assert status[k] != 0 and outs[k] is None
            continue
        pcm, n, loop, rate = m
        assert status[k] == 0
        coefs, adpcm = _encode_gc(oracle, pcm)
        ctx = _loop_ctx(oracle, adpcm, coefs, n, loop[0]) if loop else None
        want = oracle.dsp_write(adpcm, coefs, rate, n, loop, ctx)
        assert outs[k].size == want.size, k
        assert outs[k].tobytes() == want.tobytes(), (k, int(np.flatnonzero(outs[k] != want)[0]))


@pytest.mark.parametrize("keyed", [False, True])
def test_convert_wave_to_adx_matches_oracle(vg, oracle, keyed):
    from vgaudio_b200 import containers as ct

    files, meta = _batch_inputs(oracle)
    kw = {}
    okey = None
    if keyed:
        okey = oracle.adx_key(key_string="karaage")
        kw = dict(adx_has_key=1, adx_key_seed=okey[0], adx_key_mult=okey[1], adx_key_inc=okey[2], adx_encryption_type=8)
    outs, status = ct.convert_wave_batch(files, ct.convert_options(ct.CONTAINER_ADX, group_bytes=400000, **kw))
    for k, m in enumerate(meta):
        if m is None:
            assert status[k] != 0
            continue
        pcm, n, loop, rate = m
        ch = len(pcm)
        align = 0
        if loop:
            mult = 64 if ch == 1 else 32
            align = (-loop[0]) % mult
        enc = [oracle.adx_encode(p, rate, 18, 4, align, 3, 0) for p in pcm]
        want = oracle.adx_write([e[0] for e in enc], [e[1] for e in enc], rate, n, loop, align, 18, 4, 3, 500, 8 if keyed else 0, okey)
        assert outs[k].size == want.size, k
        assert outs[k].tobytes() == want.tobytes(), (k, int(np.flatnonzero(outs[k] != want)[0]))


@pytest.mark.parametrize("key_type", [-1, 56])
def test_convert_wave_to_hca_matches_oracle(vg, oracle, key_type):
    from vgaudio_b200 import containers as ct

    files, meta = _batch_inputs(oracle)
    # HCA needs at least a frame's worth of sensible input; keep the ordinary files
    keep = [k for k, m in enumerate(meta) if m is None or m[1] >= 777]
    files = [files[k] for k in keep]
    meta = [meta[k] for k in keep]
    outs, status = ct.convert_wave_batch(files, ct.convert_options(ct.CONTAINER_HCA, hca_quality=2, hca_key_type=key_type, hca_key_code=777,
                                                                   group_bytes=300000))
    table = oracle.hca_key_tables(56, 777)[1] if key_type >= 0 else None
    for k, m in enumerate(meta):
        if m is None:
            assert status[k] != 0
            continue
        pcm, n, loop, rate = m
        assert status[k] == 0, k
        info, frames = oracle.hca_encode(pcm, rate, quality=2, loop=loop)
        want = oracle.hca_write(info, frames, table, max(key_type, 0))
        assert outs[k].size == want.size, k
        assert outs[k].tobytes() == want.tobytes(), (k, int(np.flatnonzero(outs[k] != want)[0]))


def test_convert_8bit_and_empty_batches(vg, oracle):
    from vgaudio_b200 import containers as ct

    rng = np.random.default_rng(9)
    chans = [rng.integers(0, 256, 30000) for _ in range(2)]
    f = _wave8(chans, 32000)
    outs, status = ct.convert_wave_batch([f], ct.convert_options(ct.CONTAINER_DSP))
    pcm = [((c.astype(np.int32) - 0x80) << 8).astype(np.int16) for c in chans]
    coefs, adpcm = _encode_gc(oracle, pcm)
    assert status == [0] and outs[0].tobytes() == oracle.dsp_write(adpcm, coefs, 32000, 30000).tobytes()
    outs, status = ct.convert_wave_batch([], ct.convert_options(ct.CONTAINER_DSP))
    assert outs == [] and status == []


def test_convert_option_variants_match_oracle(vg, oracle):
    """The writer options of the converter: DSP interleave size / no-trim / loop alignment, ADX version 3, frame size,
    Fixed and Exponential types (padded streams take the encoder's general path there), encryption type 9, HCA quality and
    key types 0 / 1 - each against the oracle chain with the same options."""
    from vgaudio_b200 import containers as ct

    files, meta = [], []
    for k, (ch, n, loop, rate) in enumerate([(2, 30000, (1001, 29000), 48000), (1, 20000, (777, 19000), 32000), (4, 12000, None, 44100),
                                             (1, 48000, None, 48000), (2, 5000, (31, 4999), 22050)]):
        pcm = _pcm(ch, n, first=300 + 3 * k)
        files.append(oracle.wave_write16(pcm, rate, loop))
        meta.append((pcm, n, loop, rate))

    # DSP: 14 * 32 samples per interleave, loop points aligned to 14, trimming off
    outs, status = ct.convert_wave_batch(files, ct.convert_options(ct.CONTAINER_DSP, dsp_samples_per_interleave=14 * 32, dsp_loop_point_alignment=14, no_trim=1))
    for k, (pcm, n, loop, rate) in enumerate(meta):
        coefs, adpcm = _encode_gc(oracle, pcm)
        ctx = _loop_ctx(oracle, adpcm, coefs, n, loop[0]) if loop else None
        if loop and loop[1] + (-loop[0]) % 14 > n:   # the aligned loop end lies past the audio: the reference throws for mono
            continue
        want = oracle.dsp_write(adpcm, coefs, rate, n, loop, ctx, None, None, 14 * 32, 14, False)
        assert status[k] == 0 and outs[k].tobytes() == want.tobytes(), ("dsp", k)

    # ADX: version 3, 34-byte frames, each encoding type; key code -> encryption type 9
    okey = oracle.adx_key(key_code=0x123456789A)
    for typ in (2, 3, 4):
        opt = ct.convert_options(ct.CONTAINER_ADX, adx_version=3, adx_frame_size=34, adx_type=typ, adx_filter_plus1=2, adx_has_key=1,
                                 adx_key_seed=okey[0], adx_key_mult=okey[1], adx_key_inc=okey[2], adx_encryption_type=9)
        outs, status = ct.convert_wave_batch(files, opt)
        for k, (pcm, n, loop, rate) in enumerate(meta):
            ch, spf = len(pcm), 64
            align = (-loop[0]) % (spf * 2 if ch == 1 else spf) if loop else 0
            enc = [oracle.adx_encode(p, rate, 34, 3, align, typ, 1) for p in pcm]
            want = oracle.adx_write([e[0] for e in enc], [e[1] for e in enc], rate, n, loop, align, 34, 3, typ, 500, 9, okey)
            assert status[k] == 0 and outs[k].tobytes() == want.tobytes(), ("adx", typ, k, int(np.flatnonzero(outs[k] != want)[0]))

    # HCA: quality Low, key types 0 and 1
    for key_type in (0, 1):
        outs, status = ct.convert_wave_batch(files, ct.convert_options(ct.CONTAINER_HCA, hca_quality=4, hca_key_type=key_type))
        table = oracle.hca_key_tables(key_type)[1]
        for k, (pcm, n, loop, rate) in enumerate(meta):
            info, frames = oracle.hca_encode(pcm, rate, quality=4, loop=loop)
            want = oracle.hca_write(info, frames, table, key_type)
            assert status[k] == 0 and outs[k].tobytes() == want.tobytes(), ("hca", key_type, k)


def test_convert_all_bad_and_many_channels(vg, oracle):
    from vgaudio_b200 import containers as ct

    bad = [np.frombuffer(b"not a riff file", dtype=np.uint8), np.zeros(0, dtype=np.uint8), oracle.wave_write16(_pcm(1, 10), 48000, None)[:20].copy()]
    outs, status = ct.convert_wave_batch(bad, ct.convert_options(ct.CONTAINER_ADX))
    assert outs == [None, None, None] and all(s != 0 for s in status)
    pcm = _pcm(40, 3000, first=400)
    f = oracle.wave_write16(pcm, 16000, (100, 2900))
    outs, status = ct.convert_wave_batch([f], ct.convert_options(ct.CONTAINER_DSP))
    coefs, adpcm = _encode_gc(oracle, pcm)
    want = oracle.dsp_write(adpcm, coefs, 16000, 3000, (100, 2900), _loop_ctx(oracle, adpcm, coefs, 3000, 100))
    assert status == [0] and outs[0].tobytes() == want.tobytes()


def test_convert_dsp_to_wave_matches_oracle(vg, oracle):
    """The decode direction: DspReader -> GcAdpcmDecoder (header coefficients, start history) -> WaveWriter, file bytes
    against the oracle chain; a truncated and a non-DSP image fail alone."""
    from vgaudio_b200 import containers as ct

    files, want = [], []
    for k, (ch, n, loop, spi) in enumerate([(1, 14 * 300, None, 0x3800), (1, 50001, (100, 50001), 0x3800), (2, 30000, (1000, 29000), 0x3800),
                                            (3, 7777, None, 14 * 16), (6, 20011, (17, 20011), 0x3800), (2, 1, None, 0x3800), (8, 9000, None, 14 * 8)]):
        pcm = _pcm(ch, n, first=500 + 4 * k)
        coefs, adpcm = _encode_gc(oracle, pcm)
        ctx = _loop_ctx(oracle, adpcm, coefs, n, loop[0]) if loop else None
        hist = np.array([[7 * c, -3 * c] for c in range(ch)], dtype=np.int16)
        f = oracle.dsp_write(adpcm, coefs, 22050 + k, n, loop, ctx, None, hist, spi, 1, False)
        files.append(f)
        dec = [oracle.decode(a, c, n, int(h[0]), int(h[1])) for a, c, h in zip(adpcm, coefs, hist)]
        want.append(oracle.wave_write16(dec, 22050 + k, loop))
    files.append(files[2][:200].copy())
    files.append(np.frombuffer(b"\x00\x00\x00\x64" + b"\x00\x00\x00\x05" + b"\x00" * 0x58 + b"\x00" * 64, dtype=np.uint8))   # sample count 100, nibble count 5
    outs, status = ct.convert_dsp_to_wave_batch(files)
    for k, w in enumerate(want):
        assert status[k] == 0 and outs[k].size == w.size, k
        assert outs[k].tobytes() == w.tobytes(), (k, int(np.flatnonzero(outs[k] != w)[0]))
    assert status[len(want)] != 0 and status[len(want) + 1] != 0 and outs[len(want)] is None
    # and the round trip through both directions of the batch job
    back, st2 = ct.convert_wave_batch([outs[0], outs[2]], ct.convert_options(ct.CONTAINER_DSP))
    info = ct.dsp_parse(back[1])
    assert st2 == [0, 0] and info.channel_count == 2 and info.sample_count == 29000 and (info.loop_start, info.loop_end) == (1000, 29000)  # trimmed to the loop end (DspWriter.cs:22)
