"""Several devices behind the C ABI (vgb_init_devices): every host-pointer batch call is sharded over the bound devices
by longest-first bin packing, one worker thread and pipeline per device.  A single-GPU box binds device 0 three times -
the same code path (per-device contexts, worker threads, result scatter, error re-addressing); on a multi-GPU box the
distinct devices are used.  Results must equal the oracle exactly as in the single-device tests."""
import ctypes as C

import numpy as np
import pytest

from vgaudio_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture
def three_devices(vg):
    import torch

    from vgaudio_b200 import _native as N

    n = torch.cuda.device_count()
    devs = [0, 1 % n, 2 % n] if n > 1 else [0, 0, 0]
    N.check(vg.lib.vgb_shutdown())
    arr = (C.c_int32 * 3)(*devs)
    N.check(vg.lib.vgb_init_devices(arr, 3, 0))
    assert vg.lib.vgb_device_count() == 3
    yield devs
    N.check(vg.lib.vgb_shutdown())
    N.check(vg.lib.vgb_init(0, 0))
    assert vg.lib.vgb_device_count() == 1


def test_gcadpcm_sharded_over_devices(vg, oracle, three_devices):
    lens = [14 * 900 + 3, 5000, 14 * 5000, 77, 14 * 2500 + 1, 30000, 1, 14 * 1200, 48000, 0, 9999, 14 * 3100]
    chans = [synth.channel(50 + i, max(L, 1))[:L] for i, L in enumerate(lens)]
    seen = []
    coefs, adpcm = vg.gcadpcm.encode_batch(chans, progress=seen.append)
    assert sum(seen) == sum((L + 13) // 14 for L in lens)
    for c, pcm in enumerate(chans):
        co = oracle.calculate_coefficients(pcm)
        assert np.array_equal(coefs[c], co), c
        assert adpcm[c].tobytes() == oracle.encode(pcm, co).tobytes(), c
    dec = vg.gcadpcm.decode_batch(adpcm, coefs, [vg.gcadpcm.GcAdpcmParameters(L) for L in lens])
    for c, L in enumerate(lens):
        assert np.array_equal(dec[c], oracle.decode(adpcm[c], coefs[c], L)), c
    # given coefficients + history travel with their channel
    rng = np.random.default_rng(3)
    given = rng.integers(-3000, 3000, (len(chans), 16)).astype(np.int16)
    cfgs = [vg.gcadpcm.GcAdpcmParameters(-1, int(rng.integers(-500, 500)), int(rng.integers(-500, 500))) for _ in chans]
    _, adpcm2 = vg.gcadpcm.encode_batch(chans, given, cfgs)
    for c, pcm in enumerate(chans):
        assert adpcm2[c].tobytes() == oracle.encode(pcm, given[c], -1, cfgs[c].history1, cfgs[c].history2).tobytes(), c
    # an error names the caller's channel index, not the shard's
    bad = [a.copy() for a in adpcm]
    bad[8][0] |= 0x80
    with pytest.raises(vg.VgbError) as e:
        vg.gcadpcm.decode_batch(bad, coefs, [vg.gcadpcm.GcAdpcmParameters(L) for L in lens])
    assert e.value.code == -2 and "channel 8:" in str(e.value)


def test_criadx_and_crihca_sharded_over_devices(vg, oracle, three_devices):
    lens = [3200, 32 * 700 + 5, 48000, 100, 32 * 1500, 7777, 20000]
    chans = [synth.channel(70 + i, L) for i, L in enumerate(lens)]
    P = vg.criadx.CriAdxParameters
    cfgs = [P(sample_rate=48000, frame_size=18, version=4 - (i % 2), type=3 + (i % 2)) for i in range(len(chans))]
    adpcm, hist = vg.criadx.encode_batch(chans, cfgs)
    for c, pcm in enumerate(chans):
        want, h = oracle.adx_encode(pcm, 48000, 18, cfgs[c].version, 0, cfgs[c].type, 0)
        assert adpcm[c].tobytes() == want.tobytes() and int(hist[c]) == h, c
    streams = [[synth.channel(90 + 2 * s, 9000 + 700 * s), synth.channel(91 + 2 * s, 9000 + 700 * s)] for s in range(5)]
    infos, frames = vg.crihca.encode_batch(streams, 48000)
    for s, st in enumerate(streams):
        o_info, o_frames = oracle.hca_encode(st, 48000, 2)
        assert np.array_equal(frames[s], o_frames), s
        assert infos[s].frame_count == o_info.frame_count
    pcm = vg.crihca.decode_batch(infos, frames)
    for s in range(5):
        assert np.array_equal(np.stack(pcm[s]), oracle.hca_decode(infos[s], frames[s])), s
