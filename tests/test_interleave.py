"""Block (de)interleave (Utilities/Interleave.cs:9-166): oracle self-checks on CPU, CUDA vs oracle on
the GPU.  The reference's own expectations are restated from src/VGAudio.Tests/Utilities/InterleaveTests-style cases:
round trip, shorter last block, output size smaller / larger than the input."""
import numpy as np
import pytest

SHAPES = [  # (count, in_size, interleave, out_size)
    (2, 64, 16, -1), (2, 100, 16, -1), (3, 100, 16, 112), (2, 100, 16, 90), (1, 77, 8, -1), (4, 0x2000 * 3 + 0x150, 0x2000, -1),
    (2, 36, 18, -1), (6, 18 * 50, 18, 18 * 50), (2, 1000, 2, -1), (5, 33, 7, 40), (2, 128, 256, -1), (3, 5, 1, 9), (8, 4096, 512, 4096 + 512),
]


def _inputs(count, in_size, seed=1):
    rng = np.random.default_rng(seed + count * 1000 + in_size)
    return [rng.integers(0, 256, in_size, dtype=np.uint8) for _ in range(count)]


@pytest.mark.parametrize("count,in_size,interleave,out_size", SHAPES)
def test_oracle_interleave_layout_and_round_trip(oracle, count, in_size, interleave, out_size):
    ins = _inputs(count, in_size)
    out = oracle.interleave(ins, interleave, out_size)
    osz = in_size if out_size == -1 else out_size
    assert out.size == osz * count
    # independent statement of the layout: walk the output blocks
    want = np.zeros(osz * count, np.uint8)
    in_blocks, out_blocks = -(-in_size // interleave), -(-osz // interleave)
    for b in range(min(in_blocks, out_blocks)):
        cur_in = in_size - b * interleave if b == in_blocks - 1 else interleave
        cur_out = osz - b * interleave if b == out_blocks - 1 else interleave
        n = min(cur_in, cur_out)
        for i in range(count):
            want[interleave * b * count + cur_out * i: interleave * b * count + cur_out * i + n] = ins[i][interleave * b: interleave * b + n]
    assert np.array_equal(out, want)
    if out_size in (-1, in_size):  # DeInterleave undoes Interleave
        back = oracle.deinterleave(out, interleave, count)
        for i in range(count):
            assert np.array_equal(back[i], ins[i])


def test_oracle_deinterleave_rejects_indivisible_length(oracle):
    with pytest.raises(ValueError):
        oracle.deinterleave(np.zeros(7, np.uint8), 2, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("count,in_size,interleave,out_size", SHAPES)
def test_gpu_interleave_and_deinterleave_match_oracle(vg, oracle, count, in_size, interleave, out_size):
    ins = _inputs(count, in_size, seed=7)
    got = vg.interleave.interleave(ins, interleave, out_size)
    assert np.array_equal(got, oracle.interleave(ins, interleave, out_size))
    osz = in_size if out_size == -1 else out_size
    data = np.random.default_rng(3).integers(0, 256, in_size * count, dtype=np.uint8)
    back = vg.interleave.deinterleave(data, interleave, count, out_size)
    want = oracle.deinterleave(data, interleave, count, out_size)
    for i in range(count):
        assert back[i].size == osz and np.array_equal(back[i], want[i]), i


@pytest.mark.gpu
def test_gpu_deinterleave_errors(vg):
    with pytest.raises(vg.VgbError):
        vg.interleave.deinterleave(np.zeros(7, np.uint8), 2, 2)
    with pytest.raises(ValueError):
        vg.interleave.interleave([np.zeros(4, np.uint8), np.zeros(5, np.uint8)], 2)


@pytest.mark.gpu
@pytest.mark.parametrize("count,n,extra", [(1, 1000, 0), (2, 48000, 0), (6, 777, 0), (2, 100, 3), (8, 1, 0)])
def test_gpu_wav_sample_interleave_round_trip(vg, count, n, extra):
    """ShortToInterleavedByte / InterleavedByteToShort (Interleave.cs:170-208) as 2-byte block (de)interleaves."""
    rng = np.random.default_rng(5)
    chans = [rng.integers(-32768, 32768, n, dtype=np.int16) for _ in range(count)]
    data = vg.interleave.short_to_interleaved_byte(chans)
    want = np.stack(chans, axis=1).astype("<i2").tobytes()  # sample-major, little-endian
    assert data.tobytes() == want
    padded = np.concatenate([data, np.zeros(extra, np.uint8)])
    back = vg.interleave.interleaved_byte_to_short(padded, count)
    for c in range(count):
        assert np.array_equal(back[c], chans[c])


@pytest.mark.gpu
@pytest.mark.parametrize("count,size,ilv", [(2, 0x2000 * 5 + 3664, 0x2000), (4, 18 * 8 * 100, 18 * 8), (1, 4096, 1024), (3, 48000, 16000), (2, 65536 + 16, 65536)])
def test_tma_variant_matches_oracle(vg, oracle, count, size, ilv):
    """VGB_INTERLEAVE_TMA=1: the same shuffle as bulk copies (cp.async.bulk + mbarrier ring); eligible shapes (everything a
    multiple of 16 bytes, input size == output size) must give the oracle's bytes in both directions."""
    import os

    rng = np.random.default_rng(count * 1000 + ilv)
    chans = [rng.integers(0, 256, size, dtype=np.uint8) for _ in range(count)]
    os.environ["VGB_INTERLEAVE_TMA"] = "1"
    try:
        got = vg.interleave.interleave(chans, ilv)
        assert got.tobytes() == oracle.interleave(chans, ilv).tobytes()
        back = vg.interleave.deinterleave(got, ilv, count)
    finally:
        os.environ.pop("VGB_INTERLEAVE_TMA", None)
    for c in range(count):
        assert np.array_equal(back[c], chans[c]), c
