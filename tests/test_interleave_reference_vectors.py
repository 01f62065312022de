"""The golden vectors the reference holds for InterleaveExtensions (src/VGAudio.Tests/Utilities/InterleaveTests.cs:12-100,
DeinterleaveTests.cs:12-110), replayed literally: on the oracle (CPU) and through the C ABI on the GPU (`vgb_interleave` /
`vgb_deinterleave`).  These pin the block semantics the container writers rely on - shorter last block on either side,
output longer / shorter than the input, zero fill."""
import numpy as np
import pytest


def seq(n):
    return np.arange(n, dtype=np.uint8)


def rows(*lists):
    return [np.array(r, dtype=np.uint8) for r in lists]


D16S8C2 = rows([0, 1, 2, 3, 4, 5, 6, 7], [8, 9, 10, 11, 12, 13, 14, 15])
D16S4C2 = rows([0, 1, 2, 3, 8, 9, 10, 11], [4, 5, 6, 7, 12, 13, 14, 15])
D16S4C4 = rows([0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15])
D16S4OUT4 = rows([0, 1, 2, 3, 0, 0], [4, 5, 6, 7, 0, 0], [8, 9, 10, 11, 0, 0], [12, 13, 14, 15, 0, 0])
D20S4OUT5 = rows([0, 1, 2, 3, 16, 0, 0, 0, 0, 0], [4, 5, 6, 7, 17, 0, 0, 0, 0, 0], [8, 9, 10, 11, 18, 0, 0, 0, 0, 0], [12, 13, 14, 15, 19, 0, 0, 0, 0, 0])
D16S8L6 = rows([0, 1, 2, 3, 4, 5], [8, 9, 10, 11, 12, 13])
I16S8L6 = np.array([0, 1, 2, 3, 4, 5, 0, 0, 8, 9, 10, 11, 12, 13, 0, 0], dtype=np.uint8)
D26S4L6 = rows([0, 1, 2, 3, 8, 9], [4, 5, 6, 7, 12, 13])
I26S4L6 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 0, 12, 13] + [0] * 12, dtype=np.uint8)
D15S2C5L3 = rows([0, 1, 10], [2, 3, 11], [4, 5, 12], [6, 7, 13], [8, 9, 14])
D15S2C5L2 = rows([0, 1], [2, 3], [4, 5], [6, 7], [8, 9])
D12S2C4L6 = rows([0, 1, 8, 0, 0, 0], [2, 3, 9, 0, 0, 0], [4, 5, 10, 0, 0, 0], [6, 7, 11, 0, 0, 0])
D10S2C5L1 = rows([0], [2], [4], [6], [8])

# InterleaveTests.ArrayData (:84-98): (inputs, expected output, interleaveSize, outputSize)
INTERLEAVE = [
    (D16S8C2, seq(16), 8, -1), (D16S4C2, seq(16), 4, -1), (D16S4C4, seq(16), 4, -1), (D16S4OUT4, seq(16), 4, 4), (D20S4OUT5, seq(20), 4, 5),
    (D16S8L6, I16S8L6, 8, 8), (D26S4L6, I26S4L6, 4, 13), (D15S2C5L3, seq(15), 2, -1), (D15S2C5L3, seq(10), 2, 2),
]
# DeinterleaveTests.ArrayData (:83-100): (input, expected outputs, interleaveSize, numInputs, outputSize)
DEINTERLEAVE = [
    (seq(16), D16S8C2, 8, 2, -1), (seq(16), D16S8L6, 8, 2, 6), (seq(26), D26S4L6, 4, 2, 6), (seq(16), D16S4C2, 4, 2, -1), (seq(16), D16S4C4, 4, 4, -1),
    (seq(16), D16S4C4, 8, 4, -1), (seq(16), D16S4OUT4, 4, 4, 6), (seq(12), D12S2C4L6, 2, 4, 6), (seq(15), D15S2C5L2, 2, 5, 2), (seq(15), D15S2C5L3, 2, 5, 3),
    (seq(10), D10S2C5L1, 2, 5, 1), (seq(16), [seq(16)], 8, 1, -1),
]


@pytest.mark.parametrize("case", range(len(INTERLEAVE)))
def test_oracle_interleave_reference_vectors(oracle, case):
    inputs, want, size, out_size = INTERLEAVE[case]
    assert oracle.interleave(inputs, size, out_size).tolist() == want.tolist()


@pytest.mark.parametrize("case", range(len(DEINTERLEAVE)))
def test_oracle_deinterleave_reference_vectors(oracle, case):
    data, want, size, count, out_size = DEINTERLEAVE[case]
    got = oracle.deinterleave(data, size, count, out_size)
    assert [g.tolist() for g in got] == [w.tolist() for w in want]


def test_oracle_uneven_length_is_rejected(oracle):   # DeinterleaveTests.FailsIfInterleavedLengthUnevenArray (:141-153)
    for n in (11, 10):
        with pytest.raises(Exception):
            oracle.deinterleave(seq(n), 1, 3)
    assert len(oracle.deinterleave(seq(9), 1, 3)) == 3


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(INTERLEAVE)))
def test_gpu_interleave_reference_vectors(vg, case):
    inputs, want, size, out_size = INTERLEAVE[case]
    assert vg.interleave.interleave(inputs, size, out_size).tolist() == want.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(DEINTERLEAVE)))
def test_gpu_deinterleave_reference_vectors(vg, case):
    data, want, size, count, out_size = DEINTERLEAVE[case]
    got = vg.interleave.deinterleave(data, size, count, out_size)
    assert [g.tolist() for g in got] == [w.tolist() for w in want]
