"""Host-side mirror of VGAudio.Utilities.InterleaveExtensions (Utilities/Interleave.cs:9-166) for byte payloads over the
C ABI: the block (de)interleave the container writers / readers run next to the codec path."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import _native as N


def interleave(inputs: Sequence[np.ndarray], interleave_size: int, output_size: int = -1) -> np.ndarray:
    """T[] Interleave<T>(this T[][] inputs, int interleaveSize, int outputSize = -1) for bytes."""
    arrs = [np.ascontiguousarray(a, dtype=np.uint8).ravel() for a in inputs]
    if len({a.size for a in arrs}) > 1:
        raise ValueError("Inputs must be of equal length")  # ArgumentOutOfRangeException (:15-16)
    count, in_size = len(arrs), arrs[0].size
    out_size = in_size if output_size == -1 else output_size
    out = np.zeros(out_size * count, dtype=np.uint8)
    tab = (C.c_void_p * count)(*[a.ctypes.data for a in arrs])
    N.check(N.lib.vgb_interleave(tab, count, in_size, interleave_size, out_size, out.ctypes.data))
    return out


def deinterleave(data: np.ndarray, interleave_size: int, output_count: int, output_size: int = -1) -> List[np.ndarray]:
    """T[][] DeInterleave<T>(this T[] input, int interleaveSize, int outputCount, int outputSize = -1) for bytes."""
    data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    in_size = data.size // max(output_count, 1)
    out_size = in_size if output_size == -1 else output_size
    outs = [np.zeros(out_size, dtype=np.uint8) for _ in range(output_count)]
    tab = (C.c_void_p * max(output_count, 1))(*[o.ctypes.data for o in outs])
    N.check(N.lib.vgb_deinterleave(data.ctypes.data, data.size, interleave_size, output_count, out_size, tab))
    return outs


def short_to_interleaved_byte(channels: Sequence[np.ndarray]) -> np.ndarray:
    """byte[] ShortToInterleavedByte(this short[][] input) (Interleave.cs:170-187): the WAV writer's sample interleave -
    little-endian 16-bit samples, i.e. a 2-byte block interleave of the channels' bytes."""
    arrs = [np.ascontiguousarray(c, dtype="<i2") for c in channels]
    return interleave([a.view(np.uint8) for a in arrs], 2)


def interleaved_byte_to_short(data: np.ndarray, output_count: int) -> List[np.ndarray]:
    """short[][] InterleavedByteToShort(this byte[] input, int outputCount) (Interleave.cs:189-208): the WAV reader's
    front end (WaveReader.cs:47-51).  Trailing bytes that do not fill a sample of every channel are ignored, as in
    the reference (itemCount = input.Length / 2 / outputCount)."""
    data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    items = data.size // 2 // output_count
    outs = deinterleave(data[: items * 2 * output_count], 2, output_count)
    return [o.view("<i2").astype(np.int16) for o in outs]
