"""Host-side mirror of the reference's format layer for the GC-ADPCM path (the drop-in boundary).

  Pcm16Format      Formats/Pcm16/Pcm16Format.cs:12-50        short[][] Channels + SampleRate
  GcAdpcmChannel   Formats/GcAdpcm/GcAdpcmChannel.cs:6-29    Adpcm, Coefs, SampleCount
  GcAdpcmFormat    Formats/GcAdpcm/GcAdpcmFormat.cs:14-74    EncodeFromPcm16 (:58-74), ToPcm16 (:42-54)

The reference runs `Parallel.For(0, ChannelCount, i => EncodeChannel(...))`; here the whole loop is ONE batched
call into libvgaudio_b200.so.  The post-encode channel rebuild of GcAdpcmChannelBuilder is
`align_loops` below (GcAdpcmAlignment) and `gcadpcm.seek_table_and_loop_context`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import gcadpcm
from .gcadpcm import GcAdpcmParameters


@dataclass
class Pcm16Format:
    channels: List[np.ndarray]
    sample_rate: int = 48000

    def __post_init__(self):
        if isinstance(self.channels, np.ndarray) and self.channels.ndim == 2:
            base = np.ascontiguousarray(self.channels, dtype=np.int16)
            self.channels = [base[i] for i in range(base.shape[0])]
        else:
            self.channels = [np.ascontiguousarray(c, dtype=np.int16) for c in self.channels]
        lengths = {len(c) for c in self.channels}
        if len(lengths) > 1:  # Pcm16FormatBuilder.cs:14-28 throws InvalidDataException
            raise ValueError("All channels must have the same sample count")

    @property
    def channel_count(self) -> int:
        return len(self.channels)

    @property
    def sample_count(self) -> int:
        return len(self.channels[0]) if self.channels else 0


@dataclass
class GcAdpcmChannel:
    adpcm: np.ndarray
    coefs: np.ndarray
    sample_count: int
    history1: int = 0  # StartContext.Hist1
    history2: int = 0

    def __post_init__(self):
        # GcAdpcmChannel.cs:33-36
        if len(self.adpcm) < gcadpcm.sample_count_to_byte_count(self.sample_count):
            raise ValueError("Audio array length is too short for the specified number of samples.")


@dataclass
class GcAdpcmFormat:
    channels: List[GcAdpcmChannel] = field(default_factory=list)
    sample_rate: int = 48000

    @property
    def channel_count(self) -> int:
        return len(self.channels)

    @property
    def sample_count(self) -> int:
        return self.channels[0].sample_count if self.channels else 0

    def encode_from_pcm16(self, pcm16: Pcm16Format, config: Optional[GcAdpcmParameters] = None) -> "GcAdpcmFormat":
        """GcAdpcmFormat.EncodeFromPcm16(Pcm16Format, GcAdpcmParameters) (GcAdpcmFormat.cs:58-74)."""
        config = config or GcAdpcmParameters()
        if config.progress:
            pass  # SetTotal(frameCount * channels) is implied: the deltas reported sum to it (GcAdpcmFormat.cs:62-63)
        coefs, adpcm = gcadpcm.encode_batch(pcm16.channels, configs=[config] * pcm16.channel_count,
                                            progress=config.progress)
        chans = [GcAdpcmChannel(adpcm[i], coefs[i].copy(), pcm16.sample_count) for i in range(pcm16.channel_count)]
        return GcAdpcmFormat(chans, pcm16.sample_rate)

    def to_pcm16(self) -> Pcm16Format:
        """GcAdpcmFormat.ToPcm16() (GcAdpcmFormat.cs:42-54 -> GcAdpcmChannel.GetPcmAudio :57-60)."""
        if not self.channels:
            return Pcm16Format([], self.sample_rate)
        cfg = [GcAdpcmParameters(c.sample_count, c.history1, c.history2) for c in self.channels]
        pcm = gcadpcm.decode_batch([c.adpcm for c in self.channels], np.stack([c.coefs for c in self.channels]), cfg)
        return Pcm16Format(pcm, self.sample_rate)


# ---- GcAdpcmAlignment (Formats/GcAdpcm/GcAdpcmAlignment.cs:20-63) -------------------------------------------------
# The loop-alignment re-encode is a composition of primitives the library already has - decode, encode with given
# coefficients / history / sample count, decode - plus array copies; batched over channels it is three native calls.

@dataclass
class GcAdpcmAlignment:
    alignment_needed: bool
    loop_start: int
    loop_end: int
    loop_start_aligned: int = 0
    sample_count_aligned: int = 0
    adpcm_aligned: Optional[np.ndarray] = None
    pcm_aligned: Optional[np.ndarray] = None


def align_loops(adpcm_channels, coefs, multiple: int, loop_start: int, loop_end: int) -> List[GcAdpcmAlignment]:
    """GcAdpcmAlignment(multiple, loopStart, loopEnd, adpcm, coefs) for every channel of a format (the channels of one
    stream share the loop points).  Mirrors the reference step by step; the arithmetic happens in the three batched
    native calls."""
    n = len(adpcm_channels)
    co = np.ascontiguousarray(coefs, dtype=np.int16).reshape(n, 16)
    needed = not (multiple == 0 or loop_start % multiple == 0)  # Helpers.LoopPointsAreAligned
    if not needed or n == 0:
        return [GcAdpcmAlignment(False, loop_start, loop_end) for _ in range(n)]
    loop_length = loop_end - loop_start
    loop_start_aligned = loop_start + (multiple - loop_start % multiple)  # GetNextMultiple
    sample_count_aligned = loop_end + (loop_start_aligned - loop_start)
    frames_to_keep = loop_end // gcadpcm.SAMPLES_PER_FRAME
    bytes_to_keep = frames_to_keep * gcadpcm.BYTES_PER_FRAME
    samples_to_keep = frames_to_keep * gcadpcm.SAMPLES_PER_FRAME
    samples_to_encode = sample_count_aligned - samples_to_keep

    old_pcm = gcadpcm.decode_batch(adpcm_channels, co, [GcAdpcmParameters(sample_count=loop_end)] * n)   # :42
    new_pcm, configs, pcm_aligned = [], [], []
    for c in range(n):
        aligned = np.zeros(sample_count_aligned, dtype=np.int16)
        aligned[:loop_end] = old_pcm[c][:loop_end]
        tail = np.zeros(samples_to_encode, dtype=np.int16)
        tail[: loop_end - samples_to_keep] = old_pcm[c][samples_to_keep:loop_end]
        cur = loop_end - samples_to_keep
        while cur < samples_to_encode:                                                                     # :48-51
            k = min(loop_length, samples_to_encode - cur)
            tail[cur:cur + k] = aligned[loop_start:loop_start + k]
            cur += loop_length
        new_pcm.append(tail)
        pcm_aligned.append(aligned)
        configs.append(GcAdpcmParameters(sample_count=samples_to_encode,
                                         history1=int(old_pcm[c][samples_to_keep - 1]) if samples_to_keep >= 1 else 0,
                                         history2=int(old_pcm[c][samples_to_keep - 2]) if samples_to_keep >= 2 else 0))
    _, new_adpcm = gcadpcm.encode_batch(new_pcm, coefs=co, configs=configs)                               # :57
    decoded = gcadpcm.decode_batch(new_adpcm, co, configs)                                                 # :61
    out = []
    for c in range(n):
        adpcm_aligned = np.zeros(gcadpcm.sample_count_to_byte_count(sample_count_aligned), dtype=np.uint8)
        adpcm_aligned[:bytes_to_keep] = np.asarray(adpcm_channels[c], dtype=np.uint8)[:bytes_to_keep]
        adpcm_aligned[bytes_to_keep:bytes_to_keep + len(new_adpcm[c])] = new_adpcm[c]
        pcm_aligned[c][samples_to_keep:samples_to_keep + samples_to_encode] = decoded[c][:samples_to_encode]
        out.append(GcAdpcmAlignment(True, loop_start, loop_end, loop_start_aligned, sample_count_aligned,
                                    adpcm_aligned, pcm_aligned[c]))
    return out
