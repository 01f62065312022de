// common.cuh — shared device helpers for the sm_100a codec kernels.
//
// Numeric contract: the reference is C# compiled by RyuJIT for x64, i.e. IEEE-754
// binary64/binary32 with NO fused multiply-add, wrapping (unchecked) int32 arithmetic, truncating integer
// division and arithmetic right shifts.  The whole library is compiled with --fmad=false -prec-div=true
// -prec-sqrt=true; where a fused operation is wanted for speed it is written explicitly (__fmaf_rn) together
// with the argument for why its single rounding cannot differ.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vgb {

constexpr int kGcFrameSamples = 14;
constexpr int kGcFrameBytes = 8;
constexpr int kGcFrameNibbles = 16;

// C# unchecked int32 arithmetic: do it in uint32 so C++ has no UB and wraps identically (A.7).
__host__ __device__ __forceinline__ int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__host__ __device__ __forceinline__ int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
__host__ __device__ __forceinline__ int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

// Helpers.Clamp16 / Clamp4 (Utilities/Helpers.cs:32-48)
__host__ __device__ __forceinline__ int32_t clamp16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
__host__ __device__ __forceinline__ int32_t clamp4(int32_t v) { return v > 7 ? 7 : (v < -8 ? -8 : v); }

// GcAdpcmMath.cs:20-27,46 / Extensions.cs:145-146
__host__ __device__ __forceinline__ int32_t gc_sample_count_to_nibble_count(int32_t n)
{
    int32_t whole = n / kGcFrameSamples, rest = n % kGcFrameSamples;
    return kGcFrameNibbles * whole + (rest == 0 ? 0 : rest + 2);
}
__host__ __device__ __forceinline__ int32_t gc_sample_count_to_byte_count(int32_t n)
{
    int32_t nib = gc_sample_count_to_nibble_count(n);
    return nib / 2 + (nib & 1);
}
__host__ __device__ __forceinline__ int32_t gc_nibble_count_to_sample_count(int32_t nib)
{
    int32_t whole = nib / kGcFrameNibbles, rest = nib % kGcFrameNibbles;
    return kGcFrameSamples * whole + (rest < 2 ? 0 : rest - 2);
}
// DivideByRoundUp for non-negative operands (the reference goes through a double; identical for n < 2^31).
__host__ __device__ __forceinline__ int32_t div_round_up(int32_t a, int32_t b) { return (int32_t)(((int64_t)a + b - 1) / b); }

// Per-batch channel tables, all resident in HBM (uploaded once per call).
struct GcChannelTable {
    const int64_t *pcm_off;    // [ch] sample offset of the channel in the PCM slab (multiple of 8)
    const int64_t *adpcm_off;  // [ch] byte offset of the channel in the ADPCM slab (multiple of 16)
    const int64_t *rec_off;    // [ch] frame offset of the channel in the record slab (multiple of 32)
    const int32_t *n_samples;  // [ch] PCM length (coefficient analysis length / decode sample count)
    const int32_t *enc_count;  // [ch] samples to encode (<= n_samples)
    int16_t *hist;             // [ch][2] running history: [0] = hist1 (newest), [1] = hist2
    int32_t n_channels;
};

// One channel of a CRI ADX batch (mirror of CriAdxParameters, Codecs/CriAdx/CriAdxParameters.cs:3-13, plus layout).
struct AdxChannel {
    int64_t pcm_off;     // sample offset in the PCM slab
    int64_t adpcm_off;   // byte offset in the ADPCM slab
    int32_t n_samples;   // encode: pcm.Length; decode: sampleCount
    int32_t frame_size, version, padding, type, filter;
    int32_t history;     // decode only (CriAdxParameters.History)
    int16_t coef0, coef1;  // fixed-table pair or CalculateCoefficients (host, once per distinct sample rate)
};

}  // namespace vgb
