// gc_decode.cu — GC-ADPCM decoder on sm_100a.
//
// Replaces GcAdpcmDecoder.Decode (Codecs/GcAdpcm/GcAdpcmDecoder.cs:10-54).  The recurrence
//     s[t] = Clamp16((c1*s[t-1] + c2*s[t-2] + scale*nibble + 1024) >> 11)
// is non-linear (the clamp), so a channel is strictly serial; parallelism is across channels: one THREAD owns one
// channel.  To keep HBM traffic coalesced although every thread walks its own stream, a warp moves data through
// shared memory in tiles of 16 frames per channel: 32 x 128 B of ADPCM in (one fully coalesced 128-byte row per
// channel), 32 x 448 B of PCM out (28 coalesced 16-byte vectors per channel).
#include "common.cuh"
#include "kernels.h"

namespace vgb {

constexpr int kDecWarps = 2;               // warps per CTA (64 channels)
constexpr int kDecTileFrames = 16;         // frames per channel per tile: 128 B in, 448 B out
constexpr int kDecInWords = kDecTileFrames * kGcFrameBytes / 4;          // 32 words per channel
constexpr int kDecOutWords = kDecTileFrames * kGcFrameSamples * 2 / 4;   // 112 words per channel
constexpr int kDecInPitch = kDecInWords + 1;     // 33: lane-per-channel reads hit distinct banks
constexpr int kDecOutPitch = kDecOutWords + 1;   // 113 (odd): lane-per-channel writes hit distinct banks

__device__ __forceinline__ int32_t nibble_signed(uint32_t v) { return (int32_t)(v << 28) >> 28; }  // Helpers.cs:50-56

__global__ void __launch_bounds__(kDecWarps * 32)
gc_decode_kernel(const uint8_t *__restrict__ adpcm, GcChannelTable tab, const int16_t *__restrict__ coefs,
                 int16_t *__restrict__ pcm, int frame_begin, int frame_end)
{
    extern __shared__ uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *in_tile = smem + warp * (32 * kDecInPitch + 32 * kDecOutPitch);
    uint32_t *out_tile = in_tile + 32 * kDecInPitch;

    const int ch0 = (blockIdx.x * kDecWarps + warp) * 32;  // first channel of this warp
    if (ch0 >= tab.n_channels) return;
    const int ch = ch0 + lane;
    const bool live = ch < tab.n_channels;

    const int n = live ? tab.n_samples[ch] : 0;
    const int n_frames = div_round_up(n, kGcFrameSamples);
    int32_t h1 = live ? tab.hist[2 * ch] : 0, h2 = live ? tab.hist[2 * ch + 1] : 0;
    const int16_t *my_coefs = coefs + (int64_t)(live ? ch : 0) * 16;

    // frames the warp as a whole still has to visit
    int warp_frames = n_frames;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_frames = max(warp_frames, __shfl_xor_sync(0xFFFFFFFFu, warp_frames, o));
    const int f_hi_warp = min(frame_end, warp_frames);

    for (int tf = frame_begin; tf < f_hi_warp; tf += kDecTileFrames) {
        // ---- stage in: channel c's 128-byte row, one 4-byte word per lane
        for (int c = 0; c < 32; c++) {
            const int cc = ch0 + c;
            if (cc >= tab.n_channels) break;
            const int cn = tab.n_samples[cc];
            const int64_t cbytes = gc_sample_count_to_byte_count(cn);
            const int64_t b = (int64_t)tf * kGcFrameBytes + lane * 4;
            uint32_t v = 0;
            if (b < cbytes) v = __ldg(reinterpret_cast<const uint32_t *>(adpcm + tab.adpcm_off[cc] + b));
            in_tile[c * kDecInPitch + lane] = v;
        }
        __syncwarp();

        // ---- decode my channel's frames of this tile
        const uint32_t *mine = in_tile + lane * kDecInPitch;
        uint32_t *mine_out = out_tile + lane * kDecOutPitch;
        const int f_hi = min(min(frame_end, n_frames), tf + kDecTileFrames);
        for (int f = tf; f < f_hi; f++) {
            const int i = f - tf;
            const uint32_t w0 = mine[2 * i], w1 = mine[2 * i + 1];
            const uint32_t head = w0 & 0xFFu;
            const int32_t scale = (int32_t)((1u << (head & 0xFu)) * 2048u);
            const int p = (head >> 4) & 0xF;  // a hostile header may select pairs 8..15: the reference would throw;
            const int32_t c1 = my_coefs[(p & 7) * 2], c2 = my_coefs[(p & 7) * 2 + 1];  // we wrap to stay in bounds
            uint32_t packed = 0;
#pragma unroll
            for (int s = 0; s < 14; s++) {
                const int byte = 1 + s / 2;
                const uint32_t word = byte < 4 ? w0 : w1;
                const uint32_t nib = (word >> ((byte & 3) * 8 + ((s & 1) ? 0 : 4))) & 0xFu;
                const int32_t q = nibble_signed(nib);
                const int32_t guess = wadd(wmul(c1, h1), wmul(c2, h2));
                const int32_t out = clamp16(wadd(wadd(guess, wmul(scale, q)), 1024) >> 11);
                h2 = h1;
                h1 = out;
                if (s & 1) {
                    packed |= (uint32_t)(out & 0xFFFF) << 16;
                    mine_out[i * 7 + s / 2] = packed;
                } else {
                    packed = (uint32_t)(out & 0xFFFF);
                }
            }
        }
        __syncwarp();

        // ---- stage out: channel c's PCM, 4 bytes per lane per step, only the samples that exist
        for (int c = 0; c < 32; c++) {
            const int cc = ch0 + c;
            if (cc >= tab.n_channels) break;
            const int cn = tab.n_samples[cc];
            const int64_t s0 = (int64_t)tf * kGcFrameSamples;  // first sample of the tile
            int64_t avail = min((int64_t)cn, (int64_t)min(frame_end, tf + kDecTileFrames) * kGcFrameSamples) - s0;
            if (avail <= 0) continue;
            int16_t *dst = pcm + tab.pcm_off[cc] + s0;
            const uint32_t *srcw = out_tile + c * kDecOutPitch;
            for (int wi = lane; wi < kDecOutWords; wi += 32) {
                const int64_t s = (int64_t)wi * 2;
                if (s + 1 < avail) {
                    *reinterpret_cast<uint32_t *>(dst + s) = srcw[wi];
                } else if (s < avail) {
                    dst[s] = (int16_t)(srcw[wi] & 0xFFFFu);
                }
            }
        }
        __syncwarp();
    }

    if (live) {  // carried into the next time slice of the same call
        tab.hist[2 * ch] = (int16_t)h1;
        tab.hist[2 * ch + 1] = (int16_t)h2;
    }
}

void launch_gc_decode(const uint8_t *adpcm, const GcChannelTable &tab, const int16_t *coefs, int16_t *pcm,
                      int max_frames, int frame_begin, int frame_end, cudaStream_t stream)
{
    if (tab.n_channels <= 0 || max_frames <= 0) return;
    if (frame_begin >= frame_end || frame_begin >= max_frames) return;
    const int ch_per_block = kDecWarps * 32;
    int blocks = (tab.n_channels + ch_per_block - 1) / ch_per_block;
    size_t smem = (size_t)kDecWarps * (32 * kDecInPitch + 32 * kDecOutPitch) * sizeof(uint32_t);
    gc_decode_kernel<<<blocks, kDecWarps * 32, smem, stream>>>(adpcm, tab, coefs, pcm, frame_begin, frame_end);
}

}  // namespace vgb
