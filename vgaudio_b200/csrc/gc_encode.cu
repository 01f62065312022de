// gc_encode.cu — GC-ADPCM encoder on sm_100a.
//
// Replaces GcAdpcmEncoder.Encode / DspEncodeFrame / DspEncodeCoef (Codecs/GcAdpcm/GcAdpcmEncoder.cs:14-171).
//
// Dependence structure of the reference: frames of one channel are strictly serial (frame k+1 starts from the
// RECONSTRUCTED last two samples of frame k, :40-41,:80), channels are independent, and inside a frame the eight
// predictors are independent while the scale attempts of one predictor form a short chain (:127-170).
//
// Mapping: ONE WARP OWNS ONE CHANNEL.  Lane = predictor * 4 + candidate: the 8 predictors are searched in
// parallel and, for each, 4 consecutive scale powers are evaluated speculatively (every attempt is a pure function
// of (samples, history, coefs, scalePower), so the do/while chain can be replayed over finished attempts; if the
// chain would leave the 4-wide window or take the rare overflow "bump" (:166-168) the warp falls back to the
// literal loop).  The argmin over predictors (strict <, first wins, :66-76) is two REDUX.MIN on an exact integer
// key; the winner's two newest reconstructed samples are broadcast with one REDUX.OR.  PCM is staged through
// shared memory 16 frames at a time with coalesced 16-byte loads (prefetched one chunk ahead), ADPCM bytes are
// staged and written back 128 bytes at a time.
//
// The kernel is latency bound (a 14-step integer recurrence per frame), not HBM bound: see DESIGN.md §gc_encode.
#include "common.cuh"
#include "kernels.h"

namespace vgb {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kEncChunkFrames = 16;                                   // frames staged per chunk
constexpr int kEncChunkSamples = kEncChunkFrames * kGcFrameSamples;   // 224 samples = 448 B = 28 x 16 B
constexpr int kEncWarps = 2;                                          // channels per CTA

template <bool kGeneral>
struct GcTrial {
    uint32_t w0, w1;   // the 8 frame bytes as two little-endian words (byte 0 = header, filled by the winner)
    int32_t r1, r2;    // newest / second newest reconstructed sample
    int32_t over;      // maxOverflow (:132,:147-151)
    uint64_t err;      // TotalDistance (:163) - a sum of squared integers, exact in 64 bits (< 2^36)
    int32_t recon[kGeneral ? 14 : 1];  // full reconstruction, only for the independent-frames entry point
};

// The reference's cast chain (:142-144): int -> float32, divide by the power-of-two scale in float32 (exact, so a
// multiply by 2^-K gives the same float), widen, add/subtract the float32 literal 0.4999999f widened, truncate.
__device__ __forceinline__ int32_t gc_quantise(int32_t diff, float inv_scale)
{
    const float ratio = __fmul_rn(__int2float_rn(diff), inv_scale);
    const double wide = (double)ratio;
    const double half = (double)0.4999999f;
    return diff > 0 ? __double2int_rz(__dadd_rn(wide, half)) : __double2int_rz(__dsub_rn(wide, half));
}

// One pass of the do/while body (:129-164) at a fixed scalePower.
template <bool kGeneral>
__device__ __forceinline__ void gc_attempt(const int32_t (&x)[14], int n, int32_t h1, int32_t h2, int32_t c0, int32_t c1,
                                           int sp, GcTrial<kGeneral> &t)
{
    const int shift = sp + 11;
    const int32_t scale = (int32_t)(1u << shift);                    // (1 << scalePower) * 2048
    const float inv_scale = __int_as_float((127 - shift) << 23);     // 2^-shift
    int32_t r1 = h1, r2 = h2, over = 0;
    uint64_t err = 0;
    uint32_t w0 = 0, w1 = 0;
#pragma unroll
    for (int s = 0; s < 14; s++) {
        if (kGeneral && s >= n) break;
        const int32_t want = x[s] * 2048;
        const int32_t guess = wadd(wmul(r2, c1), wmul(r1, c0));
        const int32_t diff = wsub(want, guess);
        const int32_t raw = gc_quantise(diff, inv_scale);
        const int32_t q = clamp4(raw);
        over = max(over, abs(raw - q));
        const int32_t out = clamp16(wadd(wadd(guess, wmul(q, scale)), 1024) >> 11);
        const uint32_t miss = (uint32_t)(x[s] - out);
        err += (uint64_t)(miss * miss);  // (x - out)^2 < 2^32: the wrapped 32-bit product is the true value
        const int byte = 1 + s / 2, bit = (byte & 3) * 8 + ((s & 1) ? 0 : 4);
        if (byte < 4) w0 |= (uint32_t)(q & 15) << bit; else w1 |= (uint32_t)(q & 15) << bit;
        if (kGeneral) t.recon[s] = out;
        r2 = r1;
        r1 = out;
    }
    t.w0 = w0; t.w1 = w1; t.r1 = r1; t.r2 = r2; t.over = over; t.err = err;
}

// The literal do/while of DspEncodeCoef (:127-170), used when the speculative window does not cover the chain.
template <bool kGeneral>
__device__ __noinline__ void gc_try_predictor_literal(const int32_t (&x)[14], int n, int32_t h1, int32_t h2, int32_t c0,
                                                      int32_t c1, int sp_first, GcTrial<kGeneral> &t, int &sp_out)
{
    int sp = sp_first - 1;
    do {
        sp++;
        gc_attempt<kGeneral>(x, n, h1, h2, c0, c1, sp, t);
        for (int v = t.over + 8; v > 256; v >>= 1)
            if (++sp >= 12) sp = 11;
    } while (sp < 12 && t.over > 1);
    sp_out = sp;
}

// Residual of one sample against the RAW neighbours (:107-115) folded into an order-preserving key:
// larger |distance| wins, then the EARLIER sample (the reference keeps the first maximum: strict '>'), and the
// sign rides in bit 0 so the signed maxDistance can be rebuilt.
__device__ __forceinline__ uint32_t gc_peak_key(int32_t older, int32_t newer, int32_t cur, int32_t c0, int32_t c1, int s)
{
    const int32_t guess = wadd(wmul(older, c1), wmul(newer, c0)) / 2048;  // truncates toward zero (A.6)
    const int32_t diff = clamp16(wsub(cur, guess));
    return ((uint32_t)abs(diff) << 5) | ((uint32_t)(15 - s) << 1) | (diff < 0 ? 1u : 0u);
}

// DspEncodeFrame (:48-94) for one frame, executed by a full warp.  On return exactly one lane has is_winner set;
// its trial `t` / sp_final / predictor (lane >> 2) describe the chosen encoding.
template <bool kGeneral>
__device__ __forceinline__ void gc_frame_search(const int32_t (&x)[14], int n, int32_t h1, int32_t h2, int32_t c0,
                                                int32_t c1, int lane, bool &is_winner, GcTrial<kGeneral> &t,
                                                int &sp_final)
{
    const int cand = lane & 3;

    uint32_t key = 0;
#pragma unroll
    for (int s = 0; s < 14; s++) {
        if (kGeneral && s >= n) break;
        const int32_t older = s == 0 ? h2 : (s == 1 ? h1 : x[s >= 2 ? s - 2 : 0]);
        const int32_t newer = s == 0 ? h1 : x[s >= 1 ? s - 1 : 0];
        key = max(key, gc_peak_key(older, newer, x[s], c0, c1, s));
    }
    int32_t peak = (int32_t)(key >> 5);
    if (key & 1u) peak = -peak;

    // first scale guess (:118-124)
    int halvings = 0;
    while (halvings <= 12 && (peak > 7 || peak < -8)) {
        peak /= 2;
        halvings++;
    }
    const int sp_first = halvings <= 1 ? 0 : halvings - 1;  // value of scalePower in the first do/while pass

    const int sp = sp_first + cand;
    const bool valid = sp <= 12;
    gc_attempt<kGeneral>(x, n, h1, h2, c0, c1, valid ? sp : 12, t);

    const bool terminal = valid && (t.over <= 1 || sp >= 12);  // the while condition (:170) fails here
    const bool bump = valid && t.over > 248;                   // the overflow bump loop (:166-168) would run
    const uint32_t term_bits = __ballot_sync(kFull, terminal);
    const uint32_t group = (term_bits >> (lane & ~3)) & 0xFu;
    const bool slow = __any_sync(kFull, bump || group == 0u);

    sp_final = sp;
    bool pred_winner;
    if (!slow) {
        pred_winner = terminal && (group & ((1u << cand) - 1u)) == 0u;  // first candidate that ends the chain
    } else {
        pred_winner = cand == 0;
        if (pred_winner) gc_try_predictor_literal<kGeneral>(x, n, h1, h2, c0, c1, sp_first, t, sp_final);
    }

    // argmin of TotalDistance over the predictors, first minimum wins (:66-76): key = err * 8 + predictor
    const uint64_t full_key = pred_winner ? ((t.err << 3) | (uint64_t)(lane >> 2)) : ~0ull;
    const uint32_t hi = (uint32_t)(full_key >> 8);
    const uint32_t min_hi = __reduce_min_sync(kFull, hi);
    const uint32_t lo = (pred_winner && hi == min_hi) ? (uint32_t)(full_key & 0xFFu) : 0xFFFFFFFFu;
    const uint32_t min_lo = __reduce_min_sync(kFull, lo);
    is_winner = pred_winner && hi == min_hi && lo == min_lo;
}

// grid: one warp per channel; encodes frames [frame_begin, frame_end) of every channel, carrying the history
// in tab.hist between launches (frame_begin must be a multiple of 16).
__global__ void __launch_bounds__(kEncWarps * 32)
gc_encode_kernel(const int16_t *__restrict__ pcm, GcChannelTable tab, const int16_t *__restrict__ coefs,
                 uint8_t *__restrict__ adpcm, int frame_begin, int frame_end)
{
    __shared__ __align__(16) int16_t in_buf[kEncWarps][2][kEncChunkSamples];
    __shared__ __align__(16) uint8_t out_buf[kEncWarps][kEncChunkFrames * kGcFrameBytes];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kEncWarps + warp;
    if (ch >= tab.n_channels) return;

    const int n_enc = tab.enc_count[ch];
    const int n_frames = div_round_up(n_enc, kGcFrameSamples);
    const int f_hi = min(frame_end, n_frames);
    if (frame_begin >= f_hi) return;
    const int total_bytes = gc_sample_count_to_byte_count(n_enc);

    const int16_t *src = pcm + tab.pcm_off[ch];
    uint8_t *dst = adpcm + tab.adpcm_off[ch];
    const int pred = lane >> 2;
    const int32_t c0 = coefs[(int64_t)ch * 16 + 2 * pred];
    const int32_t c1 = coefs[(int64_t)ch * 16 + 2 * pred + 1];
    int32_t h1 = tab.hist[2 * ch], h2 = tab.hist[2 * ch + 1];

    // 16-byte vector `lane` of a chunk, zero beyond the encoded sample count (GcAdpcmEncoder.cs:32-34)
    auto load_vec = [&](int chunk_frame) -> uint4 {
        uint4 q = make_uint4(0, 0, 0, 0);
        const int64_t s = (int64_t)chunk_frame * kGcFrameSamples + lane * 8;
        if (lane < kEncChunkSamples / 8 && chunk_frame < f_hi && s < n_enc) {
            q = __ldg(reinterpret_cast<const uint4 *>(src + s));
            const int valid = (int)min((int64_t)8, (int64_t)n_enc - s);
            if (valid < 8) {
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (2 * i >= valid) w[i] = 0;
                    else if (2 * i + 1 >= valid) w[i] &= 0xFFFFu;
                }
                q = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        return q;
    };

    int buf = 0;
    {
        const uint4 first = load_vec(frame_begin);
        if (lane < kEncChunkSamples / 8) reinterpret_cast<uint4 *>(in_buf[warp][0])[lane] = first;
        __syncwarp();
    }

    for (int cf = frame_begin; cf < f_hi; cf += kEncChunkFrames) {
        const uint4 next = load_vec(cf + kEncChunkFrames);  // prefetch; consumed after this chunk
        const int frames_here = min(kEncChunkFrames, f_hi - cf);

        for (int i = 0; i < frames_here; i++) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(&in_buf[warp][buf][i * kGcFrameSamples]);
            int32_t x[14];
#pragma unroll
            for (int j = 0; j < 7; j++) {
                const uint32_t u = w[j];  // same address in every lane: shared-memory broadcast
                x[2 * j] = (int32_t)(int16_t)(u & 0xFFFFu);
                x[2 * j + 1] = (int32_t)(int16_t)(u >> 16);
            }

            bool is_winner;
            int sp_final;
            GcTrial<false> t;
            gc_frame_search<false>(x, 14, h1, h2, c0, c1, lane, is_winner, t, sp_final);

            if (is_winner) {
                const uint32_t head = (uint32_t)((pred << 4) | (sp_final & 0xF));  // CombineNibbles (:83)
                *reinterpret_cast<uint2 *>(&out_buf[warp][i * kGcFrameBytes]) = make_uint2(t.w0 | head, t.w1);
            }
            const uint32_t packed = __reduce_or_sync(
                kFull, is_winner ? (((uint32_t)t.r1 & 0xFFFFu) | ((uint32_t)t.r2 << 16)) : 0u);
            h1 = (int32_t)(int16_t)(packed & 0xFFFFu);  // pcmBuffer[1] = pcmBuffer[15] (:41)
            h2 = (int32_t)(int16_t)(packed >> 16);      // pcmBuffer[0] = pcmBuffer[14] (:40)
        }
        __syncwarp();

        // write the chunk's bytes; only the channel's last frame can be partial (:38)
        {
            const int64_t byte0 = (int64_t)cf * kGcFrameBytes;
            const int bytes_here = (int)min((int64_t)frames_here * kGcFrameBytes, (int64_t)total_bytes - byte0);
            if (lane < 8) {
                const int b = lane * 16;
                if (b + 16 <= bytes_here) {
                    *reinterpret_cast<uint4 *>(dst + byte0 + b) = *reinterpret_cast<const uint4 *>(&out_buf[warp][b]);
                } else {
                    for (int j = b; j < bytes_here; j++) dst[byte0 + j] = out_buf[warp][j];
                }
            }
        }
        buf ^= 1;
        if (lane < kEncChunkSamples / 8) reinterpret_cast<uint4 *>(in_buf[warp][buf])[lane] = next;
        __syncwarp();
    }

    if (lane == 0) {
        tab.hist[2 * ch] = (int16_t)h1;
        tab.hist[2 * ch + 1] = (int16_t)h2;
    }
}

// DspEncodeFrame for independent frames: one warp per frame (IDspTool.DspEncodeFrame / GcAdpcmAlignment use).
__global__ void __launch_bounds__(128)
gc_encode_frames_kernel(int16_t *__restrict__ pcm_in_out, const int32_t *__restrict__ sample_count,
                        const int16_t *__restrict__ coefs, int n_frames, uint8_t *__restrict__ adpcm_out)
{
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_frames) return;
    int16_t *io = pcm_in_out + (int64_t)f * 16;
    const int n = sample_count ? min(max(sample_count[f], 0), 14) : 14;
    int32_t x[14];
#pragma unroll
    for (int j = 0; j < 14; j++) x[j] = io[2 + j];
    const int32_t h2 = io[0], h1 = io[1];
    const int pred = lane >> 2;
    const int32_t c0 = coefs[(int64_t)f * 16 + 2 * pred], c1 = coefs[(int64_t)f * 16 + 2 * pred + 1];

    bool is_winner;
    int sp_final;
    GcTrial<true> t;
    gc_frame_search<true>(x, n, h1, h2, c0, c1, lane, is_winner, t, sp_final);
    __syncwarp();  // every lane has read io[] before the winner rewrites it
    if (is_winner) {
        const uint32_t head = (uint32_t)((pred << 4) | (sp_final & 0xF));
        uint32_t w[2] = {t.w0 | head, t.w1};
        for (int j = 0; j < 8; j++) adpcm_out[(int64_t)f * 8 + j] = (uint8_t)(w[j >> 2] >> ((j & 3) * 8));
#pragma unroll
        for (int j = 0; j < 14; j++)
            if (j < n) io[2 + j] = (int16_t)t.recon[j];
    }
}

void launch_gc_encode(const int16_t *pcm, const GcChannelTable &tab, const int16_t *coefs, uint8_t *adpcm,
                      int max_frames, int frame_begin, int frame_end, cudaStream_t stream)
{
    if (tab.n_channels <= 0 || max_frames <= 0) return;
    if (frame_begin >= frame_end || frame_begin >= max_frames) return;
    int blocks = (tab.n_channels + kEncWarps - 1) / kEncWarps;
    gc_encode_kernel<<<blocks, kEncWarps * 32, 0, stream>>>(pcm, tab, coefs, adpcm, frame_begin, frame_end);
}

void launch_gc_encode_frames(int16_t *pcm_in_out, const int32_t *sample_count, const int16_t *coefs, int n_frames,
                             uint8_t *adpcm_out, cudaStream_t stream)
{
    if (n_frames <= 0) return;
    int blocks = (n_frames + 3) / 4;
    gc_encode_frames_kernel<<<blocks, 128, 0, stream>>>(pcm_in_out, sample_count, coefs, n_frames, adpcm_out);
}

}  // namespace vgb
