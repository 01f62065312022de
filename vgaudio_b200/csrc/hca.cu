// hca.cu — CRI HCA encoder on sm_100a.
//
// Replaces CriHcaEncoder.EncodeFrame and its 12 stages (Codecs/CriHca/CriHcaEncoder.cs:271-286, :420-858),
// CriHcaPacking.PackFrame (CriHcaPacking.cs:17-58, BitWriter.cs:26-98, Crc16.cs) and Mdct.RunMdct/Dct4
// (Utilities/Mdct.cs:63-181) for NON-LOOPING streams (frame k of a stream is the k-th 1024-sample window of its PCM
// followed by zeros, CriHcaFormat.cs:53-81 + CriHcaEncoder.cs:192-242).
//
// Frames are independent given the raw PCM (the MDCT overlap is the previous 128 raw samples), so ONE CTA OF 128
// THREADS OWNS ONE (stream, frame): massive parallelism across frames, stages inside the CTA separated by barriers.
//   mdct        2 x 64 threads run two 128-point DCT-IV at a time (6 radix-2 stages in shared memory), fp64, same
//               operation order as the reference (a*cos + b*sin as mul, mul, add - no FMA)
//   scale       one thread per band: max |coef| over 8 subframes -> FindScaleFactor (binary search) -> scaled spectra
//   allocation  CalculateUsedBits is an integer sum -> block reduction; the two binary searches (noise level,
//               evaluation boundary) run ~16 probes of it
//   order-sensitive fp64 sums (intensity stereo energies, HFR group averages) stay sequential inside one thread each
//   pack        one thread appends the bits (MSB first) and computes the CRC; the frame leaves with coalesced stores
// fp64 throughout like the reference (A.14); with identical tables and operation order the frames are byte-identical
// to the oracle.  ALU/latency bound: ~135 ops per sample, algorithmic traffic 2 B/sample in +
// frame_size/1024 B/sample out.
#include "common.cuh"
#include "kernels.h"

namespace vgb {

namespace {

constexpr int kSub = 8, kBins = 128, kFrame = 1024;

__device__ __forceinline__ double dclamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// CalculateResolution (CriHcaPacking.cs:60-69)
__device__ __forceinline__ int hca_resolution(const HcaTables &T, int scale_factor, int noise_level)
{
    if (scale_factor == 0) return 0;
    int pos = noise_level - 5 * scale_factor / 2 + 2;
    pos = min(max(pos, 0), 58);
    return T.scale_to_resolution[pos];
}

// FindScaleFactor (CriHcaEncoder.cs:691-709)
__device__ __forceinline__ int hca_find_scale_factor(const HcaTables &T, double value)
{
    unsigned low = 0, high = 63;
    while (low < high) {
        const unsigned mid = (low + high) / 2;
        if (T.dequantizer_scaling[mid] <= value) low = mid + 1;
        else high = mid;
    }
    return (int)low;
}

// bits one scaled coefficient costs at a resolution (CalculateUsedBits inner loops, :566-592)
__device__ __forceinline__ int hca_coef_bits(const HcaTables &T, int resolution, double scaled)
{
    if (resolution >= 8) {
        const int bits = T.quantized_max_bits[resolution] - 1;
        return bits + (fabs(scaled) >= T.dead_zone[resolution] ? 1 : 0);
    }
    const double inv = T.inv_step[resolution];
    const double up = inv + 1;
    const int down = (int)(inv + 0.5 - 8);
    const int q = (int)(scaled * inv + up) - down;
    return T.quantize_bits[resolution][q];
}

struct BlockSum {  // integer sum over the 128 threads of the CTA, result in every thread
    int *scratch;  // 4 ints of shared memory
    __device__ int operator()(int v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
        __syncthreads();
        return scratch[0] + scratch[1] + scratch[2] + scratch[3];
    }
};

struct BitAppender {  // BitWriter.Write on a zeroed buffer: MSB-first append (BitWriter.cs:26-70)
    uint8_t *buf;
    int pos, length_bits;
    bool overflow;
    __device__ void write(int value, int count)
    {
        if (count > length_bits - pos) { overflow = true; return; }
        for (int i = count - 1; i >= 0; i--) {
            if ((value >> i) & 1) buf[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
            pos++;
        }
    }
};

}  // namespace

// Dynamic shared memory layout per CTA (nch = channel count):
//   double spectra[nch][8][128]; double scaled[nch][128][8]; then the small per-channel state below.
struct HcaChannelState {
    int scale_factors[kBins];
    int resolution[kBins];
    int intensity[kSub];
    int hfr_scales[8];
    double hfr_group_avg[8];
    int header_bits, delta_bits, type, coded_count;
};

__global__ void __launch_bounds__(128)
hca_encode_kernel(const int16_t *__restrict__ pcm, const HcaStream *__restrict__ streams, HcaConfig cfg, HcaTables T,
                  uint8_t *__restrict__ frames_out, int32_t *__restrict__ status_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nch = cfg.channel_count;
    double *spectra = reinterpret_cast<double *>(smem_raw);                   // [nch][8][128]
    double *scaled = spectra + (size_t)nch * kSub * kBins;                    // [nch][128][8]
    double *work = scaled + (size_t)nch * kSub * kBins;                       // [2][128] DCT scratch (+ fold input)
    double *fold = work + 2 * kBins;                                          // [2][128]
    HcaChannelState *chs = reinterpret_cast<HcaChannelState *>(fold + 2 * kBins);  // [nch]
    uint8_t *frame_buf = reinterpret_cast<uint8_t *>(chs + nch);              // [frame_size] (+pad)
    int *red = reinterpret_cast<int *>(frame_buf + ((cfg.frame_size + 15) & ~15));  // [8] reduction / broadcast scratch

    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const HcaStream st = streams[s];
    const int k = blockIdx.x;  // frame index
    if (k >= st.frame_count) return;
    const BlockSum block_sum{red};

    // ---- channel set-up (CriHcaFrame ctor :18-33)
    if (tid < nch) {
        chs[tid].type = cfg.channel_type[tid];
        chs[tid].coded_count = cfg.channel_type[tid] == 2 ? cfg.base_band_count : cfg.base_band_count + cfg.stereo_band_count;
    }

    // ---- PcmToFloat (:845-858) + RunMdct (:834-843, Mdct.cs:63-92): two subframes at a time
    const int grp = tid >> 6, i = tid & 63;
    double *t = work + grp * kBins;
    double *in = fold + grp * kBins;
    for (int c = 0; c < nch; c++) {
        const int16_t *src = pcm + st.pcm_off + (int64_t)c * st.channel_stride;
        for (int sf2 = 0; sf2 < kSub; sf2 += 2) {
            const int sf = sf2 + grp;
            const int64_t base = (int64_t)k * kFrame + sf * kBins;  // first sample of this subframe
            auto sample = [&](int64_t idx) -> double {
                const int16_t v = (idx >= 0 && idx < st.sample_count) ? src[idx] : (int16_t)0;
                return (double)v * (1.0 / 32768.0);
            };
            {   // window + fold into the DCT input (Mdct.cs:77-85); `previous` = the 128 samples before this subframe
                const double a = T.window[64 - i - 1] * -sample(base + 64 + i);
                const double b = T.window[64 + i] * sample(base + 64 - i - 1);
                const double cc = T.window[i] * sample(base - kBins + i);
                const double d = T.window[kBins - i - 1] * sample(base - kBins + kBins - i - 1);
                in[i] = a - b;
                in[64 + i] = cc - d;
            }
            __syncthreads();
            {   // Dct4 pre-twiddle (Mdct.cs:137-147)
                const int i2 = i * 2;
                const double a = in[i2], b = in[kBins - 1 - i2];
                const double sn = T.sin_tab[7][i], cs = T.cos_tab[7][i];
                t[i2] = a * cs + b * sn;
                t[i2 + 1] = a * sn - b * cs;
            }
            __syncthreads();
#pragma unroll 1
            for (int stage = 0; stage < 6; stage++) {  // (Mdct.cs:148-175)
                const int block_bits = 6 - stage, half_bits = block_bits - 1;
                const int block_size = 1 << block_bits, block_half = 1 << half_bits;
                const int block = i >> half_bits, j = i & (block_half - 1);
                const int front = (block * block_size + j) * 2, back = front + block_size;
                const double a = t[front] - t[back];
                const double b = t[front + 1] - t[back + 1];
                const double sn = T.sin_tab[half_bits][j], cs = T.cos_tab[half_bits][j];
                const double f0 = t[front] + t[back], f1 = t[front + 1] + t[back + 1];
                t[front] = f0;
                t[front + 1] = f1;
                t[back] = a * cs + b * sn;
                t[back + 1] = a * sn - b * cs;
                __syncthreads();
            }
            double *out = spectra + ((size_t)c * kSub + sf) * kBins;
            out[i] = t[T.shuffle[i]] * T.mdct_scale;            // (Mdct.cs:177-180)
            out[64 + i] = t[T.shuffle[64 + i]] * T.mdct_scale;
            __syncthreads();
        }
    }

    // ---- EncodeIntensityStereo (:711-764): the energy sums are order-sensitive fp64 -> one thread per (pair, sf)
    if (cfg.stereo_band_count > 0) {
        for (int c = 0; c < nch; c++) {
            if (chs[c].type != 1) continue;
            double *l = spectra + ((size_t)c * kSub) * kBins, *r = spectra + ((size_t)(c + 1) * kSub) * kBins;
            if (tid < kSub) {
                const int sf = tid;
                double el = 0, er = 0, et = 0;
                for (int b = cfg.base_band_count; b < cfg.total_band_count; b++) {
                    el += fabs(l[sf * kBins + b]);
                    er += fabs(r[sf * kBins + b]);
                    et += fabs(l[sf * kBins + b] + r[sf * kBins + b]);
                }
                et *= 2;
                const double elr = er + el;
                const double stored = 2 * el / elr;
                double ratio = elr / et;
                ratio = dclamp(ratio, 0.5, T.sqrt2 / 2);
                int q = 1;
                if (er > 0 || el > 0) {
                    while (q < 13 && T.intensity_bounds[q] >= stored) q++;
                } else {
                    q = 0;
                    ratio = 1;
                }
                chs[c + 1].intensity[sf] = q;
                for (int b = cfg.base_band_count; b < cfg.total_band_count; b++) {
                    l[sf * kBins + b] = (l[sf * kBins + b] + r[sf * kBins + b]) * ratio;
                    r[sf * kBins + b] = 0;
                }
            }
        }
        __syncthreads();
    }

    // ---- CalculateScaleFactors (:673-689) + ScaleSpectra (:651-671): thread = band
    for (int c = 0; c < nch; c++) {
        const int b = tid;
        int sfac = 0;
        if (b < chs[c].coded_count) {
            double mx = 0;
#pragma unroll
            for (int sf = 0; sf < kSub; sf++) {
                const double coeff = fabs(spectra[((size_t)c * kSub + sf) * kBins + b]);
                mx = coeff > mx ? coeff : mx;
            }
            sfac = hca_find_scale_factor(T, mx);
            const double qs = T.quantizer_scaling[sfac];
#pragma unroll
            for (int sf = 0; sf < kSub; sf++) {
                const double coeff = spectra[((size_t)c * kSub + sf) * kBins + b];
                scaled[((size_t)c * kBins + b) * kSub + sf] = sfac == 0 ? 0.0 : dclamp(coeff * qs, -0.999999999999, 0.999999999999);
            }
        }
        chs[c].scale_factors[b] = sfac;
        chs[c].resolution[b] = 0;
    }
    __syncthreads();

    // ---- CalculateHfrGroupAverages (:766-793) + CalculateHfrScale (:795-832): sequential sums, thread = (channel, group)
    if (cfg.hfr_group_count > 0) {
        const int c = tid >> 3, group = tid & 7;
        if (c < nch && group < cfg.hfr_group_count && chs[c].type != 2) {
            const int start = cfg.stereo_band_count + cfg.base_band_count;
            {
                double sum = 0.0;
                int count = 0;
                int band = start + group * cfg.bands_per_hfr_group;
                for (int ii = 0; ii < cfg.bands_per_hfr_group && band < kBins; band++, ii++) {
                    for (int sf = 0; sf < kSub; sf++) sum += fabs(spectra[((size_t)c * kSub + sf) * kBins + band]);
                    count += kSub;
                }
                chs[c].hfr_group_avg[group] = sum / count;
            }
            {
                const int hfr_bands = min(cfg.hfr_band_count, cfg.total_band_count - cfg.hfr_band_count);
                double sum = 0.0;
                int count = 0;
                int band = group * cfg.bands_per_hfr_group;
                for (int ii = 0; ii < cfg.bands_per_hfr_group && band < hfr_bands; band++, ii++) {
                    for (int sf = 0; sf < kSub; sf++) sum += fabs(scaled[((size_t)c * kBins + (start - band - 1)) * kSub + sf]);
                    count += kSub;
                }
                const double avg = sum / count;
                double g = chs[c].hfr_group_avg[group];
                if (avg > 0.0) {
                    const double inv = 1.0 / avg;
                    g *= inv < T.sqrt2 ? inv : T.sqrt2;
                }
                chs[c].hfr_group_avg[group] = g;
                chs[c].hfr_scales[group] = hca_find_scale_factor(T, g);
            }
        }
        __syncthreads();
    }

    // ---- CalculateFrameHeaderLength (:599-649): integers; one thread per channel
    auto header_lengths = [&]() {
        if (tid < nch) {
            HcaChannelState &ch = chs[tid];
            bool empty = true;
            for (int b = 0; b < ch.coded_count; b++)
                if (ch.scale_factors[b] != 0) { empty = false; break; }
            if (empty) {
                ch.header_bits = 3;
                ch.delta_bits = 0;
            } else {
                int min_delta_bits = 6;
                int min_length = 3 + 6 * ch.coded_count;
                for (int delta_bits = 1; delta_bits < 6; delta_bits++) {
                    const int max_delta = (1 << (delta_bits - 1)) - 1;
                    int length = 3 + 6;
                    for (int band = 1; band < ch.coded_count; band++) {
                        const int delta = ch.scale_factors[band] - ch.scale_factors[band - 1];
                        length += abs(delta) > max_delta ? delta_bits + 6 : delta_bits;
                    }
                    if (length < min_length) { min_length = length; min_delta_bits = delta_bits; }
                }
                ch.header_bits = min_length;
                ch.delta_bits = min_delta_bits;
            }
            if (ch.type == 2) ch.header_bits += 32;
            else if (cfg.hfr_group_count > 0) ch.header_bits += 6 * cfg.hfr_group_count;
        }
        __syncthreads();
    };
    header_lengths();

    // ---- CalculateUsedBits (:554-597): integer sum over channels x bands x subframes -> block reduction
    auto used_bits = [&](int noise_level, int eval_boundary) -> int {
        int mine = 0;
        for (int c = 0; c < nch; c++) {
            const int b = tid;
            if (b < chs[c].coded_count) {
                const int noise = b < eval_boundary ? noise_level - 1 : noise_level;
                const int resolution = hca_resolution(T, chs[c].scale_factors[b], noise);
#pragma unroll
                for (int sf = 0; sf < kSub; sf++) mine += hca_coef_bits(T, resolution, scaled[((size_t)c * kBins + b) * kSub + sf]);
            }
            if (tid == 0) mine += chs[c].header_bits;
        }
        return 16 + 16 + 16 + block_sum(mine);
    };

    // ---- CalculateNoiseLevel (:457-485) with BinarySearchLevel (:502-523); every thread runs the same control flow
    const int available = cfg.frame_size * 8;
    auto search_level = [&]() -> int {
        int low = 0, high = 255, mid_value = 0;
        while (low != high) {
            const int mid = (low + high) / 2;
            mid_value = used_bits(mid, 0);
            if (mid_value > available) low = mid + 1;
            else high = mid;
        }
        return (low == 255 && mid_value > available) ? -1 : low;
    };
    int level = search_level();
    int highest = cfg.base_band_count + cfg.stereo_band_count - 1;
    int status = 0;
    while (level < 0) {
        highest -= 2;
        if (highest < 0) { status = VGB_HCA_BITRATE_TOO_LOW; break; }  // InvalidDataException("Bitrate is set too low.")
        if (tid < nch) {
            chs[tid].scale_factors[highest + 1] = 0;
            chs[tid].scale_factors[highest + 2] = 0;
        }
        __syncthreads();
        header_lengths();
        level = search_level();
    }
    const int noise_level = status ? 0 : level;

    // ---- CalculateEvaluationBoundary (:487-500) with BinarySearchBoundary (:525-552)
    int eval_boundary = 0;
    if (!status && noise_level != 0) {
        int low = 0, high = 127;
        while (abs(high - low) > 1) {
            const int mid = (low + high) / 2;
            const int mid_value = used_bits(noise_level, mid);
            if (available < mid_value) high = mid - 1;
            else low = mid;
        }
        int found;
        if (low == high) found = low < 127 ? low : -1;
        else found = used_bits(noise_level, high) > available ? low : high;
        if (found < 0) status = VGB_HCA_NOT_IMPLEMENTED;  // NotImplementedException (:499)
        else eval_boundary = found;
    }

    // ---- CalculateFrameResolutions (:441-455) + QuantizeSpectra (:420-439): thread = band; quantised values reuse
    // the `spectra` storage as ints (the spectra themselves are no longer needed)
    int *quantized = reinterpret_cast<int *>(spectra);  // [nch][8][128] ints inside the first half of the doubles
    __syncthreads();
    for (int c = 0; c < nch; c++) {
        const int b = tid;
        int resolution = 0;
        if (b < chs[c].coded_count)
            resolution = hca_resolution(T, chs[c].scale_factors[b], b < eval_boundary ? noise_level - 1 : noise_level);
        chs[c].resolution[b] = resolution;
    }
    __syncthreads();
    for (int c = 0; c < nch; c++) {
        const int b = tid;
        int q[kSub];
        const int resolution = chs[c].resolution[b];
        const double inv = T.inv_step[resolution];
        const double up = inv + 1;
        const int down = (int)(inv + 0.5);
#pragma unroll
        for (int sf = 0; sf < kSub; sf++)
            q[sf] = b < chs[c].coded_count ? (int)(scaled[((size_t)c * kBins + b) * kSub + sf] * inv + up) - down : 0;
        __syncthreads();  // every thread has read what it needs from `spectra` long ago; ints alias its first half
#pragma unroll
        for (int sf = 0; sf < kSub; sf++) quantized[((size_t)c * kSub + sf) * kBins + b] = q[sf];
    }
    for (int b = tid; b < cfg.frame_size; b += blockDim.x) frame_buf[b] = 0;
    __syncthreads();

    // ---- PackFrame (CriHcaPacking.cs:17-58): one thread appends the bits, then the CRC (Crc16.cs)
    if (tid == 0) {
        BitAppender w{frame_buf, 0, cfg.frame_size * 8, false};
        w.write(0xffff, 16);
        w.write(noise_level, 9);
        w.write(eval_boundary, 7);
        for (int c = 0; c < nch; c++) {
            const HcaChannelState &ch = chs[c];
            const int delta_bits = ch.delta_bits;  // WriteScaleFactors (:262-295)
            w.write(delta_bits, 3);
            if (delta_bits == 6) {
                for (int b = 0; b < ch.coded_count; b++) w.write(ch.scale_factors[b], 6);
            } else if (delta_bits != 0) {
                w.write(ch.scale_factors[0], 6);
                const int max_delta = (1 << (delta_bits - 1)) - 1;
                const int escape = (1 << delta_bits) - 1;
                for (int b = 1; b < ch.coded_count; b++) {
                    const int delta = ch.scale_factors[b] - ch.scale_factors[b - 1];
                    if (abs(delta) > max_delta) {
                        w.write(escape, delta_bits);
                        w.write(ch.scale_factors[b], 6);
                    } else {
                        w.write(max_delta + delta, delta_bits);
                    }
                }
            }
            if (ch.type == 2) {
                for (int sf = 0; sf < kSub; sf++) w.write(ch.intensity[sf], 4);
            } else if (cfg.hfr_group_count > 0) {
                for (int g = 0; g < cfg.hfr_group_count; g++) w.write(ch.hfr_scales[g], 6);
            }
        }
        for (int sf = 0; sf < kSub; sf++)  // WriteSpectra (:238-260)
            for (int c = 0; c < nch; c++) {
                const HcaChannelState &ch = chs[c];
                for (int b = 0; b < ch.coded_count; b++) {
                    const int resolution = ch.resolution[b];
                    const int q = quantized[((size_t)c * kSub + sf) * kBins + b];
                    if (resolution == 0) continue;
                    if (resolution < 8) {
                        w.write(T.quantize_value[resolution][q + 8], T.quantize_bits[resolution][q + 8]);
                    } else {
                        w.write(abs(q), T.quantized_max_bits[resolution] - 1);
                        if (q != 0) w.write(q > 0 ? 0 : 1, 1);
                    }
                }
            }
        if (w.overflow && !status) status = VGB_HCA_BIT_OVERFLOW;  // InvalidOperationException (BitWriter.cs:30-33)
        uint16_t crc = 0;  // WriteChecksum (:231-236)
        for (int b = 0; b < cfg.frame_size - 2; b++) crc = (uint16_t)((crc << 8) ^ T.crc_table[(crc >> 8) ^ frame_buf[b]]);
        frame_buf[cfg.frame_size - 2] = (uint8_t)(crc >> 8);
        frame_buf[cfg.frame_size - 1] = (uint8_t)crc;
        if (status) atomicCAS(status_out + s, 0, status);
    }
    __syncthreads();
    uint8_t *dst = frames_out + st.frames_off + (int64_t)k * cfg.frame_size;
    for (int b = tid; b < cfg.frame_size; b += blockDim.x) dst[b] = frame_buf[b];
}

size_t hca_encode_smem_bytes(const HcaConfig &cfg)
{
    const size_t nch = (size_t)cfg.channel_count;
    return 2 * nch * kSub * kBins * sizeof(double) + 4 * kBins * sizeof(double) + nch * sizeof(HcaChannelState) +
           (size_t)((cfg.frame_size + 15) & ~15) + 8 * sizeof(int) + 16;
}

cudaError_t launch_hca_encode(const int16_t *pcm, const HcaStream *streams, int n_streams, int max_frames,
                              const HcaConfig &cfg, const HcaTables &tables, uint8_t *frames_out, int32_t *status_out,
                              cudaStream_t stream)
{
    if (n_streams <= 0 || max_frames <= 0) return cudaSuccess;
    const size_t smem = hca_encode_smem_bytes(cfg);
    cudaError_t e = cudaFuncSetAttribute(hca_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)max_frames, (unsigned)n_streams);
    hca_encode_kernel<<<grid, 128, smem, stream>>>(pcm, streams, cfg, tables, frames_out, status_out);
    return cudaGetLastError();
}

}  // namespace vgb
