// gc_coefs.cu — GC-ADPCM coefficient analysis on sm_100a.
//
// Replaces GcAdpcmCoefficients.CalculateCoefficients (Codecs/GcAdpcm/GcAdpcmCoefficients.cs:9-110) for a whole
// batch of channels.  The reference is one serial fp64 loop per channel; here it is split where the data
// dependences allow it:
//
//   gc_coef_frames_kernel   phase 1 (:40-61): one THREAD per 14-sample frame, all channels x frames in flight.
//                           Each frame is a pure function of 16 samples (two of history).  The five integer
//                           correlations are exact in int64, the 2x2 LU / solve / reflection step is the
//                           reference's fp64 sequence.  The record is stored already pushed through MatrixFilter
//                           (:285-305), because that direct-form pair is the only thing phase 2 ever reads
//                           (it equals ContrastVectors' `val` terms bit for bit - see DESIGN.md §gc_coef_refine).
//                           HBM-bound by design: 2 B/sample in, 16 B + 1 bit per frame out.
//
//   gc_coef_refine_kernel   phase 2 (:63-108): one WARP per channel.  Nearest-centroid search is parallel over
//                           32 records; the fp64 accumulations are applied strictly in record order
//                            by 16 accumulator lanes (8 buckets x 2 components), because any
//                           tree reduction would change the roundings and can flip a 16-bit coefficient.  Records
//                           are compacted per bucket (ballot + popc) so each lane's DADD chain only contains its
//                           own bucket's records.
#include "common.cuh"
#include "kernels.h"

namespace vgb {

constexpr int kP1Threads = 256;             // frames per CTA tile
constexpr int kP1TileSamples = kP1Threads * kGcFrameSamples;  // 3584 samples = 7168 B = 448 x 16 B

// ---------------------------------------------------------------------------------------------------------
// phase 1: per-frame record
// ---------------------------------------------------------------------------------------------------------

// x[0..1] = the two samples before the frame (older first), x[2..15] = the frame (zero padded).
// Returns true when the reference would append a record (GcAdpcmCoefficients.cs:46-57) and sets d1,d2 to
// MatrixFilter(record)[1..2].
__device__ __forceinline__ bool gc_frame_direct(const int32_t (&x)[16], double &d1, double &d2)
{
    // InnerProductMerge (:112-120): -sum x[t-lag]*x[t].  Every product and partial sum is an integer below
    // 2^34, so the reference's sequential double accumulation is exact and equals the int64 sum.
    long long s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int t = 0; t < 14; t++) {
        long long a = x[t + 2];
        s0 += a * x[t + 2];
        s1 += a * x[t + 1];
        s2 += a * x[t];
    }
    double v0 = 0.0 - (double)s0;  // "0.0 - s" keeps +0.0 for an all-zero frame like the reference's loop
    double v1 = 0.0 - (double)s1;
    double v2 = 0.0 - (double)s2;
    if (!(fabs(v0) > 10.0)) return false;  // :46

    // OuterProductMerge (:122-131) by sliding the lag-0/lag-1 sums one and two samples back (exact integers).
    long long c11 = s0 - (long long)x[15] * x[15] + (long long)x[1] * x[1];
    long long c22 = c11 - (long long)x[14] * x[14] + (long long)x[0] * x[0];
    long long c12 = s1 - (long long)x[15] * x[14] + (long long)x[1] * x[0];
    double m11 = (double)c11, m12 = (double)c12, m21 = (double)c12, m22 = (double)c22;

    // AnalyzeRanges (:133-208), 2x2 case written out.
    double big1 = fmax(fabs(m11), fabs(m12));
    if (big1 < 4.9406564584124654e-324) return false;  // double.Epsilon (A.2)
    double big2 = fmax(fabs(m21), fabs(m22));
    if (big2 < 4.9406564584124654e-324) return false;
    double inv1 = 1.0 / big1, inv2 = 1.0 / big2;

    int perm1 = 0;
    {
        double best = 0.0;
        double t = fabs(m11) * inv1;
        if (t >= best) { best = t; perm1 = 1; }
        t = fabs(m21) * inv2;
        if (t >= best) { best = t; perm1 = 2; }
    }
    if (perm1 == 2) {  // row swap (:174-183)
        double t = m21; m21 = m11; m11 = t;
        t = m22; m22 = m12; m12 = t;
    }
    {
        double t = 1.0 / m11;  // :187-192
        m21 *= t;
    }
    m22 = m22 - m21 * m12;  // :161-165 for column 2 (the pivot search there can only pick row 2)

    {
        double lo = 1.0e10, hi = 0.0;  // :195-207
        double t = fabs(m11);
        if (t < lo) lo = t;
        if (t > hi) hi = t;
        t = fabs(m22);
        if (t < lo) lo = t;
        if (t > hi) hi = t;
        if (lo / hi < 1.0e-10) return false;
    }

    // BidirectionalFilter (:210-237)
    double a = (perm1 == 2) ? v2 : v1;  // forward substitution with the row permutation
    double b = (perm1 == 2) ? v1 : v2;
    if (a != 0.0) b = b - a * m21;
    b = b / m22;
    a = (a - b * m12) / m11;

    // QuadraticMerge (:239-255)
    double den = 1.0 - (b * b);
    if (den == 0.0) return false;
    double k1 = (a - (a * b)) / den;
    if (fabs(k1) > 1.0) return false;
    double k2 = b;

    // FinishRecord (:257-269)
    if (k1 >= 1.0) k1 = 0.9999999999; else if (k1 <= -1.0) k1 = -0.9999999999;
    if (k2 >= 1.0) k2 = 0.9999999999; else if (k2 <= -1.0) k2 = -0.9999999999;
    double rec1 = (k2 * k1) + k1;
    double rec2 = k2;

    // MatrixFilter (:285-305)
    double n2 = -rec2, n1 = -rec1;
    double den2 = 1.0 - (n2 * n2);
    double p11 = ((n2 * n1) + n1) / den2;
    d1 = 0.0 + p11 * 1.0;
    d2 = (0.0 + n1 * d1) + n2 * 1.0;
    return true;
}

// grid: x = channel, y = tile of 256 frames (strided if the channel has more tiles than gridDim.y)
__global__ void __launch_bounds__(kP1Threads)
gc_coef_frames_kernel(const int16_t *__restrict__ pcm, GcChannelTable tab, double2 *__restrict__ records,
                      uint32_t *__restrict__ accept_mask, int frame_begin, int frame_end)
{
    __shared__ __align__(16) int16_t tile[8 + kP1TileSamples];  // tile[6..7] = two history samples, tile[8..] = frames

    const int ch = blockIdx.x;
    const int n = tab.n_samples[ch];
    const int n_frames = div_round_up(n, kGcFrameSamples);
    const int f_hi = min(frame_end, n_frames);
    const int16_t *src = pcm + tab.pcm_off[ch];
    double2 *rec = records + tab.rec_off[ch];
    uint32_t *mask = accept_mask + (tab.rec_off[ch] >> 5);

    for (int f0 = frame_begin + (int)blockIdx.y * kP1Threads; f0 < f_hi; f0 += (int)gridDim.y * kP1Threads) {
        const int64_t s0 = (int64_t)f0 * kGcFrameSamples;  // first sample of the tile; multiple of 8 -> 16 B aligned
        // coalesced 16-byte loads; samples at or beyond n read as zero (the reference zero-pads, :42-43)
        for (int v = threadIdx.x; v < kP1TileSamples / 8; v += kP1Threads) {
            int64_t s = s0 + (int64_t)v * 8;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (s < n) {
                q = __ldg(reinterpret_cast<const uint4 *>(src + s));
                int valid = (int)min((int64_t)8, (int64_t)n - s);
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
            *reinterpret_cast<uint4 *>(&tile[8 + v * 8]) = q;
        }
        if (threadIdx.x < 2) {
            int64_t s = s0 - 2 + threadIdx.x;
            tile[6 + threadIdx.x] = (s >= 0 && s < n) ? src[s] : (int16_t)0;
        }
        __syncthreads();

        const int f = f0 + threadIdx.x;
        bool ok = false;
        double d1 = 0.0, d2 = 0.0;
        if (f < f_hi) {
            // 16 samples = 8 aligned 32-bit words at byte offset 28*t + 12: stride of 7 words -> conflict free
            const uint32_t *w = reinterpret_cast<const uint32_t *>(&tile[6 + threadIdx.x * kGcFrameSamples]);
            int32_t x[16];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t u = w[i];
                x[2 * i] = (int32_t)(int16_t)(u & 0xFFFFu);
                x[2 * i + 1] = (int32_t)(int16_t)(u >> 16);
            }
            ok = gc_frame_direct(x, d1, d2);
            rec[f] = make_double2(d1, d2);
        }
        uint32_t bits = __ballot_sync(0xFFFFFFFFu, ok);
        if ((threadIdx.x & 31) == 0 && f < f_hi) mask[f >> 5] = bits;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// phase 2: ordered mean + three split/refine generations, one warp per channel
// ---------------------------------------------------------------------------------------------------------

struct Centroid {
    double e0, e1, e2;  // ContrastVectors' val1, val2, val3 for this centroid (:338-340)
};

// FinishRecord(double[], double[]) (:271-283)
__device__ __forceinline__ void gc_finish(double k1, double k2, double &o1, double &o2)
{
    if (k1 >= 1.0) k1 = 0.9999999999; else if (k1 <= -1.0) k1 = -0.9999999999;
    if (k2 >= 1.0) k2 = 0.9999999999; else if (k2 <= -1.0) k2 = -0.9999999999;
    o1 = (k2 * k1) + k1;
    o2 = k2;
}

// MergeFinishRecord (:307-333) for a 3-vector src -> centroid (1, o1, o2)
__device__ __forceinline__ void gc_centroid_from_mean(double s0, double s1, double s2, double &o1, double &o2)
{
    double err = s0;
    // i = 1
    double acc = 0.0;
    double t1 = (err > 0.0) ? (-(acc + s1) / err) : 0.0;
    double k1 = t1;
    err *= 1.0 - (t1 * t1);
    // i = 2
    acc = 0.0;
    acc += t1 * s1;
    double t2 = (err > 0.0) ? (-(acc + s2) / err) : 0.0;
    double k2 = t2;
    // (dst[1] += dst[2]*dst[1] and the last err update do not reach the output)
    gc_finish(k1, k2, o1, o2);
}

// (short) rounding of :94-108; Math.Round is half-to-even = rint.
__device__ __forceinline__ int16_t gc_quantise_coef(double v)
{
    double d = -v * 2048.0;
    if (d > 0.0) return d > 32767.0 ? (int16_t)32767 : (int16_t)__double2int_rn(d);
    if (d < -32768.0) return (int16_t)-32768;
    if (d != d) return 0;
    return (int16_t)__double2int_rn(d);
}

constexpr int kRefineWarps = 4;

// Per-bucket queues of one block of 32 records: record lanes append their direct-form pair to the queue of the
// bucket they were assigned to (rank = number of earlier records of the block in the same bucket, from a ballot),
// so an accumulator lane walks ONLY its bucket's records, still in record order.
struct __align__(16) RefineQueues {
    double2 q[8][32];
};

// One pass over all records of a channel with COUNT centroids (COUNT == 0: the plain ordered mean of :63-76, every
// record goes to bucket 0).  On return lane (bucket*2 + comp), bucket < max(COUNT,1), holds the ordered sum of that
// component over the bucket's records in `acc` and the record count in `hits`.
template <int COUNT>
__device__ __forceinline__ void gc_refine_pass(int lane, int n_frames, int n_blocks, const double2 *__restrict__ rec,
                                               const uint32_t *__restrict__ mask, RefineQueues *queues,
                                               const double (*best)[3], double &acc, int &hits)
{
    constexpr int NB = COUNT > 0 ? COUNT : 1;
    const int my_bucket = lane >> 1, my_comp = lane & 1;
    const uint32_t lanes_below = (1u << lane) - 1u;

    // per-centroid constants of ContrastVectors (:338-340)
    double e0[NB], e1[NB], e2[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        const double a0 = best[i][0], a1 = best[i][1], a2 = best[i][2];
        e0[i] = (a0 * a0) + (a1 * a1) + (a2 * a2);
        e1[i] = (a0 * a1) + (a1 * a2);
        e2[i] = a0 * a2;
    }

    // classify block b (nearest centroid, parallel over its 32 records) and append to the bucket queues of slot b&1;
    // bits[k] = which lanes went to bucket k
    auto stage = [&](int b, double2 r, uint32_t okbits, uint32_t (&bits)[NB]) {
        const bool ok = ((okbits >> lane) & 1u) && (b * 32 + lane < n_frames);
        int pick = 0;
        if (COUNT > 1) {
            // ContrastVectors (:335-342); its `val` is r.x and (-rec1*val - rec2) is r.y
            const double ta = 2.0 * r.x, tb = 2.0 * r.y;
            // all distances first (independent), then a tournament instead of the reference's sequential scan
            // (:373-381: `if (tempVal < value)`, value starting at 1.0e30): a later candidate replaces an earlier one
            // only when strictly smaller, so ties keep the lower index - the same first-minimum rule, depth log2(NB)
            double d[NB];
            int idx[NB];
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const double di = e0[i] + (ta * e1[i]) + (tb * e2[i]);
                d[i] = di < 1.0e30 ? di : __longlong_as_double(0x7FF0000000000000ll);  // never wins (also NaN), as in the scan
                idx[i] = i;
            }
#pragma unroll
            for (int step = 1; step < NB; step *= 2) {
#pragma unroll
                for (int i = 0; i + step < NB; i += 2 * step) {
                    const bool right = d[i + step] < d[i];
                    d[i] = right ? d[i + step] : d[i];
                    idx[i] = right ? idx[i + step] : idx[i];
                }
            }
            pick = d[0] < 1.0e30 ? idx[0] : 0;  // nothing below the initial 1.0e30: the index stays 0
        }
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < NB; k++) {
            bits[k] = __ballot_sync(0xFFFFFFFFu, ok && pick == k);
            mine = pick == k ? bits[k] : mine;
        }
        if (ok) queues[b & 1].q[pick][__popc(mine & lanes_below)] = r;
    };
    auto fetch = [&](int b, double2 &r, uint32_t &okbits) {
        r = make_double2(0.0, 0.0);
        okbits = 0;
        if (b < n_blocks) {
            okbits = mask[b];
            const int f = b * 32 + lane;
            if (f < n_frames) r = rec[f];
        }
    };

    acc = 0.0;
    hits = 0;
    // Records stream from HBM once per pass (1.6 MB per channel, no reuse): loads are issued kDepth blocks ahead of
    // their use to cover DRAM latency; ring indices are compile-time.
    constexpr int kDepth = 4;
    double2 ring_r[kDepth];
    uint32_t ring_ok[kDepth];
#pragma unroll
    for (int k = 0; k < kDepth; k++) fetch(k, ring_r[k], ring_ok[k]);
    uint32_t cur_bits[NB], next_bits[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) cur_bits[k] = next_bits[k] = 0;
    if (n_blocks > 0) stage(0, ring_r[0], ring_ok[0], cur_bits);
    fetch(kDepth, ring_r[0], ring_ok[0]);
    __syncwarp();

    for (int b0 = 0; b0 < n_blocks; b0 += kDepth) {
#pragma unroll
        for (int k = 0; k < kDepth; k++) {
            const int b = b0 + k;
            if (b >= n_blocks) break;
            // (1) classify and queue the NEXT block (independent of the chain below), refill its ring slot
            const int slot_next = (k + 1) % kDepth;
            if (b + 1 < n_blocks) stage(b + 1, ring_r[slot_next], ring_ok[slot_next], next_bits);
            fetch(b + 1 + kDepth, ring_r[slot_next], ring_ok[slot_next]);

            // (2) ordered accumulation of block b (:382-386 / :67-72): lane (bucket, comp) adds its bucket's records
            // in record order - a pure DADD chain over exactly those records
            uint32_t mine = 0;
#pragma unroll
            for (int kk = 0; kk < NB; kk++) mine = my_bucket == kk ? cur_bits[kk] : mine;
            const int n_mine = __popc(mine);
            const double *col = reinterpret_cast<const double *>(queues[b & 1].q[my_bucket & 7]) + my_comp;
            // eight queue entries per step: the loads are issued together (entries past the end read as -0.0, the
            // additive identity), then a pure DADD chain
            const int n_max = __reduce_max_sync(0xFFFFFFFFu, n_mine);
            for (int j0 = 0; j0 < n_max; j0 += 8) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = (j0 + j < n_mine) ? col[2 * (j0 + j)] : -0.0;
#pragma unroll
                for (int j = 0; j < 8; j++) acc += v[j];
            }
            hits += n_mine;
#pragma unroll
            for (int kk = 0; kk < NB; kk++) cur_bits[kk] = next_bits[kk];
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(kRefineWarps * 32)
gc_coef_refine_kernel(GcChannelTable tab, const double2 *__restrict__ records, const uint32_t *__restrict__ accept_mask,
                      int16_t *__restrict__ coefs_out)
{
    __shared__ RefineQueues queues[kRefineWarps][2];
    __shared__ double cent[kRefineWarps][8][3];  // centroids (1, c1, c2)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kRefineWarps + warp;
    if (ch >= tab.n_channels) return;  // whole warp leaves together

    double (*best)[3] = cent[warp];
    const int n_frames = div_round_up(tab.n_samples[ch], kGcFrameSamples);
    const int n_blocks = (n_frames + 31) >> 5;
    const double2 *rec = records + tab.rec_off[ch];
    const uint32_t *mask = accept_mask + (tab.rec_off[ch] >> 5);

    if (lane < 8) { best[lane][0] = 1.0; best[lane][1] = 0.0; best[lane][2] = 0.0; }
    __syncwarp();

    // pass 0 is the plain ordered mean (:63-76); passes 1..6 are the FilterRecords rounds (:79-91): split, then two
    // rounds of reassign + average with 2, 4, 8 centroids.
    int count = 1;
    for (int pass = 0; pass < 7; pass++) {
        if (pass == 1 || pass == 3 || pass == 5) {
            // split (:82-89): new centroid = (0.01 * (0,-1,0)) + old
            if (lane < count) {
                best[count + lane][0] = (0.01 * 0.0) + best[lane][0];
                best[count + lane][1] = (0.01 * -1.0) + best[lane][1];
                best[count + lane][2] = (0.01 * 0.0) + best[lane][2];
            }
            count *= 2;
            __syncwarp();
        }
        double acc;
        int hits;
        if (pass == 0) gc_refine_pass<0>(lane, n_frames, n_blocks, rec, mask, queues[warp], best, acc, hits);
        else if (count == 2) gc_refine_pass<2>(lane, n_frames, n_blocks, rec, mask, queues[warp], best, acc, hits);
        else if (count == 4) gc_refine_pass<4>(lane, n_frames, n_blocks, rec, mask, queues[warp], best, acc, hits);
        else gc_refine_pass<8>(lane, n_frames, n_blocks, rec, mask, queues[warp], best, acc, hits);

        // divide (:73-74 / :388-391) and rebuild the centroids (:76 / :393-394)
        double mean;
        if (pass == 0) mean = acc / (double)hits;  // 0/0 = NaN for a silent channel, as in the reference (A.19)
        else mean = hits > 0 ? acc / (double)hits : acc;
        const double m1 = __shfl_sync(0xFFFFFFFFu, mean, (lane & 7) * 2);
        const double m2 = __shfl_sync(0xFFFFFFFFu, mean, (lane & 7) * 2 + 1);
        const int h = __shfl_sync(0xFFFFFFFFu, hits, (lane & 7) * 2);
        __syncwarp();
        if (lane < count) {
            const double m0 = (pass == 0) ? 1.0 : (h > 0 ? 1.0 : 0.0);  // sum of h ones divided by h
            double o1, o2;
            gc_centroid_from_mean(m0, m1, m2, o1, o2);
            best[lane][0] = 1.0;
            best[lane][1] = o1;
            best[lane][2] = o2;
        }
        __syncwarp();
    }

    if (lane < 16) coefs_out[(int64_t)ch * 16 + lane] = gc_quantise_coef(best[lane >> 1][1 + (lane & 1)]);
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
void launch_gc_coef_frames(const int16_t *pcm, const GcChannelTable &tab, double2 *records, uint32_t *mask,
                           int max_frames, int frame_begin, int frame_end, cudaStream_t stream)
{
    if (tab.n_channels <= 0 || max_frames <= 0) return;
    int hi = frame_end < max_frames ? frame_end : max_frames;
    if (hi <= frame_begin) return;
    int tiles = (hi - frame_begin + kP1Threads - 1) / kP1Threads;
    if (tiles > 65535) tiles = 65535;
    dim3 grid((unsigned)tab.n_channels, (unsigned)tiles);
    gc_coef_frames_kernel<<<grid, kP1Threads, 0, stream>>>(pcm, tab, records, mask, frame_begin, frame_end);
}

void launch_gc_coef_refine(const GcChannelTable &tab, const double2 *records, const uint32_t *mask,
                           int16_t *coefs_out, cudaStream_t stream)
{
    if (tab.n_channels <= 0) return;
    int blocks = (tab.n_channels + kRefineWarps - 1) / kRefineWarps;
    gc_coef_refine_kernel<<<blocks, kRefineWarps * 32, 0, stream>>>(tab, records, mask, coefs_out);
}

}  // namespace vgb
