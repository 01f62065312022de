// interleave.cu — block (de)interleave of channel payloads on sm_100a.
//
// Replaces InterleaveExtensions.Interleave / DeInterleave (Utilities/Interleave.cs:9-166), the byte shuffles the
// container writers / readers run right after / before the codec path: DspWriter / BrstmWriter
// interleave the channels' ADPCM in 0x2000-byte blocks, AdxWriter in frame_size-byte blocks (AdxWriter.cs:131), the WAV
// reader de-interleaves 2-byte samples (WaveReader.cs:47).  Pure data movement: every byte is read once and written
// once, so the kernel is HBM-bound by construction; it moves the widest vector (16/8/4/2/1 bytes) that divides the
// block size, the channel sizes and the strides, four independent elements per thread (loads first, stores after),
// fully coalesced on the interleaved side and in block-sized runs on the planar side; 32-bit index arithmetic per item.
//
// Semantics (identical for both directions, :17-23): blocks of `interleave` bytes per channel; the LAST block of the
// input / output may be shorter (size - (blocks-1)*interleave); only min(inBlocks, outBlocks) blocks are copied and
// in a block only min(currentInputBlock, currentOutputBlock) bytes per channel; the rest of the output stays zero.
#include "common.cuh"
#include "kernels.h"

namespace vgb {

namespace {

struct Shape {
    int64_t in_size, out_size, interleave;  // bytes per channel / per block
    int64_t in_blocks, out_blocks, last_in, last_out, blocks_to_copy;
    int count;
};

__host__ __device__ inline Shape make_shape(int count, int64_t in_size, int64_t interleave, int64_t out_size)
{
    Shape s;
    s.count = count; s.in_size = in_size; s.out_size = out_size; s.interleave = interleave;
    s.in_blocks = (in_size + interleave - 1) / interleave;
    s.out_blocks = (out_size + interleave - 1) / interleave;
    s.last_in = in_size - (s.in_blocks - 1) * interleave;
    s.last_out = out_size - (s.out_blocks - 1) * interleave;
    s.blocks_to_copy = s.in_blocks < s.out_blocks ? s.in_blocks : s.out_blocks;
    return s;
}

constexpr int kIlvThreads = 256, kIlvIter = 4;  // elements per thread: independent loads in flight

// planar [count][in_size] -> interleaved [out_size * count].  A CTA works on kIlvThreads * kIlvIter consecutive output
// elements of ONE item, so all index arithmetic is 32-bit (an item's payload is < 2 GiB) and the item offset is added
// once; one output element of width sizeof(V) per thread and iteration, consecutive threads on consecutive elements.
template <typename V>
__global__ void __launch_bounds__(kIlvThreads)
interleave_kernel(const uint8_t *__restrict__ in, int64_t in_channel_stride, int64_t in_item_stride, uint8_t *__restrict__ out,
                  int64_t out_item_stride, uint32_t blocks_per_item, Shape sh)
{
    const uint32_t item = blockIdx.x / blocks_per_item, bx = blockIdx.x - item * blocks_per_item;
    const uint32_t per_item = (uint32_t)(sh.out_size * sh.count / (int64_t)sizeof(V));
    const uint32_t full_block = (uint32_t)(sh.interleave * sh.count), ilv = (uint32_t)sh.interleave;
    const uint32_t out_blocks = (uint32_t)sh.out_blocks, in_blocks = (uint32_t)sh.in_blocks, to_copy = (uint32_t)sh.blocks_to_copy;
    const uint32_t last_out = (uint32_t)sh.last_out, last_in = (uint32_t)sh.last_in;
    const uint8_t *src = in + (int64_t)item * in_item_stride;
    uint8_t *dst = out + (int64_t)item * out_item_stride;
    V v[kIlvIter];
    uint32_t o[kIlvIter];
#pragma unroll
    for (int k = 0; k < kIlvIter; k++) {
        const uint32_t e = (bx * kIlvIter + k) * kIlvThreads + threadIdx.x;
        o[k] = e * (uint32_t)sizeof(V);                         // byte offset inside the item's output
        v[k] = V{};
        if (e < per_item) {
            const uint32_t b = o[k] / full_block;               // block index (all earlier blocks are full)
            const uint32_t r = o[k] - b * full_block;
            const uint32_t cur_out = b == out_blocks - 1 ? last_out : ilv;
            const uint32_t cur_in = b == in_blocks - 1 ? last_in : ilv;
            const uint32_t ch = r / cur_out, off = r - ch * cur_out;
            if (b < to_copy && off < (cur_in < cur_out ? cur_in : cur_out))
                v[k] = *reinterpret_cast<const V *>(src + (int64_t)ch * in_channel_stride + (ilv * b + off));
        }
    }
#pragma unroll
    for (int k = 0; k < kIlvIter; k++)
        if (o[k] / (uint32_t)sizeof(V) < per_item) *reinterpret_cast<V *>(dst + o[k]) = v[k];
}

// interleaved [in_size * count] -> planar [count][out_size]; elements are walked in the order of the interleaved stream
// (one block of one channel, then the same block of the next channel): reads stay sequential, writes are block runs
template <typename V>
__global__ void __launch_bounds__(kIlvThreads)
deinterleave_kernel(const uint8_t *__restrict__ in, int64_t in_item_stride, uint8_t *__restrict__ out, int64_t out_channel_stride,
                    int64_t out_item_stride, uint32_t blocks_per_item, Shape sh)
{
    const uint32_t item = blockIdx.x / blocks_per_item, bx = blockIdx.x - item * blocks_per_item;
    const uint32_t per_item = (uint32_t)(sh.out_size / (int64_t)sizeof(V)) * (uint32_t)sh.count;
    const uint32_t ilv = (uint32_t)sh.interleave, count = (uint32_t)sh.count;
    const uint32_t block_elems = ilv / (uint32_t)sizeof(V), group = block_elems * count;
    const uint32_t out_blocks = (uint32_t)sh.out_blocks, in_blocks = (uint32_t)sh.in_blocks, to_copy = (uint32_t)sh.blocks_to_copy;
    const uint32_t last_out = (uint32_t)sh.last_out, last_in = (uint32_t)sh.last_in;
    const uint32_t last_elems = last_out / (uint32_t)sizeof(V);
    const uint8_t *src = in + (int64_t)item * in_item_stride;
    uint8_t *dst = out + (int64_t)item * out_item_stride;
    V v[kIlvIter];
    int64_t where[kIlvIter];
#pragma unroll
    for (int k = 0; k < kIlvIter; k++) {
        const uint32_t q = (bx * kIlvIter + k) * kIlvThreads + threadIdx.x;
        v[k] = V{};
        where[k] = -1;
        if (q < per_item) {
            uint32_t b = q / group, ch, off;
            if (b < out_blocks - 1) {
                const uint32_t r = q - b * group;
                ch = r / block_elems;
                off = (r - ch * block_elems) * (uint32_t)sizeof(V);
            } else {  // last output block: last_out bytes per channel
                b = out_blocks - 1;
                const uint32_t r = q - b * group;
                ch = r / last_elems;
                off = (r - ch * last_elems) * (uint32_t)sizeof(V);
            }
            const uint32_t cur_out = b == out_blocks - 1 ? last_out : ilv;
            const uint32_t cur_in = b == in_blocks - 1 ? last_in : ilv;
            if (b < to_copy && off < (cur_in < cur_out ? cur_in : cur_out))
                v[k] = *reinterpret_cast<const V *>(src + ((int64_t)ilv * b * count + cur_in * ch + off));
            where[k] = (int64_t)ch * out_channel_stride + (ilv * b + off);
        }
    }
#pragma unroll
    for (int k = 0; k < kIlvIter; k++)
        if (where[k] >= 0) *reinterpret_cast<V *>(dst + where[k]) = v[k];
}

struct alignas(16) Vec16 { uint32_t x, y, z, w; };

inline int vector_width(std::initializer_list<int64_t> values)
{
    int w = 16;
    for (int64_t v : values)
        while (w > 1 && (v % w) != 0) w >>= 1;
    return w;
}

inline uint32_t blocks_per_item_for(int64_t elements_per_item)
{
    const int64_t per_block = (int64_t)kIlvThreads * kIlvIter;
    return (uint32_t)((elements_per_item + per_block - 1) / per_block);
}

}  // namespace

cudaError_t launch_interleave(const void *in, int64_t in_channel_stride, int64_t in_item_stride, void *out, int64_t out_item_stride,
                              int n_items, int count, int64_t in_size, int64_t interleave, int64_t out_size, cudaStream_t stream)
{
    if (n_items <= 0 || count <= 0 || out_size <= 0) return cudaSuccess;
    const Shape sh = make_shape(count, in_size, interleave, out_size);
    const int w = vector_width({interleave, in_size, out_size, sh.last_in, sh.last_out, in_channel_stride, in_item_stride, out_item_stride,
                                (int64_t)reinterpret_cast<uintptr_t>(in), (int64_t)reinterpret_cast<uintptr_t>(out)});
    if (out_size * count >= (int64_t)1 << 31 || in_size >= (int64_t)1 << 31) return cudaErrorInvalidValue;  // 32-bit index arithmetic per item
    const uint32_t bpi = blocks_per_item_for(out_size * count / w);
    const int64_t grid64 = (int64_t)bpi * n_items;
    if (grid64 >= (int64_t)1 << 31) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)grid64;
    const uint8_t *i8 = static_cast<const uint8_t *>(in);
    uint8_t *o8 = static_cast<uint8_t *>(out);
    switch (w) {
    case 16: interleave_kernel<Vec16><<<grid, kIlvThreads, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, bpi, sh); break;
    case 8: interleave_kernel<uint64_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, bpi, sh); break;
    case 4: interleave_kernel<uint32_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, bpi, sh); break;
    case 2: interleave_kernel<uint16_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, bpi, sh); break;
    default: interleave_kernel<uint8_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, bpi, sh); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_deinterleave(const void *in, int64_t in_item_stride, void *out, int64_t out_channel_stride, int64_t out_item_stride,
                                int n_items, int count, int64_t in_size, int64_t interleave, int64_t out_size, cudaStream_t stream)
{
    if (n_items <= 0 || count <= 0 || out_size <= 0) return cudaSuccess;
    const Shape sh = make_shape(count, in_size, interleave, out_size);
    const int w = vector_width({interleave, in_size, out_size, sh.last_in, sh.last_out, out_channel_stride, in_item_stride, out_item_stride,
                                (int64_t)reinterpret_cast<uintptr_t>(in), (int64_t)reinterpret_cast<uintptr_t>(out)});
    if (out_size * count >= (int64_t)1 << 31 || in_size * count >= (int64_t)1 << 31) return cudaErrorInvalidValue;
    const uint32_t bpi = blocks_per_item_for(out_size / w * count);
    const int64_t grid64 = (int64_t)bpi * n_items;
    if (grid64 >= (int64_t)1 << 31) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)grid64;
    const uint8_t *i8 = static_cast<const uint8_t *>(in);
    uint8_t *o8 = static_cast<uint8_t *>(out);
    switch (w) {
    case 16: deinterleave_kernel<Vec16><<<grid, kIlvThreads, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, bpi, sh); break;
    case 8: deinterleave_kernel<uint64_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, bpi, sh); break;
    case 4: deinterleave_kernel<uint32_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, bpi, sh); break;
    case 2: deinterleave_kernel<uint16_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, bpi, sh); break;
    default: deinterleave_kernel<uint8_t><<<grid, kIlvThreads, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, bpi, sh); break;
    }
    return cudaGetLastError();
}

}  // namespace vgb
