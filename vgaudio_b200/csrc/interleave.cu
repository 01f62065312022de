// interleave.cu — block (de)interleave of channel payloads on sm_100a.
//
// Replaces InterleaveExtensions.Interleave / DeInterleave (Utilities/Interleave.cs:9-166), the byte shuffles the
// container writers / readers run right after / before the codec path: DspWriter / BrstmWriter
// interleave the channels' ADPCM in 0x2000-byte blocks, AdxWriter in frame_size-byte blocks (AdxWriter.cs:131), the WAV
// reader de-interleaves 2-byte samples (WaveReader.cs:47).  Pure data movement: every byte is read once and written
// once, so the kernel is HBM-bound by construction; it moves the widest vector (16/8/4/2/1 bytes) that divides the
// block size, the channel sizes and the strides, one output element per thread-iteration, fully coalesced on the
// interleaved side and in block-sized runs on the planar side.
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

// planar [count][in_size] -> interleaved [out_size * count]; one output element of width sizeof(V) per iteration
template <typename V>
__global__ void __launch_bounds__(256)
interleave_kernel(const uint8_t *__restrict__ in, int64_t in_channel_stride, int64_t in_item_stride, uint8_t *__restrict__ out,
                  int64_t out_item_stride, int n_items, Shape sh)
{
    const int64_t per_item = sh.out_size * sh.count / (int64_t)sizeof(V);  // elements of one item's output
    const int64_t total = per_item * n_items;
    const int64_t full_block = sh.interleave * sh.count;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t item = e / per_item;
        const int64_t o = (e - item * per_item) * (int64_t)sizeof(V);  // byte offset inside the item's output
        const int64_t b = o / full_block;                                // block index (all earlier blocks are full)
        const int64_t r = o - b * full_block;
        const int64_t cur_out = b == sh.out_blocks - 1 ? sh.last_out : sh.interleave;
        const int64_t cur_in = b == sh.in_blocks - 1 ? sh.last_in : sh.interleave;
        const int64_t ch = r / cur_out, off = r - ch * cur_out;
        V v{};
        if (b < sh.blocks_to_copy && off < (cur_in < cur_out ? cur_in : cur_out))
            v = *reinterpret_cast<const V *>(in + item * in_item_stride + ch * in_channel_stride + sh.interleave * b + off);
        *reinterpret_cast<V *>(out + item * out_item_stride + o) = v;
    }
}

// interleaved [in_size * count] -> planar [count][out_size]; one output element per iteration
template <typename V>
__global__ void __launch_bounds__(256)
deinterleave_kernel(const uint8_t *__restrict__ in, int64_t in_item_stride, uint8_t *__restrict__ out, int64_t out_channel_stride,
                    int64_t out_item_stride, int n_items, Shape sh)
{
    const int64_t per_channel = sh.out_size / (int64_t)sizeof(V);
    const int64_t per_item = per_channel * sh.count;
    const int64_t total = per_item * n_items;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t item = e / per_item;
        const int64_t q = e - item * per_item;
        // consecutive threads walk one block of one channel, then the same block of the next channel: reads stay
        // sequential in the interleaved stream, writes are block-sized runs
        const int64_t block_elems = sh.interleave / (int64_t)sizeof(V);
        const int64_t group = block_elems * sh.count;  // elements of one full block over all channels
        const int64_t b = q / group;
        int64_t ch, off;
        if (b < sh.out_blocks - 1) {
            const int64_t r = q - b * group;
            ch = r / block_elems;
            off = (r - ch * block_elems) * (int64_t)sizeof(V);
        } else {  // last output block: last_out bytes per channel
            const int64_t last_elems = sh.last_out / (int64_t)sizeof(V);
            const int64_t r = q - (sh.out_blocks - 1) * group;
            ch = r / last_elems;
            off = (r - ch * last_elems) * (int64_t)sizeof(V);
        }
        const int64_t bb = b < sh.out_blocks - 1 ? b : sh.out_blocks - 1;
        const int64_t cur_out = bb == sh.out_blocks - 1 ? sh.last_out : sh.interleave;
        const int64_t cur_in = bb == sh.in_blocks - 1 ? sh.last_in : sh.interleave;
        V v{};
        if (bb < sh.blocks_to_copy && off < (cur_in < cur_out ? cur_in : cur_out))
            v = *reinterpret_cast<const V *>(in + item * in_item_stride + sh.interleave * bb * sh.count + cur_in * ch + off);
        *reinterpret_cast<V *>(out + item * out_item_stride + ch * out_channel_stride + sh.interleave * bb + off) = v;
    }
}

struct alignas(16) Vec16 { uint32_t x, y, z, w; };

inline int vector_width(std::initializer_list<int64_t> values)
{
    int w = 16;
    for (int64_t v : values)
        while (w > 1 && (v % w) != 0) w >>= 1;
    return w;
}

inline int grid_for(int64_t elements)
{
    const int64_t want = (elements + 255) / 256;
    const int64_t cap = 148 * 16;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

cudaError_t launch_interleave(const void *in, int64_t in_channel_stride, int64_t in_item_stride, void *out, int64_t out_item_stride,
                              int n_items, int count, int64_t in_size, int64_t interleave, int64_t out_size, cudaStream_t stream)
{
    if (n_items <= 0 || count <= 0 || out_size <= 0) return cudaSuccess;
    const Shape sh = make_shape(count, in_size, interleave, out_size);
    const int w = vector_width({interleave, in_size, out_size, sh.last_in, sh.last_out, in_channel_stride, in_item_stride, out_item_stride,
                                (int64_t)reinterpret_cast<uintptr_t>(in), (int64_t)reinterpret_cast<uintptr_t>(out)});
    const int64_t elems = out_size * count / w * n_items;
    const int grid = grid_for(elems);
    const uint8_t *i8 = static_cast<const uint8_t *>(in);
    uint8_t *o8 = static_cast<uint8_t *>(out);
    switch (w) {
    case 16: interleave_kernel<Vec16><<<grid, 256, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, n_items, sh); break;
    case 8: interleave_kernel<uint64_t><<<grid, 256, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, n_items, sh); break;
    case 4: interleave_kernel<uint32_t><<<grid, 256, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, n_items, sh); break;
    case 2: interleave_kernel<uint16_t><<<grid, 256, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, n_items, sh); break;
    default: interleave_kernel<uint8_t><<<grid, 256, 0, stream>>>(i8, in_channel_stride, in_item_stride, o8, out_item_stride, n_items, sh); break;
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
    const int64_t elems = out_size / w * count * n_items;
    const int grid = grid_for(elems);
    const uint8_t *i8 = static_cast<const uint8_t *>(in);
    uint8_t *o8 = static_cast<uint8_t *>(out);
    switch (w) {
    case 16: deinterleave_kernel<Vec16><<<grid, 256, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, n_items, sh); break;
    case 8: deinterleave_kernel<uint64_t><<<grid, 256, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, n_items, sh); break;
    case 4: deinterleave_kernel<uint32_t><<<grid, 256, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, n_items, sh); break;
    case 2: deinterleave_kernel<uint16_t><<<grid, 256, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, n_items, sh); break;
    default: deinterleave_kernel<uint8_t><<<grid, 256, 0, stream>>>(i8, in_item_stride, o8, out_channel_stride, out_item_stride, n_items, sh); break;
    }
    return cudaGetLastError();
}

}  // namespace vgb
