/* interleave.c — CPU restatement of InterleaveExtensions.Interleave / DeInterleave for bytes
 * (src/VGAudio/Utilities/Interleave.cs:9-41, :81-117).  TEST INFRASTRUCTURE (see vgoracle.h). */
#include <string.h>

#include "vgoracle.h"

static int div_up(int a, int b) { return (a + b - 1) / b; }

/* T[] Interleave<T>(this T[][] inputs, int interleaveSize, int outputSize = -1)  (:9-41).  output: out_size * count bytes */
int vgo_interleave(const uint8_t *const *inputs, int count, int in_size, int interleave_size, int out_size, uint8_t *output)
{
    if (out_size == -1) out_size = in_size;
    int in_blocks = div_up(in_size, interleave_size), out_blocks = div_up(out_size, interleave_size);
    int last_in = in_size - (in_blocks - 1) * interleave_size, last_out = out_size - (out_blocks - 1) * interleave_size;
    int blocks = in_blocks < out_blocks ? in_blocks : out_blocks;
    memset(output, 0, (size_t)out_size * (size_t)count);
    for (int b = 0; b < blocks; b++) {
        int cur_in = b == in_blocks - 1 ? last_in : interleave_size;
        int cur_out = b == out_blocks - 1 ? last_out : interleave_size;
        int n = cur_in < cur_out ? cur_in : cur_out;
        for (int i = 0; i < count; i++)
            memcpy(output + (size_t)interleave_size * b * count + (size_t)cur_out * i, inputs[i] + (size_t)interleave_size * b, (size_t)n);
    }
    return 0;
}

/* T[][] DeInterleave<T>(this T[] input, int interleaveSize, int outputCount, int outputSize = -1)  (:81-117) */
int vgo_deinterleave(const uint8_t *input, int length, int interleave_size, int count, int out_size, uint8_t *const *outputs)
{
    if (length % count != 0) return -1; /* ArgumentOutOfRangeException */
    int in_size = length / count;
    if (out_size == -1) out_size = in_size;
    int in_blocks = div_up(in_size, interleave_size), out_blocks = div_up(out_size, interleave_size);
    int last_in = in_size - (in_blocks - 1) * interleave_size, last_out = out_size - (out_blocks - 1) * interleave_size;
    int blocks = in_blocks < out_blocks ? in_blocks : out_blocks;
    for (int o = 0; o < count; o++) memset(outputs[o], 0, (size_t)out_size);
    for (int b = 0; b < blocks; b++) {
        int cur_in = b == in_blocks - 1 ? last_in : interleave_size;
        int cur_out = b == out_blocks - 1 ? last_out : interleave_size;
        int n = cur_in < cur_out ? cur_in : cur_out;
        for (int o = 0; o < count; o++)
            memcpy(outputs[o] + (size_t)interleave_size * b, input + (size_t)interleave_size * b * count + (size_t)cur_in * o, (size_t)n);
    }
    return 0;
}
