/* adx_quantiser_check.c — proves the integer form of the CRI ADX quantiser used by adx_encode_kernel.
 *
 * Reference (CriAdxCodec.cs:126,167-171):   scaled = Clamp16((int)(raw * gain)),  gain = 32767.0 / maxDistance (double)
 *                                           q = Clamp4((scaled + 2340 * Sign(scaled)) / 4681)          ("/" truncates)
 * Integer form:                             a = min(|raw|, 65535)
 *                                           q = sign(raw) * #{ k in 1..7 : a * 32767 >= T_k * maxDistance },  T_k = 4681 k - 2340
 * (maxDistance == 0: gain is 0 and q = 0;  a * 32767 >= 2^31 * maxDistance: the (int) cast overflows, q = -7.)
 *
 * Why it can hold: q counts the thresholds T_k = 2341, 7022, ... 30427 that |scaled| reaches; (int) truncation and
 * Clamp16 do not move a value across an integer threshold, so |scaled| >= T_k  <=>  fl(a * fl(32767/md)) >= T_k.  The two
 * roundings move the product by < 1e-11, while a*32767/md differs from T_k by >= 1/md unless they are EQUAL, which needs
 * 4681 | md (T_k is coprime to 31 and 151).  For md = 4681 j, j in {1,2,4,7} the gain 7/j is exact; j in {3,5,6} are the
 * only maxDistance values where the double chain could disagree - this program finds out (exhaustively over every
 * maxDistance and, by monotonicity in a, around every threshold), and prints the exceptions the kernel must treat.
 *
 * Build/run: gcc -O2 -ffp-contract=off -o /tmp/adx_q tools/adx_quantiser_check.c && /tmp/adx_q
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static int clamp16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
static int clamp4(int v) { return v > 7 ? 7 : (v < -8 ? -8 : v); }
static int sign(int v) { return (v > 0) - (v < 0); }

static int q_reference(int raw, int md)
{
    const double gain = md == 0 ? 0.0 : 32767.0 / (double)md;
    volatile double prod = (double)raw * gain; /* one rounding, no contraction */
    /* (int)double on x64 is cvttsd2si: out-of-range gives INT_MIN; |raw| <= 2^18 and gain <= 32767 keep it in range */
    const int scaled = clamp16(prod >= 2147483648.0 || prod < -2147483648.0 ? INT32_MIN : (int)prod);
    return clamp4((scaled + 2340 * sign(scaled)) / 4681);
}

static int q_integer(int raw, int md)
{
    if (md == 0) return 0;
    uint32_t a = (uint32_t)abs(raw);
    /* (int)(raw * gain) leaves int32 when a * 32767 / md >= 2^31: x64 cvttsd2si then yields INT_MIN, Clamp16 -32768, q -7
     * whatever the sign of raw.  Needs a >= 65538 * md, i.e. md <= 3 for |raw| < 2^18. */
    if ((uint64_t)a * 32767u >= ((uint64_t)md << 31)) return -7;
    if (a > 65535u) a = 65535u;
    const uint32_t v = a * 32767u;
    int k = 0;
    for (int t = 1; t <= 7; t++) k += v >= (uint32_t)(4681 * t - 2340) * (uint32_t)md;
    return raw < 0 ? -k : k;
}

int main(void)
{
    long long checked = 0, bad = 0;
    for (int md = 0; md <= 32768; md++) {
        /* dense sweep for small md, threshold neighbourhoods (+-3) and extremes for every md */
        if (md <= 64)
            for (int raw = -140000; raw <= 140000; raw++) {
                checked++;
                if (q_reference(raw, md) != q_integer(raw, md)) { if (bad++ < 20) printf("MISMATCH md=%d raw=%d ref=%d int=%d\n", md, raw, q_reference(raw, md), q_integer(raw, md)); }
            }
        for (int t = 1; t <= 7; t++) {
            const long long T = 4681 * t - 2340;
            const long long a0 = md ? (T * md + 32766) / 32767 : 0;
            for (long long a = a0 - 3; a <= a0 + 3; a++) {
                if (a < 0) continue;
                for (int s = -1; s <= 1; s += 2) {
                    const int raw = (int)(s * a);
                    checked++;
                    if (q_reference(raw, md) != q_integer(raw, md)) { if (bad++ < 20) printf("MISMATCH md=%d raw=%d ref=%d int=%d\n", md, raw, q_reference(raw, md), q_integer(raw, md)); }
                }
            }
        }
        const int extremes[] = {0, 1, -1, 30427, 30428, 32767, -32768, 65535, -65535, 65536, 100000, -100000, 140000, -140000, 262143, -262143};
        for (unsigned i = 0; i < sizeof extremes / sizeof *extremes; i++) {
            checked++;
            if (q_reference(extremes[i], md) != q_integer(extremes[i], md)) { if (bad++ < 20) printf("MISMATCH md=%d raw=%d ref=%d int=%d\n", md, extremes[i], q_reference(extremes[i], md), q_integer(extremes[i], md)); }
        }
    }
    printf("checked %lld (md 0..32768), mismatches %lld\n", checked, bad);
    return bad != 0;
}
