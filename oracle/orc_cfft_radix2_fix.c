/*
 * oracle/orc_cfft_radix2_fix.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restatement of the reference's deprecated fixed-point radix-2 transforms, generic-C branch:
 *   arm_cfft_radix2_q31   Source/TransformFunctions/arm_cfft_radix2_q31.c:62-81  (forward butterflies :87-201, inverse :204-318)
 *   arm_cfft_radix2_q15   Source/TransformFunctions/arm_cfft_radix2_q15.c:62-78  (forward butterflies :275-386, inverse :577-681)
 *   arm_bitreversal_q31 / _q15   Source/TransformFunctions/arm_bitreversal.c:121-175, 196-254
 * Decimation in frequency, log2(N) radix-2 stages in place, then the bit reversal -- which both functions apply
 * whatever the instance's bitReverseFlag says (:80, :77) -- leaves the spectrum in natural order.
 *   stage 1        inputs >> 1; the sum is halved again; the difference is rotated by W_N^i
 *   middle stages  sum >> 1; the difference is rotated by W^(j * 2^stage)
 *   last stage     plain sum and difference
 * q31 rotates with the rounding multiply-accumulates of Include/dsp/none.h:185-194; q15 with truncating >> 16 products
 * stored to int16 (the difference xt itself is an int16 variable: it wraps).
 * Twiddles: the instance points at the 4096-point table and reads it with twidCoefModifier = 4096 / fftLen
 * (arm_cfft_radix2_init_q31.c:79-170): entry k * modifier of that table is entry k of the fftLen-point table.
 */
#include "orc_fft.h"

static int32_t rhi(int32_t x, int32_t y) { return (int32_t)(((int64_t)x * y + 0x80000000LL) >> 32); }
static int32_t rhi_acc(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) + (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
static int32_t rhi_sub(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) - (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
static int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

static uint32_t bitrev(uint32_t k, uint32_t n)
{
    uint32_t r = 0;
    for (uint32_t m = n >> 1; m; m >>= 1, k >>= 1) r = (r << 1) | (k & 1u);
    return r;
}

void orc_cfft_radix2_q31(uint32_t N, int32_t *p, int ifftFlag)
{
    const int32_t *tw = orc_twiddle_q31(N);
    if (!tw) return;
    uint32_t n2 = N, mod = 1;
    for (uint32_t stage = 0; n2 > 1; stage++, mod <<= 1) {
        const uint32_t n1 = n2;
        n2 >>= 1;
        const int first = (stage == 0), last = (n2 == 1);
        for (uint32_t j = 0; j < n2; j++) {
            const int32_t co = tw[2 * j * mod], si = tw[2 * j * mod + 1];
            for (uint32_t i = j; i < N; i += n1) {
                const uint32_t l = i + n2;
                int32_t xt, yt;
                if (first) {                                                  /* :116-122 */
                    xt = wsub(p[2 * i] >> 1, p[2 * l] >> 1);
                    p[2 * i] = wadd(p[2 * i] >> 1, p[2 * l] >> 1) >> 1;
                    yt = wsub(p[2 * i + 1] >> 1, p[2 * l + 1] >> 1);
                    p[2 * i + 1] = wadd(p[2 * l + 1] >> 1, p[2 * i + 1] >> 1) >> 1;
                } else if (!last) {                                           /* :155-159 */
                    xt = wsub(p[2 * i], p[2 * l]);
                    p[2 * i] = wadd(p[2 * i], p[2 * l]) >> 1;
                    yt = wsub(p[2 * i + 1], p[2 * l + 1]);
                    p[2 * i + 1] = wadd(p[2 * l + 1], p[2 * i + 1]) >> 1;
                } else {                                                      /* :185-197 */
                    xt = wsub(p[2 * i], p[2 * l]);
                    p[2 * i] = wadd(p[2 * i], p[2 * l]);
                    yt = wsub(p[2 * i + 1], p[2 * l + 1]);
                    p[2 * i + 1] = wadd(p[2 * l + 1], p[2 * i + 1]);
                    p[2 * l] = xt;
                    p[2 * l + 1] = yt;
                    continue;
                }
                int32_t p0 = rhi(xt, co), p1 = rhi(yt, co);
                if (!ifftFlag) { p0 = rhi_acc(p0, yt, si); p1 = rhi_sub(p1, xt, si); }      /* :124-127 */
                else           { p0 = rhi_sub(p0, yt, si); p1 = rhi_acc(p1, xt, si); }      /* :241-244 */
                p[2 * l] = p0;
                p[2 * l + 1] = p1;
            }
        }
    }
    for (uint32_t k = 0; k < N; k++) {                                        /* arm_bitreversal_q31 */
        const uint32_t r = bitrev(k, N);
        if (k < r) {
            int32_t t = p[2 * k]; p[2 * k] = p[2 * r]; p[2 * r] = t;
            t = p[2 * k + 1]; p[2 * k + 1] = p[2 * r + 1]; p[2 * r + 1] = t;
        }
    }
}

void orc_cfft_radix2_q15(uint32_t N, int16_t *p, int ifftFlag)
{
    const int16_t *tw = orc_twiddle_q15(N);
    if (!tw) return;
    uint32_t n2 = N, mod = 1;
    for (uint32_t stage = 0; n2 > 1; stage++, mod <<= 1) {
        const uint32_t n1 = n2;
        n2 >>= 1;
        const int first = (stage == 0), last = (n2 == 1);
        for (uint32_t j = 0; j < n2; j++) {
            const int16_t co = tw[2 * j * mod], si = tw[2 * j * mod + 1];
            for (uint32_t i = j; i < N; i += n1) {
                const uint32_t l = i + n2;
                int16_t xt, yt;
                if (first) {                                                  /* :300-306 */
                    xt = (int16_t)((p[2 * i] >> 1) - (p[2 * l] >> 1));
                    p[2 * i] = (int16_t)(((p[2 * i] >> 1) + (p[2 * l] >> 1)) >> 1);
                    yt = (int16_t)((p[2 * i + 1] >> 1) - (p[2 * l + 1] >> 1));
                    p[2 * i + 1] = (int16_t)(((p[2 * l + 1] >> 1) + (p[2 * i + 1] >> 1)) >> 1);
                } else if (!last) {                                           /* :337-341 */
                    xt = (int16_t)(p[2 * i] - p[2 * l]);
                    p[2 * i] = (int16_t)((p[2 * i] + p[2 * l]) >> 1);
                    yt = (int16_t)(p[2 * i + 1] - p[2 * l + 1]);
                    p[2 * i + 1] = (int16_t)((p[2 * l + 1] + p[2 * i + 1]) >> 1);
                } else {                                                      /* :370-379 */
                    xt = (int16_t)(p[2 * i] - p[2 * l]);
                    p[2 * i] = (int16_t)(p[2 * i] + p[2 * l]);
                    yt = (int16_t)(p[2 * i + 1] - p[2 * l + 1]);
                    p[2 * i + 1] = (int16_t)(p[2 * l + 1] + p[2 * i + 1]);
                    p[2 * l] = xt;
                    p[2 * l + 1] = yt;
                    continue;
                }
                const int16_t xc = (int16_t)(((int32_t)xt * co) >> 16), ys = (int16_t)(((int32_t)yt * si) >> 16);
                const int16_t yc = (int16_t)(((int32_t)yt * co) >> 16), xs = (int16_t)(((int32_t)xt * si) >> 16);
                if (!ifftFlag) { p[2 * l] = (int16_t)(xc + ys); p[2 * l + 1] = (int16_t)(yc - xs); }     /* :308-312 */
                else           { p[2 * l] = (int16_t)(xc - ys); p[2 * l + 1] = (int16_t)(yc + xs); }     /* :610-614 */
            }
        }
    }
    for (uint32_t k = 0; k < N; k++) {                                        /* arm_bitreversal_q15 */
        const uint32_t r = bitrev(k, N);
        if (k < r) {
            int16_t t = p[2 * k]; p[2 * k] = p[2 * r]; p[2 * r] = t;
            t = p[2 * k + 1]; p[2 * k + 1] = p[2 * r + 1]; p[2 * r + 1] = t;
        }
    }
}

void orc_cfft_radix2_q31_batch(uint32_t N, int32_t *p, uint64_t nFrames, int ifft)
{
    for (uint64_t f = 0; f < nFrames; f++) orc_cfft_radix2_q31(N, p + 2ull * N * f, ifft);
}
void orc_cfft_radix2_q15_batch(uint32_t N, int16_t *p, uint64_t nFrames, int ifft)
{
    for (uint64_t f = 0; f < nFrames; f++) orc_cfft_radix2_q15(N, p + 2ull * N * f, ifft);
}
