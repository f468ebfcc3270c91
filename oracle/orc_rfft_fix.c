/*
 * oracle/orc_rfft_fix.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restatement of the reference's generic-C (non-DSP, non-Neon, non-MVE) fixed-point real FFT:
 *   arm_rfft_q31   Source/TransformFunctions/arm_rfft_q31.c:145-181  (+ arm_split_rfft_q31 :257-342,
 *                  arm_split_rifft_q31 :406-478, arm_shift_q31 BasicMathFunctions/arm_shift_q31.c)
 *   arm_rfft_q15   Source/TransformFunctions/arm_rfft_q15.c:148-182  (+ arm_split_rfft_q15 :267-408 generic
 *                  branch :352-405, arm_split_rifft_q15 :486-581 generic branch :552-570, arm_shift_q15)
 * and of the generation rule of realCoefA/BQ31, realCoefA/BQ15
 * (Source/CommonTables/arm_common_tables.c:39063-45280: round(x * 2^31 | 2^15), clipped).
 *
 * N = real length 32..8192, L2 = N/2 = length of the inner complex FFT.
 *   forward: pSrc N scalars (destroyed: holds the L2-point CFFT afterwards), pDst 2N scalars =
 *            N complex bins, bins L2+1..N-1 being the conjugate mirror (written explicitly).
 *   inverse: pSrc bins 0..L2 are read (N+2 scalars), pDst N scalars.
 */
#include "orc_fft.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

#define RC_N 4096               /* table generation length n (arm_common_tables.c:39065) */
static int32_t g_a31[2 * RC_N], g_b31[2 * RC_N];
static int16_t g_a15[2 * RC_N], g_b15[2 * RC_N];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static int64_t rclip(double v, int64_t lo, int64_t hi)
{
    v = floor(v + 0.5);
    if (v < (double)lo) return lo;
    if (v > (double)hi) return hi;
    return (int64_t)v;
}

static void build(void)
{
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < RC_N; i++) {
        const double a = 2.0 * pi / (double)(2 * RC_N) * (double)i;
        const double a0 = 0.5 * (1.0 - sin(a)), a1 = 0.5 * (-1.0 * cos(a));
        const double b0 = 0.5 * (1.0 + sin(a)), b1 = 0.5 * (1.0 * cos(a));
        g_a31[2 * i] = (int32_t)rclip(a0 * 2147483648.0, INT32_MIN, INT32_MAX);
        g_a31[2 * i + 1] = (int32_t)rclip(a1 * 2147483648.0, INT32_MIN, INT32_MAX);
        g_b31[2 * i] = (int32_t)rclip(b0 * 2147483648.0, INT32_MIN, INT32_MAX);
        g_b31[2 * i + 1] = (int32_t)rclip(b1 * 2147483648.0, INT32_MIN, INT32_MAX);
        g_a15[2 * i] = (int16_t)rclip(a0 * 32768.0, INT16_MIN, INT16_MAX);
        g_a15[2 * i + 1] = (int16_t)rclip(a1 * 32768.0, INT16_MIN, INT16_MAX);
        g_b15[2 * i] = (int16_t)rclip(b0 * 32768.0, INT16_MIN, INT16_MAX);
        g_b15[2 * i + 1] = (int16_t)rclip(b1 * 32768.0, INT16_MIN, INT16_MAX);
    }
}

const int32_t *orc_real_coef_q31(int b) { pthread_once(&g_once, build); return b ? g_b31 : g_a31; }
const int16_t *orc_real_coef_q15(int b) { pthread_once(&g_once, build); return b ? g_b15 : g_a15; }

static int supported(uint32_t N) { return N >= 32 && N <= 8192 && (N & (N - 1)) == 0; }

/* Include/dsp/none.h:185-194 */
static int32_t mulr(int32_t x, int32_t y) { return (int32_t)(((int64_t)x * y + 0x80000000LL) >> 32); }
static int32_t accr(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) + (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
static int32_t subr(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) - (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
static int32_t wneg(int32_t a) { return (int32_t)(0u - (uint32_t)a); }

void orc_rfft_q31(uint32_t N, int32_t *pSrc, int32_t *pDst, int ifftFlagR, int bitReverseFlagR)
{
    if (!supported(N)) return;
    pthread_once(&g_once, build);
    const uint32_t L2 = N >> 1, mod = 8192u / N;
    const int32_t *A = g_a31, *B = g_b31;
    if (ifftFlagR) {
        /* arm_split_rifft_q31 (:406-478): i = 0..L2-1, S1 = X[i], S2 = X[L2-i] */
        for (uint32_t i = 0; i < L2; i++) {
            const int32_t a1 = A[2 * i * mod], a2 = A[2 * i * mod + 1], b1 = B[2 * i * mod];
            const int32_t s1r = pSrc[2 * i], s1i = pSrc[2 * i + 1], s2r = pSrc[2 * (L2 - i)], s2i = pSrc[2 * (L2 - i) + 1];
            int32_t outR = mulr(s1r, a1);
            int32_t outI = mulr(s1r, wneg(a2));
            outR = accr(outR, s1i, a2);
            outI = accr(outI, s1i, a1);
            outR = accr(outR, s2i, a2);
            outI = subr(outI, s2i, b1);
            outR = accr(outR, s2r, b1);
            outI = accr(outI, s2r, a2);
            pDst[2 * i] = outR;
            pDst[2 * i + 1] = outI;
        }
        orc_cfft_q31(L2, pDst, 1, bitReverseFlagR);
        /* arm_shift_q31(pDst, 1, pDst, N): clip_q63_to_q31((q63_t)x << 1) */
        for (uint32_t k = 0; k < N; k++) {
            const int64_t v = (int64_t)pDst[k] * 2;
            pDst[k] = v > INT32_MAX ? INT32_MAX : (v < INT32_MIN ? INT32_MIN : (int32_t)v);
        }
    } else {
        orc_cfft_q31(L2, pSrc, 0, bitReverseFlagR);
        /* arm_split_rfft_q31 (:257-342): i = 1..L2-1, S1 = X[i], S2 = X[L2-i] */
        for (uint32_t i = 1; i < L2; i++) {
            const int32_t a1 = A[2 * i * mod], a2 = A[2 * i * mod + 1], b1 = B[2 * i * mod];
            const int32_t s1r = pSrc[2 * i], s1i = pSrc[2 * i + 1], s2r = pSrc[2 * (L2 - i)], s2i = pSrc[2 * (L2 - i) + 1];
            int32_t outR = mulr(s1r, a1);
            int32_t outI = mulr(s1r, a2);
            outR = subr(outR, s1i, a2);
            outI = accr(outI, s1i, a1);
            outR = subr(outR, s2i, a2);
            outI = subr(outI, s2i, b1);
            outR = accr(outR, s2r, b1);
            outI = subr(outI, s2r, a2);
            pDst[2 * i] = outR;
            pDst[2 * i + 1] = outI;
            pDst[4 * L2 - 2 * i] = outR;
            pDst[4 * L2 - 2 * i + 1] = wneg(outI);
        }
        pDst[2 * L2] = (int32_t)((uint32_t)pSrc[0] - (uint32_t)pSrc[1]) >> 1;
        pDst[2 * L2 + 1] = 0;
        pDst[0] = (int32_t)((uint32_t)pSrc[0] + (uint32_t)pSrc[1]) >> 1;
        pDst[1] = 0;
    }
}

void orc_rfft_q15(uint32_t N, int16_t *pSrc, int16_t *pDst, int ifftFlagR, int bitReverseFlagR)
{
    if (!supported(N)) return;
    pthread_once(&g_once, build);
    const uint32_t L2 = N >> 1, mod = 8192u / N;
    const int16_t *A = g_a15, *B = g_b15;
    if (ifftFlagR) {
        /* arm_split_rifft_q15 generic branch (:552-570) */
        for (uint32_t i = 0; i < L2; i++) {
            const int32_t a0 = A[2 * i * mod], a1 = A[2 * i * mod + 1], b0 = B[2 * i * mod], b1 = B[2 * i * mod + 1];
            const int32_t s1r = pSrc[2 * i], s1i = pSrc[2 * i + 1], s2r = pSrc[2 * (L2 - i)], s2i = pSrc[2 * (L2 - i) + 1];
            uint32_t outR = (uint32_t)(s2r * b0);
            outR -= (uint32_t)(s2i * b1);
            outR += (uint32_t)(s1r * a0);
            outR += (uint32_t)(s1i * a1);
            uint32_t outI = (uint32_t)(s1i * a0);
            outI -= (uint32_t)(s1r * a1);
            outI -= (uint32_t)(s2r * b1);
            outI -= (uint32_t)(s2i * b0);
            pDst[2 * i] = (int16_t)((int32_t)outR >> 16);
            pDst[2 * i + 1] = (int16_t)((int32_t)outI >> 16);
        }
        orc_cfft_q15(L2, pDst, 1, bitReverseFlagR);
        /* arm_shift_q15(pDst, 1, pDst, N): __SSAT((q31_t)x << 1, 16) */
        for (uint32_t k = 0; k < N; k++) {
            const int32_t v = (int32_t)pDst[k] * 2;
            pDst[k] = (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
        }
    } else {
        orc_cfft_q15(L2, pSrc, 0, bitReverseFlagR);
        /* arm_split_rfft_q15 generic branch (:352-405) */
        for (uint32_t i = 1; i < L2; i++) {
            const int32_t a0 = A[2 * i * mod], a1 = A[2 * i * mod + 1], b0 = B[2 * i * mod], b1 = B[2 * i * mod + 1];
            const int32_t s1r = pSrc[2 * i], s1i = pSrc[2 * i + 1], s2r = pSrc[2 * (L2 - i)], s2i = pSrc[2 * (L2 - i) + 1];
            uint32_t outR = (uint32_t)(s1r * a0);
            outR -= (uint32_t)(s1i * a1);
            outR += (uint32_t)(s2r * b0);
            outR += (uint32_t)(s2i * b1);
            uint32_t outI = (uint32_t)(s2r * b1);
            outI -= (uint32_t)(s2i * b0);
            outI += (uint32_t)(s1i * a0);
            outI += (uint32_t)(s1r * a1);
            const int32_t r = (int32_t)outR >> 16, im = (int32_t)outI >> 16;
            pDst[2 * i] = (int16_t)r;
            pDst[2 * i + 1] = (int16_t)im;
            pDst[4 * L2 - 2 * i] = (int16_t)r;
            pDst[4 * L2 - 2 * i + 1] = (int16_t)(-im);
        }
        pDst[2 * L2] = (int16_t)(((int32_t)pSrc[0] - (int32_t)pSrc[1]) >> 1);
        pDst[2 * L2 + 1] = 0;
        pDst[0] = (int16_t)(((int32_t)pSrc[0] + (int32_t)pSrc[1]) >> 1);
        pDst[1] = 0;
    }
}

/* ---- batch drivers: forward frames N in -> 2N out, inverse frames 2N in (bins 0..N/2 read) -> N out ---- */
typedef struct { int q15; uint32_t N; const void *src; void *dst; uint64_t f0, f1; int ifft, bitrev; } rjob_t;

static void *rworker(void *arg)
{
    rjob_t *j = arg;
    const uint32_t N = j->N;
    const uint64_t inStride = j->ifft ? 2ull * N : N, outStride = j->ifft ? N : 2ull * N;
    if (j->q15) {
        int16_t *tmp = malloc(sizeof(int16_t) * 2 * N);
        for (uint64_t f = j->f0; f < j->f1; f++) {
            for (uint64_t k = 0; k < inStride; k++) tmp[k] = ((const int16_t *)j->src)[f * inStride + k];
            orc_rfft_q15(N, tmp, (int16_t *)j->dst + f * outStride, j->ifft, j->bitrev);
        }
        free(tmp);
    } else {
        int32_t *tmp = malloc(sizeof(int32_t) * 2 * N);
        for (uint64_t f = j->f0; f < j->f1; f++) {
            for (uint64_t k = 0; k < inStride; k++) tmp[k] = ((const int32_t *)j->src)[f * inStride + k];
            orc_rfft_q31(N, tmp, (int32_t *)j->dst + f * outStride, j->ifft, j->bitrev);
        }
        free(tmp);
    }
    return NULL;
}

static void rrun(int q15, uint32_t N, const void *src, void *dst, uint64_t nFrames, int ifft, int bitrev, int nthreads)
{
    if (!supported(N)) return;
    pthread_once(&g_once, build);
    (void)orc_twiddle_q31(16);
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nFrames) nthreads = nFrames ? (int)nFrames : 1;
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    rjob_t *jobs = malloc((size_t)nthreads * sizeof *jobs);
    const uint64_t per = (nFrames + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t f0 = per * (uint64_t)t, f1 = f0 + per;
        if (f0 > nFrames) f0 = nFrames;
        if (f1 > nFrames) f1 = nFrames;
        jobs[t] = (rjob_t){q15, N, src, dst, f0, f1, ifft, bitrev};
        if (nthreads == 1) rworker(&jobs[t]);
        else pthread_create(&th[t], NULL, rworker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

void orc_rfft_q31_batch(uint32_t N, const int32_t *src, int32_t *dst, uint64_t n, int ifft, int bitrev, int nt) { rrun(0, N, src, dst, n, ifft, bitrev, nt); }
void orc_rfft_q15_batch(uint32_t N, const int16_t *src, int16_t *dst, uint64_t n, int ifft, int bitrev, int nt) { rrun(1, N, src, dst, n, ifft, bitrev, nt); }
