/*
 * oracle/orc_rfft_f32.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restatement of the reference's generic-C arm_rfft_fast_f32: an N/2-point
 * CFFT on the packed real input followed by the split stage (forward), or the
 * merge stage followed by an N/2-point inverse CFFT (inverse).
 *
 *   arm_rfft_fast_f32   Source/TransformFunctions/arm_rfft_fast_f32.c:675-699
 *   stage_rfft_f32      arm_rfft_fast_f32.c:316-402
 *   merge_rfft_f32      arm_rfft_fast_f32.c:405-462
 */
#include "orc_fft.h"

/* forward split: out[0] = ReX0 + ImX0, out[1] = ReX0 - ImX0 (Nyquist packed),
 * bins k = 1..Nh-1 from A = X[k], B = X[Nh-k], tw = table[k]. */
static void split_stage(uint32_t Nh, const float *tw, const float *p, float *out)
{
    {
        float xBR = p[0], xBI = p[1], xAR = p[0], xAI = p[1];
        float t1a = xBR + xAR, t1b = xBI + xAI;
        out[0] = 0.5f * (t1a + t1b);
        out[1] = 0.5f * (t1a - t1b);
    }
    for (uint32_t k = 1; k < Nh; k++) {
        const float *pA = p + 2 * k, *pB = p + 2 * (Nh - k);
        float xBI = pB[1], xBR = pB[0], xAR = pA[0], xAI = pA[1];
        float twR = tw[2 * k], twI = tw[2 * k + 1];
        float t1a = xBR - xAR, t1b = xBI + xAI;
        float p0 = twR * t1a, p1 = twI * t1a, p2 = twR * t1b, p3 = twI * t1b;
        out[2 * k]     = 0.5f * (xAR + xBR + p0 + p3);
        out[2 * k + 1] = 0.5f * (xAI - xBI + p1 - p2);
    }
}

/* inverse merge: builds the N/2 complex points fed to the inverse CFFT. */
static void merge_stage(uint32_t Nh, const float *tw, const float *p, float *out)
{
    {
        float xAR = p[0], xAI = p[1];
        out[0] = 0.5f * (xAR + xAI);
        out[1] = 0.5f * (xAR - xAI);
    }
    for (uint32_t k = 1; k < Nh; k++) {
        const float *pA = p + 2 * k, *pB = p + 2 * (Nh - k);
        float xBI = pB[1], xBR = pB[0], xAR = pA[0], xAI = pA[1];
        float twR = tw[2 * k], twI = tw[2 * k + 1];
        float t1a = xAR - xBR, t1b = xAI + xBI;
        float r = twR * t1a, s = twI * t1b, t = twI * t1a, u = twR * t1b;
        out[2 * k]     = 0.5f * (xAR + xBR - r - s);
        out[2 * k + 1] = 0.5f * (xAI - xBI + t - u);
    }
}

void orc_rfft_fast_f32(uint32_t N, float *p, float *pOut, int ifftFlag)
{
    const float *tw = orc_twiddle_rfft_f32(N);
    if (!tw) return;
    uint32_t Nh = N / 2;
    if (ifftFlag) {
        merge_stage(Nh, tw, p, pOut);
        orc_cfft_f32(Nh, pOut, ifftFlag, 1);
    } else {
        orc_cfft_f32(Nh, p, ifftFlag, 1);
        split_stage(Nh, tw, p, pOut);
    }
}
