/*
 * oracle/orc_mfcc_f32.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 * CPU restatement of arm_mfcc_f32 (generic C, RFFT based):
 *   Source/TransformFunctions/arm_mfcc_f32.c:88-174, with the helper kernels it calls:
 *   arm_absmax_f32   Source/StatisticsFunctions/arm_absmax_f32.c (generic, no loop unrolling)
 *   arm_scale_f32    Source/BasicMathFunctions/arm_scale_f32.c
 *   arm_mult_f32     Source/BasicMathFunctions/arm_mult_f32.c
 *   arm_cmplx_mag_f32 Source/ComplexMathFunctions/arm_cmplx_mag_f32.c:155-266 (sqrtf(re*re + im*im))
 *   arm_dot_prod_f32 Source/BasicMathFunctions/arm_dot_prod_f32.c (sequential sum)
 *   arm_offset_f32, arm_vlog_f32 (= logf, Source/FastMathFunctions/arm_vlog_f32.c:104-110)
 *   arm_mat_vec_mult_f32 Source/MatrixFunctions/arm_mat_vec_mult_f32.c (sequential per row)
 * tests/test_oracle_vs_ref.py checks it bit for bit against the compiled reference.
 */
#include "orc_fft.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* pSrc (fftLen floats) is destroyed, pTmp needs 2*fftLen floats, like the reference */
void orc_mfcc_f32(uint32_t fftLen, uint32_t nbMel, uint32_t nbDct, const float *dct, const uint32_t *pos,
                  const uint32_t *len, const float *coefs, const float *window, float *pSrc, float *pDst, float *pTmp)
{
    float maxValue = 0.0f;
    for (uint32_t i = 0; i < fftLen; i++) {                        /* arm_absmax_f32 */
        float a = fabsf(pSrc[i]);
        if (a > maxValue) maxValue = a;
    }
    if (maxValue != 0.0f) {
        const float s = 1.0f / maxValue;
        for (uint32_t i = 0; i < fftLen; i++) pSrc[i] = pSrc[i] * s;   /* arm_scale_f32 */
    }
    for (uint32_t i = 0; i < fftLen; i++) pSrc[i] = pSrc[i] * window[i];   /* arm_mult_f32 */
    orc_rfft_fast_f32(fftLen, pSrc, pTmp, 0);                      /* arm_mfcc_f32.c:137 */
    pTmp[1] = 0.0f;                                                /* :138 drops the packed Nyquist bin */
    /* :141 asks for fftLen magnitudes; only the first fftLen/2 read transform output, the rest is
     * never used by the filters -- the restatement computes the valid half */
    for (uint32_t k = 0; k < fftLen / 2; k++) {
        const float re = pTmp[2 * k], im = pTmp[2 * k + 1];
        pSrc[k] = sqrtf(re * re + im * im);
    }
    if (maxValue != 0.0f)
        for (uint32_t k = 0; k < fftLen / 2; k++) pSrc[k] = pSrc[k] * maxValue;
    const float *c = coefs;
    for (uint32_t f = 0; f < nbMel; f++) {                         /* :150-161 */
        float sum = 0.0f;
        for (uint32_t t = 0; t < len[f]; t++) sum += pSrc[pos[f] + t] * c[t];
        c += len[f];
        pTmp[f] = sum;
    }
    for (uint32_t f = 0; f < nbMel; f++) pTmp[f] = logf(pTmp[f] + 1.0e-6f);   /* :164-165 */
    for (uint32_t r = 0; r < nbDct; r++) {                         /* :171 */
        float sum = 0.0f;
        for (uint32_t f = 0; f < nbMel; f++) sum += dct[r * nbMel + f] * pTmp[f];
        pDst[r] = sum;
    }
}

typedef struct {
    uint32_t fftLen, nbMel, nbDct; const float *dct; const uint32_t *pos, *len; const float *coefs, *window;
    const float *src; uint64_t stride; float *dst; uint64_t f0, f1;
} mjob_t;

static void *mworker(void *arg)
{
    mjob_t *j = arg;
    float *frame = malloc(sizeof(float) * j->fftLen), *tmp = malloc(sizeof(float) * 2 * j->fftLen);
    for (uint64_t f = j->f0; f < j->f1; f++) {
        memcpy(frame, j->src + f * j->stride, sizeof(float) * j->fftLen);
        orc_mfcc_f32(j->fftLen, j->nbMel, j->nbDct, j->dct, j->pos, j->len, j->coefs, j->window, frame,
                     j->dst + f * j->nbDct, tmp);
    }
    free(frame); free(tmp);
    return NULL;
}

/* frame f starts at src + f*stride floats (stride = fftLen for back-to-back frames, < fftLen for overlap);
 * src is left untouched */
void orc_mfcc_f32_batch(uint32_t fftLen, uint32_t nbMel, uint32_t nbDct, const float *dct, const uint32_t *pos,
                        const uint32_t *len, const float *coefs, const float *window, const float *src,
                        uint64_t stride, float *dst, uint64_t nFrames, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nFrames) nthreads = nFrames ? (int)nFrames : 1;
    (void)orc_twiddle_f32(16);
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    mjob_t *jobs = malloc((size_t)nthreads * sizeof *jobs);
    uint64_t per = (nFrames + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t f0 = per * (uint64_t)t, f1 = f0 + per;
        if (f0 > nFrames) f0 = nFrames;
        if (f1 > nFrames) f1 = nFrames;
        jobs[t] = (mjob_t){fftLen, nbMel, nbDct, dct, pos, len, coefs, window, src, stride, dst, f0, f1};
        if (nthreads == 1) mworker(&jobs[t]);
        else pthread_create(&th[t], NULL, mworker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}
