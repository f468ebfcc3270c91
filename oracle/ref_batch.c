/*
 * oracle/ref_batch.c -- TEST INFRASTRUCTURE.
 * Thin driver compiled TOGETHER WITH the reference's own sources (taken where
 * they lie under /root/reference; nothing is copied) into
 * oracle/_ref/libcmsisdsp_ref.so.  It only adds pthread batch loops around the
 * reference's public single-frame API and accessors for its preset instances,
 * so tests and bench.py can drive the real reference through ctypes.
 */
#include "arm_math_types.h"
#include "dsp/transform_functions.h"
#include "arm_const_structs.h"
#include "arm_common_tables.h"
#include <pthread.h>
#include <stdlib.h>

typedef struct {
    int kind; uint32_t N; void *p, *out; uint64_t f0, f1; int ifft, bitrev;
} job_t;

static void *worker(void *arg)
{
    job_t *j = arg;
    arm_cfft_instance_f32 Sf; arm_cfft_instance_q31 S31; arm_cfft_instance_q15 S15;
    arm_rfft_fast_instance_f32 Sr;
    switch (j->kind) {
    case 0: if (arm_cfft_init_f32(&Sf, (uint16_t)j->N) != ARM_MATH_SUCCESS) return NULL; break;
    case 1: if (arm_cfft_init_q31(&S31, (uint16_t)j->N) != ARM_MATH_SUCCESS) return NULL; break;
    case 2: if (arm_cfft_init_q15(&S15, (uint16_t)j->N) != ARM_MATH_SUCCESS) return NULL; break;
    default: if (arm_rfft_fast_init_f32(&Sr, (uint16_t)j->N) != ARM_MATH_SUCCESS) return NULL; break;
    }
    for (uint64_t f = j->f0; f < j->f1; f++) {
        switch (j->kind) {
        case 0: arm_cfft_f32(&Sf, (float32_t *)j->p + 2ull * j->N * f, (uint8_t)j->ifft, (uint8_t)j->bitrev); break;
        case 1: arm_cfft_q31(&S31, (q31_t *)j->p + 2ull * j->N * f, (uint8_t)j->ifft, (uint8_t)j->bitrev); break;
        case 2: arm_cfft_q15(&S15, (q15_t *)j->p + 2ull * j->N * f, (uint8_t)j->ifft, (uint8_t)j->bitrev); break;
        default:
            arm_rfft_fast_f32(&Sr, (float32_t *)j->p + (uint64_t)j->N * f,
                              (float32_t *)j->out + (uint64_t)j->N * f, (uint8_t)j->ifft);
            break;
        }
    }
    return NULL;
}

static void run(int kind, uint32_t N, void *p, void *out, uint64_t nFrames, int ifft, int bitrev, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nFrames) nthreads = nFrames ? (int)nFrames : 1;
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    job_t *jobs = malloc((size_t)nthreads * sizeof *jobs);
    uint64_t per = (nFrames + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t f0 = per * (uint64_t)t, f1 = f0 + per;
        if (f0 > nFrames) f0 = nFrames;
        if (f1 > nFrames) f1 = nFrames;
        jobs[t] = (job_t){kind, N, p, out, f0, f1, ifft, bitrev};
        if (nthreads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

void ref_cfft_f32_batch(uint32_t N, float *p, uint64_t n, int ifft, int bitrev, int nt) { run(0, N, p, 0, n, ifft, bitrev, nt); }
void ref_cfft_q31_batch(uint32_t N, int32_t *p, uint64_t n, int ifft, int bitrev, int nt) { run(1, N, p, 0, n, ifft, bitrev, nt); }
void ref_cfft_q15_batch(uint32_t N, int16_t *p, uint64_t n, int ifft, int bitrev, int nt) { run(2, N, p, 0, n, ifft, bitrev, nt); }
void ref_rfft_fast_f32_batch(uint32_t N, float *p, float *out, uint64_t n, int ifft, int nt) { run(3, N, p, out, n, ifft, 0, nt); }

/* table accessors (the instance structs carry the pointers) */
const float *ref_twiddle_f32(uint32_t N) { arm_cfft_instance_f32 S; return arm_cfft_init_f32(&S, (uint16_t)N) ? 0 : S.pTwiddle; }
const int32_t *ref_twiddle_q31(uint32_t N) { arm_cfft_instance_q31 S; return arm_cfft_init_q31(&S, (uint16_t)N) ? 0 : S.pTwiddle; }
const int16_t *ref_twiddle_q15(uint32_t N) { arm_cfft_instance_q15 S; return arm_cfft_init_q15(&S, (uint16_t)N) ? 0 : S.pTwiddle; }
const float *ref_twiddle_rfft_f32(uint32_t N) { arm_rfft_fast_instance_f32 S; return arm_rfft_fast_init_f32(&S, (uint16_t)N) ? 0 : S.pTwiddleRFFT; }
const uint16_t *ref_bitrev_f32(uint32_t N, uint16_t *len)
{ arm_cfft_instance_f32 S; if (arm_cfft_init_f32(&S, (uint16_t)N)) return 0; *len = S.bitRevLength; return S.pBitRevTable; }
const uint16_t *ref_bitrev_fixed(uint32_t N, uint16_t *len)
{ arm_cfft_instance_q31 S; if (arm_cfft_init_q31(&S, (uint16_t)N)) return 0; *len = S.bitRevLength; return S.pBitRevTable; }
uint32_t ref_sizeof_cfft_instance_f32(void) { return (uint32_t)sizeof(arm_cfft_instance_f32); }
uint32_t ref_sizeof_cfft_instance_q31(void) { return (uint32_t)sizeof(arm_cfft_instance_q31); }
uint32_t ref_sizeof_cfft_instance_q15(void) { return (uint32_t)sizeof(arm_cfft_instance_q15); }
uint32_t ref_sizeof_rfft_fast_instance_f32(void) { return (uint32_t)sizeof(arm_rfft_fast_instance_f32); }

/* ---- arm_mfcc_f32 batch driver (frames start `stride` floats apart; src is left untouched) ---- */
#include <string.h>
typedef struct {
    arm_mfcc_instance_f32 S; const float *src; uint64_t stride; float *dst; uint64_t f0, f1;
} mjob_t;
static void *mworker(void *arg)
{
    mjob_t *j = arg;
    const uint32_t n = j->S.fftLen;
    float *frame = malloc(sizeof(float) * n), *tmp = malloc(sizeof(float) * 2 * n);
    for (uint64_t f = j->f0; f < j->f1; f++) {
        memcpy(frame, j->src + f * j->stride, sizeof(float) * n);
        arm_mfcc_f32(&j->S, frame, j->dst + f * j->S.nbDctOutputs, tmp);
    }
    free(frame); free(tmp);
    return NULL;
}
int ref_mfcc_f32_batch(uint32_t fftLen, uint32_t nbMel, uint32_t nbDct, const float *dct, const uint32_t *pos,
                       const uint32_t *len, const float *coefs, const float *window, const float *src,
                       uint64_t stride, float *dst, uint64_t nFrames, int nthreads)
{
    arm_mfcc_instance_f32 S;
    if (arm_mfcc_init_f32(&S, fftLen, nbMel, nbDct, dct, pos, len, coefs, window) != ARM_MATH_SUCCESS) return -1;
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nFrames) nthreads = nFrames ? (int)nFrames : 1;
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    mjob_t *jobs = malloc((size_t)nthreads * sizeof *jobs);
    uint64_t per = (nFrames + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t f0 = per * (uint64_t)t, f1 = f0 + per;
        if (f0 > nFrames) f0 = nFrames;
        if (f1 > nFrames) f1 = nFrames;
        jobs[t] = (mjob_t){S, src, stride, dst, f0, f1};
        if (nthreads == 1) mworker(&jobs[t]);
        else pthread_create(&th[t], NULL, mworker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
uint32_t ref_sizeof_mfcc_instance_f32(void) { return (uint32_t)sizeof(arm_mfcc_instance_f32); }

/* ---- arm_rfft_q31 / arm_rfft_q15 batch drivers: forward frames N in -> 2N out, inverse frames 2N in -> N out;
 * the source is copied per frame (the reference's forward transform destroys it) ---- */
typedef struct { int q15; uint32_t N; const void *src; void *dst; uint64_t f0, f1; int ifft, bitrev; } rjob_t;
static void *rworker(void *arg)
{
    rjob_t *j = arg;
    const uint32_t N = j->N;
    const uint64_t inStride = j->ifft ? 2ull * N : N, outStride = j->ifft ? N : 2ull * N;
    arm_rfft_instance_q31 S31; arm_rfft_instance_q15 S15;
    if (j->q15) { if (arm_rfft_init_q15(&S15, N, (uint32_t)j->ifft, (uint32_t)j->bitrev) != ARM_MATH_SUCCESS) return NULL; }
    else        { if (arm_rfft_init_q31(&S31, N, (uint32_t)j->ifft, (uint32_t)j->bitrev) != ARM_MATH_SUCCESS) return NULL; }
    void *tmp = malloc((j->q15 ? 2 : 4) * 2 * (size_t)N);
    for (uint64_t f = j->f0; f < j->f1; f++) {
        if (j->q15) {
            memcpy(tmp, (const q15_t *)j->src + f * inStride, 2 * inStride);
            arm_rfft_q15(&S15, (q15_t *)tmp, (q15_t *)j->dst + f * outStride);
        } else {
            memcpy(tmp, (const q31_t *)j->src + f * inStride, 4 * inStride);
            arm_rfft_q31(&S31, (q31_t *)tmp, (q31_t *)j->dst + f * outStride);
        }
    }
    free(tmp);
    return NULL;
}
static void rrun(int q15, uint32_t N, const void *src, void *dst, uint64_t nFrames, int ifft, int bitrev, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nFrames) nthreads = nFrames ? (int)nFrames : 1;
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    rjob_t *jobs = malloc((size_t)nthreads * sizeof *jobs);
    uint64_t per = (nFrames + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
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
void ref_rfft_q31_batch(uint32_t N, const int32_t *src, int32_t *dst, uint64_t n, int ifft, int bitrev, int nt) { rrun(0, N, src, dst, n, ifft, bitrev, nt); }
void ref_rfft_q15_batch(uint32_t N, const int16_t *src, int16_t *dst, uint64_t n, int ifft, int bitrev, int nt) { rrun(1, N, src, dst, n, ifft, bitrev, nt); }
const int32_t *ref_real_coef_q31(int b) { return b ? realCoefBQ31 : realCoefAQ31; }
const int16_t *ref_real_coef_q15(int b) { return b ? realCoefBQ15 : realCoefAQ15; }
uint32_t ref_sizeof_rfft_instance_q31(void) { return (uint32_t)sizeof(arm_rfft_instance_q31); }
uint32_t ref_sizeof_rfft_instance_q15(void) { return (uint32_t)sizeof(arm_rfft_instance_q15); }

/* ---- arm_cfft_f32 + arm_cmplx_mag[_squared]_f32 (+ arm_max_f32): the reference's own functions, frame by frame ---- */
#include "dsp/complex_math_functions.h"
#include "dsp/statistics_functions.h"
void ref_cfft_mag_f32_batch(uint32_t N, const float *src, float *mag, float *val, uint32_t *idx, uint64_t nFrames,
                            int ifftFlag, int squared)
{
    arm_cfft_instance_f32 S;
    if (arm_cfft_init_f32(&S, (uint16_t)N) != ARM_MATH_SUCCESS) return;
    float *p = malloc(sizeof(float) * 2 * N), *m = malloc(sizeof(float) * N);
    for (uint64_t f = 0; f < nFrames; f++) {
        memcpy(p, src + f * 2 * N, sizeof(float) * 2 * N);
        arm_cfft_f32(&S, p, (uint8_t)ifftFlag, 1);
        if (squared) arm_cmplx_mag_squared_f32(p, m, N);
        else arm_cmplx_mag_f32(p, m, N);
        if (mag) memcpy(mag + f * N, m, sizeof(float) * N);
        if (val && idx) arm_max_f32(m, N, val + f, idx + f);
    }
    free(p); free(m);
}

/* ---- deprecated radix-4 / radix-2 instance API of the reference, frame by frame (kind 0 f32, 1 q31, 2 q15) ---- */
int ref_cfft_radix_batch(int kind, int radix, uint32_t N, void *p, uint64_t nFrames, int ifft, int bitrev)
{
    arm_cfft_radix4_instance_f32 Sf4; arm_cfft_radix2_instance_f32 Sf2; arm_cfft_radix4_instance_q31 S31; arm_cfft_radix4_instance_q15 S15;
    arm_cfft_radix2_instance_q31 R31; arm_cfft_radix2_instance_q15 R15;
    arm_status st = ARM_MATH_ARGUMENT_ERROR;
    if (kind == 1 && radix == 2) st = arm_cfft_radix2_init_q31(&R31, (uint16_t)N, (uint8_t)ifft, (uint8_t)bitrev);
    if (kind == 2 && radix == 2) st = arm_cfft_radix2_init_q15(&R15, (uint16_t)N, (uint8_t)ifft, (uint8_t)bitrev);
    if (kind == 0 && radix == 4) st = arm_cfft_radix4_init_f32(&Sf4, (uint16_t)N, (uint8_t)ifft, (uint8_t)bitrev);
    if (kind == 0 && radix == 2) st = arm_cfft_radix2_init_f32(&Sf2, (uint16_t)N, (uint8_t)ifft, (uint8_t)bitrev);
    if (kind == 1 && radix == 4) st = arm_cfft_radix4_init_q31(&S31, (uint16_t)N, (uint8_t)ifft, (uint8_t)bitrev);
    if (kind == 2 && radix == 4) st = arm_cfft_radix4_init_q15(&S15, (uint16_t)N, (uint8_t)ifft, (uint8_t)bitrev);
    if (st != ARM_MATH_SUCCESS) return -1;
    for (uint64_t f = 0; f < nFrames; f++) {
        if (kind == 0 && radix == 4) arm_cfft_radix4_f32(&Sf4, (float32_t *)p + 2ull * N * f);
        else if (kind == 0) arm_cfft_radix2_f32(&Sf2, (float32_t *)p + 2ull * N * f);
        else if (kind == 1 && radix == 2) arm_cfft_radix2_q31(&R31, (q31_t *)p + 2ull * N * f);
        else if (kind == 2 && radix == 2) arm_cfft_radix2_q15(&R15, (q15_t *)p + 2ull * N * f);
        else if (kind == 1) arm_cfft_radix4_q31(&S31, (q31_t *)p + 2ull * N * f);
        else arm_cfft_radix4_q15(&S15, (q15_t *)p + 2ull * N * f);
    }
    return 0;
}
const uint16_t *ref_arm_bit_rev_table(void) { return armBitRevTable; }
uint32_t ref_sizeof_cfft_radix4_instance_f32(void) { return (uint32_t)sizeof(arm_cfft_radix4_instance_f32); }
uint32_t ref_sizeof_cfft_radix4_instance_q31(void) { return (uint32_t)sizeof(arm_cfft_radix4_instance_q31); }
uint32_t ref_sizeof_cfft_radix4_instance_q15(void) { return (uint32_t)sizeof(arm_cfft_radix4_instance_q15); }
uint32_t ref_sizeof_cfft_radix2_instance_q31(void) { return (uint32_t)sizeof(arm_cfft_radix2_instance_q31); }
uint32_t ref_sizeof_cfft_radix2_instance_q15(void) { return (uint32_t)sizeof(arm_cfft_radix2_instance_q15); }

/* ---- arm_cfft_f64, frame by frame (single thread: used for parity only) ---- */
int ref_cfft_f64_batch(uint32_t N, double *p, uint64_t nFrames, int ifft, int bitrev)
{
    arm_cfft_instance_f64 S;
    if (arm_cfft_init_f64(&S, (uint16_t)N) != ARM_MATH_SUCCESS) return -1;
    for (uint64_t f = 0; f < nFrames; f++) arm_cfft_f64(&S, p + 2ull * N * f, (uint8_t)ifft, (uint8_t)bitrev);
    return 0;
}
const double *ref_twiddle_f64(uint32_t N) { arm_cfft_instance_f64 S; return arm_cfft_init_f64(&S, (uint16_t)N) ? 0 : S.pTwiddle; }
const uint16_t *ref_bitrev_f64(uint32_t N, uint16_t *len)
{ arm_cfft_instance_f64 S; if (arm_cfft_init_f64(&S, (uint16_t)N)) return 0; *len = S.bitRevLength; return S.pBitRevTable; }
uint32_t ref_sizeof_cfft_instance_f64(void) { return (uint32_t)sizeof(arm_cfft_instance_f64); }

/* ---- arm_rfft_fast_f64, frame by frame; the source is copied per frame (the forward transform destroys it) ---- */
int ref_rfft_fast_f64_batch(uint32_t N, const double *p, double *pOut, uint64_t nFrames, int ifft)
{
    arm_rfft_fast_instance_f64 S;
    if (arm_rfft_fast_init_f64(&S, (uint16_t)N) != ARM_MATH_SUCCESS) return -1;
    double *tmp = malloc(sizeof(double) * N);
    for (uint64_t f = 0; f < nFrames; f++) {
        memcpy(tmp, p + f * N, sizeof(double) * N);
        arm_rfft_fast_f64(&S, tmp, pOut + f * N, (uint8_t)ifft);
    }
    free(tmp);
    return 0;
}
const double *ref_twiddle_rfft_f64(uint32_t N) { arm_rfft_fast_instance_f64 S; return arm_rfft_fast_init_f64(&S, (uint16_t)N) ? 0 : S.pTwiddleRFFT; }
uint32_t ref_sizeof_rfft_fast_instance_f64(void) { return (uint32_t)sizeof(arm_rfft_fast_instance_f64); }
