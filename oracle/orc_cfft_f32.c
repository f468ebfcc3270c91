/*
 * oracle/orc_cfft_f32.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restatement of the reference's generic-C arm_cfft_f32: mixed radix-8/4/2
 * decimation-in-frequency, in place, followed by the swap-list permutation.
 * Every floating-point operation is performed in the reference's order so the
 * result is bit-identical when both are built with -ffp-contract=off.
 *
 *   arm_cfft_f32                Source/TransformFunctions/arm_cfft_f32.c:1243-1298
 *   arm_cfft_radix8by2_f32      arm_cfft_f32.c:846-958
 *   arm_cfft_radix8by4_f32      arm_cfft_f32.c:960-1201
 *   arm_radix8_butterfly_f32    Source/TransformFunctions/arm_cfft_radix8_f32.c:51-291
 *   arm_bitreversal_32          Source/TransformFunctions/arm_bitreversal2.c:84-108
 */
#include "orc_fft.h"

typedef struct { float re, im; } c32;

/* out = v * conj(w), with w = (co, si): (co*re + si*im, co*im - si*re).
 * Four separate products, then one add / one subtract (radix8_f32.c:220-225). */
static inline c32 mul_conj_tw(c32 v, float co, float si)
{
    float p1 = co * v.re, p2 = si * v.im, p3 = co * v.im, p4 = si * v.re;
    c32 o = {p1 + p2, p3 - p4};
    return o;
}

/* 8-point DFT kernel without twiddles; x[m] is the point at offset m*n2, and
 * y[m] is what the reference leaves in (or, for the twiddled columns, is about
 * to scale into) that same slot.  Operation order: radix8_f32.c:87-138 / :188-257. */
static inline void dft8_slots(const c32 x[8], c32 y[8])
{
    const float C81 = 0.70710678118f;
    float r1 = x[0].re + x[4].re, r5 = x[0].re - x[4].re;
    float r2 = x[1].re + x[5].re, r6 = x[1].re - x[5].re;
    float r3 = x[2].re + x[6].re, r7 = x[2].re - x[6].re;
    float r4 = x[3].re + x[7].re, r8 = x[3].re - x[7].re;
    float t1 = r1 - r3;
    r1 = r1 + r3;
    r3 = r2 - r4;
    r2 = r2 + r4;
    y[0].re = r1 + r2;
    y[4].re = r1 - r2;

    float s1 = x[0].im + x[4].im, s5 = x[0].im - x[4].im;
    float s2 = x[1].im + x[5].im, s6 = x[1].im - x[5].im;
    float s3 = x[2].im + x[6].im, s7 = x[2].im - x[6].im;
    float s4 = x[3].im + x[7].im, s8 = x[3].im - x[7].im;
    float t2 = s1 - s3;
    s1 = s1 + s3;
    s3 = s2 - s4;
    s2 = s2 + s4;
    y[0].im = s1 + s2;
    y[4].im = s1 - s2;
    y[2].re = t1 + s3;
    y[6].re = t1 - s3;
    y[2].im = t2 - r3;
    y[6].im = t2 + r3;

    float u1 = (r6 - r8) * C81;
    float u6 = (r6 + r8) * C81;
    float v1 = (s6 - s8) * C81;
    float v6 = (s6 + s8) * C81;
    float a1 = r5 - u1;      /* t1 */
    float a5 = r5 + u1;      /* r5 */
    float a8 = r7 - u6;      /* r8 */
    float a7 = r7 + u6;      /* r7 */
    float b2 = s5 - v1;      /* t2 */
    float b5 = s5 + v1;      /* s5 */
    float b8 = s7 - v6;      /* s8 */
    float b7 = s7 + v6;      /* s7 */
    y[1].re = a5 + b7;  y[1].im = b5 - a7;
    y[7].re = a5 - b7;  y[7].im = b5 + a7;
    y[5].re = a1 + b8;  y[5].im = b2 - a8;
    y[3].re = a1 - b8;  y[3].im = b2 + a8;
}

/* log8(len) in-place DIF radix-8 passes over `len` complex points.
 * tw = N-point (cos,+sin) table, mod = table stride of this sub-transform. */
static void radix8_passes(c32 *d, uint32_t len, const float *tw, uint32_t mod)
{
    uint32_t n2 = len;
    do {
        uint32_t n1 = n2;
        n2 >>= 3;
        /* column j = 0: no twiddles (radix8_f32.c:78-141) */
        for (uint32_t i = 0; i < len; i += n1) {
            c32 x[8], y[8];
            for (int m = 0; m < 8; m++) x[m] = d[i + (uint32_t)m * n2];
            dft8_slots(x, y);
            for (int m = 0; m < 8; m++) d[i + (uint32_t)m * n2] = y[m];
        }
        if (n2 < 8) break;
        /* columns j = 1..n2-1: slot m is scaled by conj(W^(m*j*mod)) (radix8_f32.c:149-287) */
        for (uint32_t j = 1; j < n2; j++) {
            uint32_t id = j * mod;
            for (uint32_t i = j; i < len; i += n1) {
                c32 x[8], y[8];
                for (int m = 0; m < 8; m++) x[m] = d[i + (uint32_t)m * n2];
                dft8_slots(x, y);
                d[i] = y[0];
                for (uint32_t m = 1; m < 8; m++)
                    d[i + m * n2] = mul_conj_tw(y[m], tw[2 * m * id], tw[2 * m * id + 1]);
            }
        }
        mod <<= 3;
    } while (n2 > 7);
}

/* N in {16,128,1024}: one radix-2 DIF pass, then two N/2 radix-8 transforms
 * (arm_cfft_f32.c:846-958). */
static void radix8by2(c32 *d, uint32_t N, const float *tw)
{
    uint32_t h = N / 2, q = N / 4;
    for (uint32_t i = 0; i < q; i++) {
        c32 a = d[i], b = d[i + h], c = d[i + q], e = d[i + h + q];
        float twR = tw[2 * i], twI = tw[2 * i + 1];
        d[i].re = a.re + b.re;  d[i].im = a.im + b.im;
        c32 t2 = {a.re - b.re, a.im - b.im};
        d[i + q].re = c.re + e.re;  d[i + q].im = c.im + e.im;
        c32 t4 = {e.re - c.re, e.im - c.im};           /* note: q3 - q1 (:904-907) */
        float m0 = t2.re * twR, m1 = t2.im * twI, m2 = t2.im * twR, m3 = t2.re * twI;
        d[i + h].re = m0 + m1;
        d[i + h].im = m2 - m3;
        /* quarter-wave symmetry of the table (:923-931) */
        m0 = t4.re * twI; m1 = t4.im * twR; m2 = t4.im * twI; m3 = t4.re * twR;
        d[i + h + q].re = m0 - m1;
        d[i + h + q].im = m2 + m3;
    }
    radix8_passes(d, h, tw, 2);
    radix8_passes(d + h, h, tw, 2);
}

/* N in {32,256,2048}: one radix-4 DIF pass walking up from 0 and down from Q,
 * then four N/4 radix-8 transforms (arm_cfft_f32.c:960-1201). */
static void radix8by4(c32 *d, uint32_t N, const float *tw)
{
    uint32_t Q = N / 4;
    c32 *c1 = d, *c2 = d + Q, *c3 = d + 2 * Q, *c4 = d + 3 * Q;

    for (uint32_t i = 0; i <= Q / 2; i++) {
        /* TOP half-butterfly at index i (:993-1009, :1026-1041, :1138-1154) */
        c32 p1 = c1[i], p2 = c2[i], p3 = c3[i], p4 = c4[i];
        float ap0 = p1.re + p3.re, sp0 = p1.re - p3.re;
        float ap1 = p1.im + p3.im, sp1 = p1.im - p3.im;
        c32 t2 = {sp0 + p2.im - p4.im, sp1 - p2.re + p4.re};
        c32 t3 = {ap0 - p2.re - p4.re, ap1 - p2.im - p4.im};
        c32 t4 = {sp0 - p2.im + p4.im, sp1 + p2.re - p4.re};
        c1[i].re = ap0 + p2.re + p4.re;
        c1[i].im = ap1 + p2.im + p4.im;

        if (i == 0) {                      /* twiddles are ones (:1011-1017) */
            c2[0] = t2; c3[0] = t3; c4[0] = t4;
            continue;
        }
        if (i < Q / 2) {
            /* BOTTOM half-butterfly at index e = Q - i (:1043-1059) */
            uint32_t e = Q - i;
            c32 e1 = c1[e], e2 = c2[e], e3 = c3[e], e4 = c4[e];
            float bar = e1.re + e3.re, bsr = e1.re - e3.re;   /* p1ap3_1 / p1sp3_1 */
            float bai = e1.im + e3.im, bsi = e1.im - e3.im;   /* p1ap3_0 / p1sp3_0 */
            float t2_2 = e2.im - e4.im + bsr;
            float t2_3 = e1.im - e3.im - e2.re + e4.re;
            float t3_2 = bar - e2.re - e4.re;
            float t3_3 = bai - e2.im - e4.im;
            float t4_2 = e2.im - e4.im - bsr;
            float t4_3 = e4.re - e2.re - bsi;
            c1[e].im = bai + e2.im + e4.im;
            c1[e].re = bar + e2.re + e4.re;

            float twR, twI, m0, m1, m2, m3;
            /* COL 2, W^i (:1061-1086) */
            twR = tw[2 * i]; twI = tw[2 * i + 1];
            m0 = t2.re * twR; m1 = t2.im * twI; m2 = t2.im * twR; m3 = t2.re * twI;
            c2[i].re = m0 + m1; c2[i].im = m2 - m3;
            m0 = t2_3 * twI; m1 = t2_2 * twR; m2 = t2_2 * twI; m3 = t2_3 * twR;
            c2[e].im = m0 - m1; c2[e].re = m2 + m3;
            /* COL 3, W^2i (:1088-1109) */
            twR = tw[4 * i]; twI = tw[4 * i + 1];
            m0 = t3.re * twR; m1 = t3.im * twI; m2 = t3.im * twR; m3 = t3.re * twI;
            c3[i].re = m0 + m1; c3[i].im = m2 - m3;
            m0 = -t3_3 * twR; m1 = t3_2 * twI; m2 = t3_2 * twR; m3 = t3_3 * twI;
            c3[e].im = m0 - m1; c3[e].re = m3 - m2;
            /* COL 4, W^3i (:1111-1132) */
            twR = tw[6 * i]; twI = tw[6 * i + 1];
            m0 = t4.re * twR; m1 = t4.im * twI; m2 = t4.im * twR; m3 = t4.re * twI;
            c4[i].re = m0 + m1; c4[i].im = m2 - m3;
            m0 = t4_3 * twI; m1 = t4_2 * twR; m2 = t4_2 * twI; m3 = t4_3 * twR;
            c4[e].im = m0 - m1; c4[e].re = m2 + m3;
        } else {
            /* MIDDLE, i == Q/2: top only (:1135-1188) */
            float twR, twI, m0, m1, m2, m3;
            twR = tw[2 * i]; twI = tw[2 * i + 1];
            m0 = t2.re * twR; m1 = t2.im * twI; m2 = t2.im * twR; m3 = t2.re * twI;
            c2[i].re = m0 + m1; c2[i].im = m2 - m3;
            twR = tw[4 * i]; twI = tw[4 * i + 1];
            m0 = t3.re * twR; m1 = t3.im * twI; m2 = t3.im * twR; m3 = t3.re * twI;
            c3[i].re = m0 + m1; c3[i].im = m2 - m3;
            twR = tw[6 * i]; twI = tw[6 * i + 1];
            m0 = t4.re * twR; m1 = t4.im * twI; m2 = t4.im * twR; m3 = t4.re * twI;
            c4[i].re = m0 + m1; c4[i].im = m2 - m3;
        }
    }
    radix8_passes(c1, Q, tw, 4);
    radix8_passes(c2, Q, tw, 4);
    radix8_passes(c3, Q, tw, 4);
    radix8_passes(c4, Q, tw, 4);
}

/* arm_bitreversal_32 (arm_bitreversal2.c:84-108) on 64-bit (re,im) pairs. */
static void apply_swaps_c32(c32 *d, const uint16_t *tab, uint16_t len)
{
    for (uint32_t i = 0; i < len; i += 2) {
        uint32_t a = tab[i] >> 3, b = tab[i + 1] >> 3;
        c32 t = d[a]; d[a] = d[b]; d[b] = t;
    }
}

void orc_cfft_f32(uint32_t N, float *p, int ifftFlag, int bitReverseFlag)
{
    const float *tw = orc_twiddle_f32(N);
    if (!tw) return;                                   /* unsupported length: no-op, like the switch */
    c32 *d = (c32 *)p;
    if (ifftFlag == 1)
        for (uint32_t l = 0; l < N; l++) d[l].im = -d[l].im;

    switch (N) {
    case 16: case 128: case 1024: radix8by2(d, N, tw); break;
    case 32: case 256: case 2048: radix8by4(d, N, tw); break;
    default:                      radix8_passes(d, N, tw, 1); break;
    }
    if (bitReverseFlag) {
        uint16_t len;
        const uint16_t *tab = orc_bitrev_f32(N, &len);
        apply_swaps_c32(d, tab, len);
    }
    if (ifftFlag == 1) {
        float invL = 1.0f / (float)N;
        for (uint32_t l = 0; l < N; l++) {
            d[l].re *= invL;
            d[l].im = -d[l].im * invL;
        }
    }
}

/* ------------------------------------------------------------------ spectrum epilogues
 * arm_cfft_f32 followed by arm_cmplx_mag_f32 (Source/ComplexMathFunctions/arm_cmplx_mag_f32.c:252-264:
 * sqrtf(re*re + im*im) through arm_sqrt_f32, Include/dsp/fast_math_functions.h:234-292) or
 * arm_cmplx_mag_squared_f32 (arm_cmplx_mag_squared_f32.c: re*re + im*im), and arm_max_f32
 * (Source/StatisticsFunctions/arm_max_f32.c generic loop: the FIRST maximum wins) -- the pipeline of
 * Examples/ARM/arm_fft_bin_example/arm_fft_bin_example_f32.c:141-149.  p is transformed in place. */
#include <math.h>
#include <stdlib.h>
void orc_cfft_mag_f32(uint32_t N, float *p, float *mag, int ifftFlag, int squared)
{
    orc_cfft_f32(N, p, ifftFlag, 1);
    for (uint32_t k = 0; k < N; k++) {
        const float re = p[2 * k], im = p[2 * k + 1];
        const float s = (re * re) + (im * im);
        mag[k] = squared ? s : (s >= 0.0f ? sqrtf(s) : 0.0f);
    }
}
void orc_max_f32(const float *src, uint32_t n, float *val, uint32_t *idx)
{
    float out = src[0];
    uint32_t outIndex = 0;
    for (uint32_t k = 1; k < n; k++)
        if (out < src[k]) { out = src[k]; outIndex = k; }
    *val = out;
    *idx = outIndex;
}
/* batch: src nFrames*2N floats (untouched), mag nFrames*N floats (may be NULL), val/idx nFrames (may be NULL) */
void orc_cfft_mag_f32_batch(uint32_t N, const float *src, float *mag, float *val, uint32_t *idx, uint64_t nFrames,
                            int ifftFlag, int squared)
{
    float *p = malloc(sizeof(float) * 2 * N), *m = malloc(sizeof(float) * N);
    for (uint64_t f = 0; f < nFrames; f++) {
        for (uint32_t k = 0; k < 2 * N; k++) p[k] = src[f * 2 * N + k];
        orc_cfft_mag_f32(N, p, m, ifftFlag, squared);
        if (mag) for (uint32_t k = 0; k < N; k++) mag[f * N + k] = m[k];
        if (val && idx) orc_max_f32(m, N, val + f, idx + f);
    }
    free(p); free(m);
}
