/*
 * oracle/orc_cfft_q15.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restatement of the reference's generic-C (non-ARM_MATH_DSP) arm_cfft_q15:
 * in-place radix-4 DIF on int16 with __SSAT at the reference's points, int32
 * products truncated by >>16, optional radix-2 pre-pass, binary bit reversal.
 *
 *   arm_cfft_q15                       Source/TransformFunctions/arm_cfft_q15.c:671-722
 *   arm_cfft_radix4by2_q15 / _inverse  arm_cfft_q15.c:782-827 / :881-927 (generic branch)
 *   arm_radix4_butterfly_q15           Source/TransformFunctions/arm_cfft_radix4_q15.c:572-970
 *   arm_radix4_butterfly_inverse_q15   arm_cfft_radix4_q15.c:1434-1813
 *   __SSAT                             Include/dsp/none.h:78-94
 *   arm_bitreversal_16                 Source/TransformFunctions/arm_bitreversal2.c:123-148
 */
#include "orc_fft.h"

typedef struct { int16_t re, im; } cq15;

static inline int32_t sat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
static inline int16_t q15(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }   /* wrapping store */

enum { ST_FIRST, ST_MIDDLE, ST_LAST };

/* forward: (co*x + si*y, -si*x + co*y) >> 16 ; inverse: (co*x - si*y, si*x + co*y) >> 16 */
static inline cq15 rot(int32_t x, int32_t y, int32_t co, int32_t si, int inv)
{
    cq15 o;
    if (!inv) { o.re = q15((co * x + si * y) >> 16); o.im = q15((-si * x + co * y) >> 16); }
    else      { o.re = q15((co * x - si * y) >> 16); o.im = q15((si * x + co * y) >> 16); }
    return o;
}

static inline void bfly(cq15 *d, uint32_t i0, uint32_t n2, const int16_t *tw, uint32_t ic,
                        int stage, int inv)
{
    int sh = (stage == ST_FIRST) ? 2 : 0;
    int32_t T0 = d[i0].re >> sh,          T1 = d[i0].im >> sh;
    int32_t S0 = d[i0 + 2 * n2].re >> sh, S1 = d[i0 + 2 * n2].im >> sh;
    int32_t B0 = d[i0 + n2].re >> sh,     B1 = d[i0 + n2].im >> sh;
    int32_t U0 = d[i0 + 3 * n2].re >> sh, U1 = d[i0 + 3 * n2].im >> sh;

    int32_t R0 = sat16(T0 + S0), R1 = sat16(T1 + S1);      /* xa+xc, ya+yc */
    S0 = sat16(T0 - S0); S1 = sat16(T1 - S1);              /* xa-xc, ya-yc */
    T0 = sat16(B0 + U0); T1 = sat16(B1 + U1);              /* xb+xd, yb+yd */
    int32_t D0 = sat16(B0 - U0), D1 = sat16(B1 - U1);      /* xb-xd, yb-yd */

    cq15 oa, oc, ob, od;
    if (stage == ST_FIRST) {
        /* radix4_q15.c:650-719 (fwd), :1512-1580 (inv) */
        oa.re = q15((R0 >> 1) + (T0 >> 1));
        oa.im = q15((R1 >> 1) + (T1 >> 1));
        R0 = sat16(R0 - T0); R1 = sat16(R1 - T1);
        oc = rot(R0, R1, tw[4 * ic], tw[4 * ic + 1], inv);
        int32_t P0, P1, Q0, Q1;                            /* P -> W^1 slot, Q -> W^3 slot */
        if (!inv) { P0 = sat16(S0 + D1); P1 = sat16(S1 - D0); Q0 = sat16(S0 - D1); Q1 = sat16(S1 + D0); }
        else      { P0 = sat16(S0 - D1); P1 = sat16(S1 + D0); Q0 = sat16(S0 + D1); Q1 = sat16(S1 - D0); }
        ob = rot(P0, P1, tw[2 * ic], tw[2 * ic + 1], inv);
        od = rot(Q0, Q1, tw[6 * ic], tw[6 * ic + 1], inv);
    } else if (stage == ST_MIDDLE) {
        /* radix4_q15.c:803-846 (fwd), :1664-1706 (inv) */
        oa.re = q15(((R0 >> 1) + (T0 >> 1)) >> 1);
        oa.im = q15(((R1 >> 1) + (T1 >> 1)) >> 1);
        R0 = (R0 >> 1) - (T0 >> 1); R1 = (R1 >> 1) - (T1 >> 1);
        oc = rot(R0, R1, tw[4 * ic], tw[4 * ic + 1], inv);
        int32_t P0, P1, Q0, Q1;
        if (!inv) { P0 = (S0 >> 1) + (D1 >> 1); P1 = (S1 >> 1) - (D0 >> 1); Q0 = (S0 >> 1) - (D1 >> 1); Q1 = (S1 >> 1) + (D0 >> 1); }
        else      { P0 = (S0 >> 1) - (D1 >> 1); P1 = (S1 >> 1) + (D0 >> 1); Q0 = (S0 >> 1) + (D1 >> 1); Q1 = (S1 >> 1) - (D0 >> 1); }
        ob = rot(P0, P1, tw[2 * ic], tw[2 * ic + 1], inv);
        od = rot(Q0, Q1, tw[6 * ic], tw[6 * ic + 1], inv);
    } else {
        /* last stage, twiddle free: radix4_q15.c:886-961 (fwd), :1745-1810 (inv) */
        oa.re = q15((R0 >> 1) + (T0 >> 1));
        oa.im = q15((R1 >> 1) + (T1 >> 1));
        oc.re = q15((R0 >> 1) - (T0 >> 1));
        oc.im = q15((R1 >> 1) - (T1 >> 1));
        cq15 p = {q15((S0 >> 1) + (D1 >> 1)), q15((S1 >> 1) - (D0 >> 1))};
        cq15 q = {q15((S0 >> 1) - (D1 >> 1)), q15((S1 >> 1) + (D0 >> 1))};
        if (!inv) { ob = p; od = q; } else { ob = q; od = p; }
    }
    d[i0] = oa; d[i0 + n2] = oc; d[i0 + 2 * n2] = ob; d[i0 + 3 * n2] = od;
}

static void radix4_passes(cq15 *d, uint32_t len, const int16_t *tw, uint32_t mod, int inv)
{
    uint32_t n2 = len >> 2;
    for (uint32_t i0 = 0; i0 < n2; i0++)
        bfly(d, i0, n2, tw, i0 * mod, ST_FIRST, inv);
    mod <<= 2;
    for (uint32_t k = len / 4; k > 4; k >>= 2) {
        uint32_t n1 = n2;
        n2 >>= 2;
        for (uint32_t j = 0; j < n2; j++)
            for (uint32_t i0 = j; i0 < len; i0 += n1)
                bfly(d, i0, n2, tw, j * mod, ST_MIDDLE, inv);
        mod <<= 2;
    }
    for (uint32_t i0 = 0; i0 < len; i0 += 4)
        bfly(d, i0, 1, tw, 0, ST_LAST, inv);
}

static void radix4by2(cq15 *d, uint32_t N, const int16_t *tw, int inv)
{
    uint32_t h = N >> 1;
    for (uint32_t i = 0; i < h; i++) {
        int32_t co = tw[2 * i], si = tw[2 * i + 1];
        cq15 a = d[i], b = d[i + h];
        int16_t xt = q15((a.re >> 1) - (b.re >> 1));
        int16_t yt = q15((a.im >> 1) - (b.im >> 1));
        d[i].re = q15(((a.re >> 1) + (b.re >> 1)) >> 1);
        d[i].im = q15(((b.im >> 1) + (a.im >> 1)) >> 1);
        if (!inv) {
            d[i + h].re = q15((int16_t)((xt * co) >> 16) + (int16_t)((yt * si) >> 16));
            d[i + h].im = q15((int16_t)((yt * co) >> 16) - (int16_t)((xt * si) >> 16));
        } else {
            d[i + h].re = q15((int16_t)((xt * co) >> 16) - (int16_t)((yt * si) >> 16));
            d[i + h].im = q15((int16_t)((yt * co) >> 16) + (int16_t)((xt * si) >> 16));
        }
    }
    radix4_passes(d, h, tw, 2, inv);
    radix4_passes(d + h, h, tw, 2, inv);
    for (uint32_t i = 0; i < N; i++) {
        d[i].re = q15((int32_t)d[i].re << 1);
        d[i].im = q15((int32_t)d[i].im << 1);
    }
}

void orc_cfft_q15(uint32_t N, int16_t *p, int ifftFlag, int bitReverseFlag)
{
    const int16_t *tw = orc_twiddle_q15(N);
    if (!tw) return;
    cq15 *d = (cq15 *)p;
    int inv = (ifftFlag == 1);
    switch (N) {
    case 16: case 64: case 256: case 1024: case 4096: radix4_passes(d, N, tw, 1, inv); break;
    default:                                          radix4by2(d, N, tw, inv); break;
    }
    if (bitReverseFlag) {
        uint16_t len;
        const uint16_t *tab = orc_bitrev_fixed(N, &len);
        for (uint32_t i = 0; i < len; i += 2) {
            uint32_t a = tab[i] >> 3, b = tab[i + 1] >> 3;   /* (>>2 on int16 words) / 2 */
            cq15 t = d[a]; d[a] = d[b]; d[b] = t;
        }
    }
}
