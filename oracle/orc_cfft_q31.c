/*
 * oracle/orc_cfft_q31.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restatement of the reference's generic-C arm_cfft_q31: in-place radix-4 DIF
 * (first / middle / last stage scaling), optional radix-2 pre-pass, binary
 * bit-reversal swap list.  All int32 arithmetic wraps (the reference is built
 * with -fwrapv); products are truncating or rounding high words exactly as in
 * the reference.
 *
 *   arm_cfft_q31                       Source/TransformFunctions/arm_cfft_q31.c:704-755
 *   arm_cfft_radix4by2_q31 / _inverse  arm_cfft_q31.c:763-822 / :824-881
 *   arm_radix4_butterfly_q31           Source/TransformFunctions/arm_cfft_radix4_q31.c:153-473
 *   arm_radix4_butterfly_inverse_q31   arm_cfft_radix4_q31.c:524-834
 *   mult_32x32_keep32_R & co           Include/dsp/none.h:185-194
 */
#include "orc_fft.h"

typedef struct { int32_t re, im; } cq31;

static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wshl1(int32_t a) { return (int32_t)((uint32_t)a << 1); }
/* truncating high word: (int32)(((q63)a*b) >> 32) */
static inline int32_t hi32(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 32); }
/* SMMULR / SMMLAR / SMMLSR emulation (none.h:185-194) */
static inline int32_t rhi32(int32_t x, int32_t y)
{
    return (int32_t)(((int64_t)x * y + 0x80000000LL) >> 32);
}
static inline int32_t rhi32_acc(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) + (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
static inline int32_t rhi32_sub(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) - (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}

enum { ST_FIRST, ST_MIDDLE };

/* (R,S) times conj(W) forward, times W inverse; truncating products. */
static inline cq31 rot(int32_t R, int32_t S, int32_t co, int32_t si, int inv)
{
    cq31 o;
    if (!inv) { o.re = wadd(hi32(R, co), hi32(S, si)); o.im = wsub(hi32(S, co), hi32(R, si)); }
    else      { o.re = wsub(hi32(R, co), hi32(S, si)); o.im = wadd(hi32(S, co), hi32(R, si)); }
    return o;
}

/* One first- or middle-stage butterfly on slots (i0, i0+n2, i0+2n2, i0+3n2).
 * ia = twiddle index of W^1 for this butterfly (W^2 at 2ia, W^3 at 3ia).
 * first : inputs >>4, a' unshifted, products <<1      (radix4_q31.c:199-274)
 * middle: inputs as stored, a' >>2, products >>1      (radix4_q31.c:329-391) */
static inline void bfly(cq31 *d, uint32_t i0, uint32_t n2, const int32_t *tw, uint32_t ia,
                        int stage, int inv)
{
    cq31 a = d[i0], b = d[i0 + n2], c = d[i0 + 2 * n2], e = d[i0 + 3 * n2];
    if (stage == ST_FIRST) {
        a.re >>= 4; a.im >>= 4; b.re >>= 4; b.im >>= 4;
        c.re >>= 4; c.im >>= 4; e.re >>= 4; e.im >>= 4;
    }
    int32_t r1 = wadd(a.re, c.re), r2 = wsub(a.re, c.re);
    int32_t s1 = wadd(a.im, c.im), s2 = wsub(a.im, c.im);
    int32_t t1 = wadd(b.re, e.re), t2 = wadd(b.im, e.im);
    int32_t u1 = wsub(b.im, e.im), u2 = wsub(b.re, e.re);
    cq31 oa = {wadd(r1, t1), wadd(s1, t2)};
    if (stage == ST_MIDDLE) { oa.re >>= 2; oa.im >>= 2; }

    cq31 oc = rot(wsub(r1, t1), wsub(s1, t2), tw[4 * ia], tw[4 * ia + 1], inv);
    cq31 ob, od;
    if (!inv) {
        ob = rot(wadd(r2, u1), wsub(s2, u2), tw[2 * ia], tw[2 * ia + 1], 0);
        od = rot(wsub(r2, u1), wadd(s2, u2), tw[6 * ia], tw[6 * ia + 1], 0);
    } else {
        ob = rot(wsub(r2, u1), wadd(s2, u2), tw[2 * ia], tw[2 * ia + 1], 1);
        od = rot(wadd(r2, u1), wsub(s2, u2), tw[6 * ia], tw[6 * ia + 1], 1);
    }
    if (stage == ST_FIRST) {
        oc.re = wshl1(oc.re); oc.im = wshl1(oc.im);
        ob.re = wshl1(ob.re); ob.im = wshl1(ob.im);
        od.re = wshl1(od.re); od.im = wshl1(od.im);
    } else {
        oc.re >>= 1; oc.im >>= 1; ob.re >>= 1; ob.im >>= 1; od.re >>= 1; od.im >>= 1;
    }
    /* slot order a', c', b', d' so that binary bit reversal gives natural order */
    d[i0] = oa; d[i0 + n2] = oc; d[i0 + 2 * n2] = ob; d[i0 + 3 * n2] = od;
}

static void radix4_passes(cq31 *d, uint32_t len, const int32_t *tw, uint32_t mod, int inv)
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
    /* last stage: groups of 4 consecutive points, no twiddles, no scaling (:411-464 / :768-831) */
    for (uint32_t g = 0; g < len; g += 4) {
        cq31 a = d[g], b = d[g + 1], c = d[g + 2], e = d[g + 3];
        cq31 oa = {wadd(wadd(a.re, b.re), wadd(c.re, e.re)), wadd(wadd(a.im, b.im), wadd(c.im, e.im))};
        cq31 oc = {wsub(wadd(wsub(a.re, b.re), c.re), e.re), wsub(wadd(wsub(a.im, b.im), c.im), e.im)};
        /* p = (xa + yb - xc - yd, ya - xb - yc + xd), q = (xa - yb - xc + yd, ya + xb - yc - xd) */
        cq31 p = {wsub(wsub(wadd(a.re, b.im), c.re), e.im), wadd(wsub(wsub(a.im, b.re), c.im), e.re)};
        cq31 q = {wadd(wsub(wsub(a.re, b.im), c.re), e.im), wsub(wsub(wadd(a.im, b.re), c.im), e.re)};
        d[g] = oa; d[g + 1] = oc;
        if (!inv) { d[g + 2] = p; d[g + 3] = q; }
        else      { d[g + 2] = q; d[g + 3] = p; }
    }
}

static void radix4by2(cq31 *d, uint32_t N, const int32_t *tw, int inv)
{
    uint32_t h = N >> 1;
    for (uint32_t i = 0; i < h; i++) {
        int32_t co = tw[2 * i], si = tw[2 * i + 1];
        cq31 a = d[i], b = d[i + h];
        int32_t xt = wsub(a.re >> 2, b.re >> 2), yt = wsub(a.im >> 2, b.im >> 2);
        d[i].re = wadd(a.re >> 2, b.re >> 2);
        d[i].im = wadd(b.im >> 2, a.im >> 2);
        int32_t p0 = rhi32(xt, co), p1 = rhi32(yt, co);
        if (!inv) { p0 = rhi32_acc(p0, yt, si); p1 = rhi32_sub(p1, xt, si); }
        else      { p0 = rhi32_sub(p0, yt, si); p1 = rhi32_acc(p1, xt, si); }
        d[i + h].re = wshl1(p0);
        d[i + h].im = wshl1(p1);
    }
    radix4_passes(d, h, tw, 2, inv);
    radix4_passes(d + h, h, tw, 2, inv);
    for (uint32_t i = 0; i < N; i++) { d[i].re = wshl1(d[i].re); d[i].im = wshl1(d[i].im); }
}

void orc_cfft_q31(uint32_t N, int32_t *p, int ifftFlag, int bitReverseFlag)
{
    const int32_t *tw = orc_twiddle_q31(N);
    if (!tw) return;
    cq31 *d = (cq31 *)p;
    int inv = (ifftFlag == 1);
    switch (N) {
    case 16: case 64: case 256: case 1024: case 4096: radix4_passes(d, N, tw, 1, inv); break;
    default:                                          radix4by2(d, N, tw, inv); break;
    }
    if (bitReverseFlag) {
        uint16_t len;
        const uint16_t *tab = orc_bitrev_fixed(N, &len);
        for (uint32_t i = 0; i < len; i += 2) {          /* arm_bitreversal_32, bitreversal2.c:84-108 */
            uint32_t a = tab[i] >> 3, b = tab[i + 1] >> 3;
            cq31 t = d[a]; d[a] = d[b]; d[b] = t;
        }
    }
}
