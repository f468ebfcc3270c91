/*
 * oracle/orc_cfft_f64.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restatement of the reference's arm_cfft_f64: in-place radix-4 decimation-in-frequency
 * stages without scaling, one radix-2 pre-pass for N = 2*4^m, binary bit reversal of the
 * result, and the inverse as conjugate -> forward -> conjugate / N.  Built with
 * -ffp-contract=off like the reference, and every sum and product is taken in the
 * reference's order, so with the SAME twiddle table the output is bit-identical
 * (tests/test_oracle_vs_ref.py feeds it the compiled reference's twiddleCoefF64_N).
 *
 *   arm_cfft_f64              Source/TransformFunctions/arm_cfft_f64.c:262-312
 *   arm_radix4_butterfly_f64  arm_cfft_f64.c:58-183
 *   arm_cfft_radix4by2_f64    arm_cfft_f64.c:193-239
 *   arm_bitreversal_64        Source/TransformFunctions/arm_bitreversal2.c:45-70
 *                             (armBitRevIndexTableF64_N == armBitRevIndexTable_fixed_N)
 *
 * Parity status of the TABLE: the reference's twiddleCoefF64_N literals
 * (Source/CommonTables/arm_common_tables.c:191-8506) are not reproducible by a rule; the
 * table generated here (first-quadrant sines in double precision, other quadrants by
 * symmetry) differs from them by 1 ulp in 15-20 % of the entries and never by more, which
 * tests/test_oracle_vs_ref.py checks.  f64 parity against the reference is therefore a
 * tolerance (relative RMS <= 1e-15), not bit-exactness.
 */
#include "orc_fft.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

typedef struct { double re, im; } cf64_t;

static double *g_tw[9];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static const uint32_t k_len[9] = {16, 32, 64, 128, 256, 512, 1024, 2048, 4096};

static void build_tables(void)
{
    const double two_pi = 6.283185307179586476925286766559;
    for (int li = 0; li < 9; li++) {
        const uint32_t N = k_len[li], q = N / 4;
        double *qs = malloc((q + 1) * sizeof *qs), *t = malloc(2u * N * sizeof *t);
        for (uint32_t i = 0; i <= q; i++) qs[i] = sin(two_pi * (double)i / (double)N);
        for (uint32_t i = 0; i < N; i++) {
            const uint32_t quad = i / q, r = i % q;
            const double s = qs[r], c = qs[q - r];
            double co, si;
            switch (quad) {
            case 0: co = c; si = s; break;
            case 1: co = -s; si = c; break;
            case 2: co = -c; si = -s; break;
            default: co = s; si = -c; break;
            }
            t[2 * i] = (co == 0.0) ? 0.0 : co;
            t[2 * i + 1] = (si == 0.0) ? 0.0 : si;
        }
        free(qs);
        g_tw[li] = t;
    }
}

const double *orc_twiddle_f64(uint32_t N)
{
    pthread_once(&g_once, build_tables);
    for (int li = 0; li < 9; li++)
        if (k_len[li] == N) return g_tw[li];
    return NULL;
}

/* (r, s) * conj(W), W = (co, +si)   (arm_cfft_f64.c:139-143) */
static inline cf64_t rot(double r, double s, double co, double si)
{
    cf64_t o = {(r * co) + (s * si), (s * co) - (r * si)};
    return o;
}

/* all radix-4 stages of an n-point transform on d; twiddle W^k of this transform = tw[k * step]
 * (arm_cfft_f64.c:76-182: n2 = quarter span, outputs a' -> i0, c' (W^2k) -> i0+n2, b' (W^k) -> i0+2n2,
 * d' (W^3k) -> i0+3n2) */
static void radix4_stages(cf64_t *d, uint32_t n, const cf64_t *tw, uint32_t step)
{
    for (uint32_t span = n; span > 1; span >>= 2, step <<= 2) {
        const uint32_t n2 = span >> 2;
        for (uint32_t j = 0; j < n2; j++) {
            const cf64_t w1 = tw[j * step], w2 = tw[2 * j * step], w3 = tw[3 * j * step];
            for (uint32_t i0 = j; i0 < n; i0 += span) {
                cf64_t *pa = d + i0, *pb = pa + n2, *pc = pb + n2, *pd = pc + n2;
                double r1 = pa->re + pc->re, r2 = pa->re - pc->re;
                double s1 = pa->im + pc->im, s2 = pa->im - pc->im;
                double t1 = pb->re + pd->re;
                const double ar = r1 + t1;
                r1 = r1 - t1;
                double t2 = pb->im + pd->im;
                const double ai = s1 + t2;
                s1 = s1 - t2;
                t1 = pb->im - pd->im;
                t2 = pb->re - pd->re;
                const cf64_t oc = rot(r1, s1, w2.re, w2.im);
                r1 = r2 + t1;
                r2 = r2 - t1;
                s1 = s2 - t2;
                s2 = s2 + t2;
                pa->re = ar; pa->im = ai;
                *pb = oc;
                *pc = rot(r1, s1, w1.re, w1.im);
                *pd = rot(r2, s2, w3.re, w3.im);
            }
        }
    }
}

static uint32_t bitrev(uint32_t k, uint32_t lg)
{
    uint32_t r = 0;
    for (uint32_t b = 0; b < lg; b++) r |= ((k >> b) & 1u) << (lg - 1u - b);
    return r;
}

void orc_cfft_f64(uint32_t N, double *p, int ifftFlag, int bitReverseFlag, const double *twiddle)
{
    const cf64_t *tw = (const cf64_t *)(twiddle ? twiddle : orc_twiddle_f64(N));
    cf64_t *d = (cf64_t *)p;
    if (!tw) return;                       /* unsupported length: no-op like the reference's switch (:272-291) */
    uint32_t lg = 0;
    while ((1u << lg) < N) lg++;
    if (ifftFlag == 1)
        for (uint32_t i = 0; i < N; i++) d[i].im = -d[i].im;
    if (lg & 1u) {
        const uint32_t h = N >> 1;                                      /* :205-230 */
        for (uint32_t i = 0; i < h; i++) {
            const double a0 = d[i].re + d[i + h].re, xt = d[i].re - d[i + h].re;
            const double yt = d[i].im - d[i + h].im, a1 = d[i + h].im + d[i].im;
            const double p0 = xt * tw[i].re, p1 = yt * tw[i].im, p2 = yt * tw[i].re, p3 = xt * tw[i].im;
            d[i].re = a0; d[i].im = a1;
            d[i + h].re = p0 + p1; d[i + h].im = p2 - p3;
        }
        radix4_stages(d, h, tw, 2);
        radix4_stages(d + h, h, tw, 2);
    } else {
        radix4_stages(d, N, tw, 1);
    }
    if (bitReverseFlag)
        for (uint32_t k = 0; k < N; k++) {
            const uint32_t r = bitrev(k, lg);
            if (r > k) { const cf64_t t = d[k]; d[k] = d[r]; d[r] = t; }
        }
    if (ifftFlag == 1) {
        const double invL = 1.0 / (double)N;
        for (uint32_t i = 0; i < N; i++) { d[i].re *= invL; d[i].im = -(d[i].im) * invL; }
    }
}

typedef struct { uint32_t N; double *p; uint64_t f0, f1; int ifft, bitrev; const double *tw; } job64_t;
static void *worker64(void *arg)
{
    job64_t *j = arg;
    for (uint64_t f = j->f0; f < j->f1; f++) orc_cfft_f64(j->N, j->p + 2ull * j->N * f, j->ifft, j->bitrev, j->tw);
    return NULL;
}
void orc_cfft_f64_batch(uint32_t N, double *p, uint64_t nFrames, int ifft, int bitrev, const double *twiddle, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nFrames) nthreads = nFrames ? (int)nFrames : 1;
    (void)orc_twiddle_f64(16);
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    job64_t *jobs = malloc((size_t)nthreads * sizeof *jobs);
    const uint64_t per = (nFrames + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t f0 = per * (uint64_t)t, f1 = f0 + per;
        if (f0 > nFrames) f0 = nFrames;
        if (f1 > nFrames) f1 = nFrames;
        jobs[t] = (job64_t){N, p, f0, f1, ifft, bitrev, twiddle};
        if (nthreads == 1) worker64(&jobs[t]);
        else pthread_create(&th[t], NULL, worker64, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------ arm_rfft_fast_f64
 *   arm_rfft_fast_f64   Source/TransformFunctions/arm_rfft_fast_f64.c:207-233
 *   stage_rfft_f64      arm_rfft_fast_f64.c:30-118     (split after the forward N/2-point CFFT, which runs in place on p)
 *   merge_rfft_f64      arm_rfft_fast_f64.c:121-181    (merge into pOut before the inverse CFFT, in place on pOut)
 * Real-stage table: (sin, cos)(2 pi k / N), k < N/2 -- the reference's twiddleCoefF64_rfft_N literals
 * (arm_common_tables.c:26703-30810) are within 1 ulp of the generated one, like the complex table. */
static double *g_twr[9];
static pthread_once_t g_once_r = PTHREAD_ONCE_INIT;
static void build_rfft_tables(void)
{
    const double two_pi = 6.283185307179586476925286766559;
    for (int li = 1; li < 9; li++) {
        const uint32_t N = k_len[li], q = N / 4;
        double *qs = malloc((q + 1) * sizeof *qs), *t = malloc(N * sizeof *t);
        for (uint32_t i = 0; i <= q; i++) qs[i] = sin(two_pi * (double)i / (double)N);
        for (uint32_t i = 0; i < N / 2; i++) {
            const double s = (i < q) ? qs[i] : qs[2 * q - i], c = (i < q) ? qs[q - i] : -qs[i - q];
            t[2 * i] = (s == 0.0) ? 0.0 : s;
            t[2 * i + 1] = (c == 0.0) ? 0.0 : c;
        }
        free(qs);
        g_twr[li] = t;
    }
}
const double *orc_twiddle_rfft_f64(uint32_t N)
{
    pthread_once(&g_once_r, build_rfft_tables);
    for (int li = 1; li < 9; li++)
        if (k_len[li] == N) return g_twr[li];
    return NULL;
}

void orc_rfft_fast_f64(uint32_t N, double *p, double *pOut, int ifftFlag, const double *twiddleC, const double *twiddleR)
{
    const uint32_t L = N / 2;
    const cf64_t *tw = (const cf64_t *)(twiddleR ? twiddleR : orc_twiddle_rfft_f64(N));
    if (!tw) return;
    if (ifftFlag) {
        const cf64_t *x = (const cf64_t *)p;
        cf64_t *y = (cf64_t *)pOut;
        y[0].re = 0.5 * (x[0].re + x[0].im);
        y[0].im = 0.5 * (x[0].re - x[0].im);
        for (uint32_t k = 1; k < L; k++) {
            const cf64_t a = x[k], b = x[L - k];
            const double t1a = a.re - b.re, t1b = a.im + b.im;
            const double r = tw[k].re * t1a, s = tw[k].im * t1b, t = tw[k].im * t1a, u = tw[k].re * t1b;
            y[k].re = 0.5 * (a.re + b.re - r - s);
            y[k].im = 0.5 * (a.im - b.im + t - u);
        }
        orc_cfft_f64(L, pOut, 1, 1, twiddleC);
        return;
    }
    orc_cfft_f64(L, p, 0, 1, twiddleC);
    const cf64_t *x = (const cf64_t *)p;
    cf64_t *y = (cf64_t *)pOut;
    {
        const double t1a = x[0].re + x[0].re, t1b = x[0].im + x[0].im;
        y[0].re = 0.5 * (t1a + t1b);
        y[0].im = 0.5 * (t1a - t1b);
    }
    for (uint32_t k = 1; k < L; k++) {
        const cf64_t a = x[k], b = x[L - k];
        const double t1a = b.re - a.re, t1b = b.im + a.im;
        const double p0 = tw[k].re * t1a, p1 = tw[k].im * t1a, p2 = tw[k].re * t1b, p3 = tw[k].im * t1b;
        y[k].re = 0.5 * (a.re + b.re + p0 + p3);
        y[k].im = 0.5 * (a.im - b.im + p1 - p2);
    }
}

/* p is copied per frame (the forward transform destroys it); single thread: parity only */
void orc_rfft_fast_f64_batch(uint32_t N, const double *p, double *pOut, uint64_t nFrames, int ifft, const double *twiddleC,
                             const double *twiddleR)
{
    double *tmp = malloc(N * sizeof *tmp);
    for (uint64_t f = 0; f < nFrames; f++) {
        for (uint32_t i = 0; i < N; i++) tmp[i] = p[f * N + i];
        orc_rfft_fast_f64(N, tmp, pOut + f * N, ifft, twiddleC, twiddleR);
    }
    free(tmp);
}
