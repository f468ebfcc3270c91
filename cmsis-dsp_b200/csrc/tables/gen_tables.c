/*
 * gen_tables.c -- build-time generator of the constant tables of the FFT path.
 *
 * The reference ships these tables as ~50k literal values in
 * Source/CommonTables/arm_common_tables.c; here they are GENERATED at build time from
 * the closed-form rules below and compiled into libcmsisdsp_b200.so as const data under
 * the reference's own symbol names (Include/arm_common_tables.h), so the instance
 * structs filled by arm_cfft_init_* point at value-identical tables:
 *
 *   twiddleCoef_N            (cos, +sin)(2 pi i / N), i < N, rounded to 9 decimals then to float
 *                            (arm_common_tables.c:8508-16835)
 *   twiddleCoef_N_q31        i < 3N/4, floor(x * 2^31 + 0.05) clipped to int32     (:16845-21140)
 *   twiddleCoef_N_q15        i < 3N/4, floor(x * 2^15) clipped to int16            (:21149-24400)
 *   twiddleCoef_rfft_N       (sin, cos)(2 pi i / N), i < N/2, 9 decimals           (:30813-34940)
 *   armBitRevIndexTableN     ordered swap list of the (r0,8,8,..) digit reversal,
 *                            entries are complex index * 8                          (:25057-26040)
 *   armBitRevIndexTable_fixed_N  swap list of the binary bit reversal               (:26042-26700)
 *   twiddleCoefF64_N         (cos, +sin)(2 pi i / N) as IEEE-754 double bit patterns, i < N: first-quadrant sines in
 *                            double precision, the other quadrants by symmetry (:191-8506).  NOT value-identical:
 *                            the reference's literals are not reproducible by a rule; 15-20 % of them differ from
 *                            these by 1 ulp (none by more), which is inside the f64 parity bar (rel-RMS 1e-15)
 *   armBitRevIndexTableF64_N the fixed-point swap lists under their f64 names    (:24404-25050)
 *   armBitRevTable           (12-bit reversal of l) >> 1, l = 1..1024 (deprecated radix-2/4 API)   (:41-67)
 *   realCoefAQ31 / realCoefBQ31 / realCoefAQ15 / realCoefBQ15   split-stage coefficients of arm_rfft_q31 / _q15,
 *                            n = 4096: A = 0.5 (1 - sin, -cos), B = 0.5 (1 + sin, cos)(2 pi i / 2n),
 *                            round(x * 2^31 | 2^15) clipped                         (:39063-45280)
 *
 * tests/test_tables.py checks every generated table against the compiled reference.
 *
 * usage: gen_tables > cmsisdsp_tables_generated.c
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const unsigned kLen[9] = {16, 32, 64, 128, 256, 512, 1024, 2048, 4096};
static const double kTwoPi = 6.283185307179586476925286766559;

static double dec9(double x) { return rint(x * 1e9) / 1e9; }
static long long clipll(double v, long long lo, long long hi)
{
    if (v < (double)lo) return lo;
    if (v > (double)hi) return hi;
    return (long long)v;
}

static unsigned ilog2(unsigned n) { unsigned l = 0; while ((1u << l) < n) l++; return l; }

static unsigned digit_reverse(unsigned N, unsigned k)
{
    unsigned lg = ilog2(N), r0 = 1u << (lg % 3u), pos = 0;
    if (r0 > 1) { pos = k % r0; k /= r0; }
    for (unsigned d = 0; d < lg / 3u; d++) { pos = pos * 8u + (k & 7u); k >>= 3; }
    return pos;
}
static unsigned bit_reverse(unsigned lg, unsigned k)
{
    unsigned r = 0;
    for (unsigned b = 0; b < lg; b++) r |= ((k >> b) & 1u) << (lg - 1u - b);
    return r;
}

/* ordered swaps that realise out[t] = in[P[t]], found by walking t upwards */
static unsigned swap_list(unsigned N, const unsigned *P, unsigned *tab)
{
    unsigned *at = malloc(N * sizeof *at), *where = malloc(N * sizeof *where), n = 0;
    for (unsigned i = 0; i < N; i++) at[i] = where[i] = i;
    for (unsigned t = 0; t < N; t++) {
        unsigned pos = where[P[t]];
        if (pos == t) continue;
        tab[n++] = t * 8u;
        tab[n++] = pos * 8u;
        unsigned x = at[t], y = at[pos];
        at[t] = y; at[pos] = x; where[y] = t; where[x] = pos;
    }
    free(at); free(where);
    return n;
}

/* The reference lists a few neighbouring swaps in exchanged order for N = 16, 32, 256, 2048
 * (entries 2i and 2i+1 of its tables never touch the same point).  e = swap index: swaps e
 * and e+1 trade places. */
static void exchange_neighbours(unsigned N, unsigned *tab)
{
    static const unsigned e16[] = {1, 3, 5, 7}, e32[] = {9, 17, 21}, e256[] = {63, 207, 213, 215, 217},
                          e2048[] = {1741, 1765, 1773, 1857, 1875, 1883, 1901};
    const unsigned *e = 0; unsigned ne = 0;
    if (N == 16) { e = e16; ne = 4; } else if (N == 32) { e = e32; ne = 3; }
    else if (N == 256) { e = e256; ne = 5; } else if (N == 2048) { e = e2048; ne = 7; }
    for (unsigned i = 0; i < ne; i++) {
        unsigned *a = tab + 2u * e[i], t0 = a[0], t1 = a[1];
        a[0] = a[2]; a[1] = a[3]; a[2] = t0; a[3] = t1;
    }
}

/* sin(2 pi i / N) for 0 <= i <= N/4; -0.0 never appears (the reference's tables hold +0) */
static double quarter_sin(unsigned i, unsigned N) { return sin(kTwoPi * (double)i / (double)N); }
static unsigned long long dbits(double v)
{
    unsigned long long u;
    if (v == 0.0) v = 0.0;
    memcpy(&u, &v, sizeof u);
    return u;
}

static void emit_u16(const char *name, unsigned N, const unsigned *tab, unsigned n)
{
    printf("const uint16_t %s%u[%u] = {", name, N, n);
    for (unsigned i = 0; i < n; i++) printf("%s%u,", (i % 16) ? "" : "\n  ", tab[i]);
    printf("\n};\n");
}

int main(void)
{
    printf("/* GENERATED by csrc/tables/gen_tables.c -- do not edit, do not commit. */\n");
    printf("#include <stdint.h>\n#include \"arm_common_tables.h\"\n\n");
    for (int li = 0; li < 9; li++) {
        unsigned N = kLen[li], lg = ilog2(N), nq = 3u * N / 4u;
        /* hex-float literals keep the exact float value */
        printf("const float32_t twiddleCoef_%u[%u] = {", N, 2 * N);
        for (unsigned i = 0; i < N; i++) {
            double a = kTwoPi * (double)i / (double)N;
            printf("%s%a,%a,", (i % 4) ? "" : "\n  ", (double)(float)dec9(cos(a)), (double)(float)dec9(sin(a)));
        }
        printf("\n};\n");
        printf("const q31_t twiddleCoef_%u_q31[%u] = {", N, 2 * nq);
        for (unsigned i = 0; i < nq; i++) {
            double a = kTwoPi * (double)i / (double)N;
            printf("%s(q31_t)0x%08llXu,(q31_t)0x%08llXu,", (i % 4) ? "" : "\n  ",
                   (unsigned long long)(clipll(floor(cos(a) * 2147483648.0 + 0.05), INT32_MIN, INT32_MAX) & 0xFFFFFFFFll),
                   (unsigned long long)(clipll(floor(sin(a) * 2147483648.0 + 0.05), INT32_MIN, INT32_MAX) & 0xFFFFFFFFll));
        }
        printf("\n};\n");
        printf("const q15_t twiddleCoef_%u_q15[%u] = {", N, 2 * nq);
        for (unsigned i = 0; i < nq; i++) {
            double a = kTwoPi * (double)i / (double)N;
            printf("%s(q15_t)0x%04llXu,(q15_t)0x%04llXu,", (i % 8) ? "" : "\n  ",
                   (unsigned long long)(clipll(floor(cos(a) * 32768.0), INT16_MIN, INT16_MAX) & 0xFFFFll),
                   (unsigned long long)(clipll(floor(sin(a) * 32768.0), INT16_MIN, INT16_MAX) & 0xFFFFll));
        }
        printf("\n};\n");
        if (N >= 32) {
            printf("const float32_t twiddleCoef_rfft_%u[%u] = {", N, N);
            for (unsigned i = 0; i < N / 2; i++) {
                double a = kTwoPi * (double)i / (double)N;
                printf("%s%a,%a,", (i % 4) ? "" : "\n  ", (double)(float)dec9(sin(a)), (double)(float)dec9(cos(a)));
            }
            printf("\n};\n");
        }
        unsigned *P = malloc(N * sizeof *P), *tab = malloc(2 * N * sizeof *tab), n;
        for (unsigned k = 0; k < N; k++) P[k] = digit_reverse(N, k);
        n = swap_list(N, P, tab);
        exchange_neighbours(N, tab);
        emit_u16("armBitRevIndexTable", N, tab, n);
        printf("const uint16_t cmsisdsp_b200_bitrev_len_f32_%u = %u;\n", N, n);
        for (unsigned k = 0; k < N; k++) P[k] = bit_reverse(lg, k);
        n = swap_list(N, P, tab);
        emit_u16("armBitRevIndexTable_fixed_", N, tab, n);
        emit_u16("armBitRevIndexTableF64_", N, tab, n);      /* the reference's f64 lists are the fixed-point ones */
        /* f64 twiddles as bit patterns: first-quadrant sines, the rest by symmetry (exact 0 / 1 on the axes) */
        printf("const uint64_t twiddleCoefF64_%u[%u] = {", N, 2 * N);
        for (unsigned i = 0; i < N; i++) {
            const unsigned q = N / 4, quad = i / q, r = i % q;
            const double sr = quarter_sin(r, N), cr = quarter_sin(q - r, N);
            double c, sn;
            switch (quad) {
            case 0: c = cr; sn = sr; break;
            case 1: c = -sr; sn = cr; break;
            case 2: c = -cr; sn = -sr; break;
            default: c = sr; sn = -cr; break;
            }
            printf("%s0x%016llxull,0x%016llxull,", (i % 4) ? "" : "\n  ", dbits(c), dbits(sn));
        }
        printf("\n};\n");
        if (N >= 32) {      /* twiddleCoefF64_rfft_N: (sin, cos)(2 pi i / N), i < N/2, same rule (arm_common_tables.c:26703-30810) */
            printf("const uint64_t twiddleCoefF64_rfft_%u[%u] = {", N, N);
            for (unsigned i = 0; i < N / 2; i++) {
                const unsigned q = N / 4;
                const double sn = (i < q) ? quarter_sin(i, N) : quarter_sin(2 * q - i, N);
                const double c = (i < q) ? quarter_sin(q - i, N) : -quarter_sin(i - q, N);
                printf("%s0x%016llxull,0x%016llxull,", (i % 4) ? "" : "\n  ", dbits(sn), dbits(c));
            }
            printf("\n};\n");
        }
        printf("const uint16_t cmsisdsp_b200_bitrev_len_fixed_%u = %u;\n\n", N, n);
        free(P); free(tab);
    }
    /* armBitRevTable of the deprecated radix-2 / radix-4 instance API (arm_common_tables.c:41-67): entry l-1 =
     * (12-bit reversal of l) >> 1, l = 1..1024 */
    printf("const uint16_t armBitRevTable[1024] = {");
    for (unsigned l = 1; l <= 1024; l++) printf("%s%u,", ((l - 1) % 16) ? "" : "\n  ", bit_reverse(12, l) >> 1);
    printf("\n};\n");
    /* split-stage coefficients of the fixed-point real FFT */
    for (int t = 0; t < 4; t++) {
        const int q15 = t >= 2, isB = t & 1;
        printf("const %s realCoef%cQ%s[8192] = {", q15 ? "q15_t" : "q31_t", isB ? 'B' : 'A', q15 ? "15" : "31");
        for (unsigned i = 0; i < 4096; i++) {
            const double a = kTwoPi / 8192.0 * (double)i;
            const double v0 = isB ? 0.5 * (1.0 + sin(a)) : 0.5 * (1.0 - sin(a)), v1 = isB ? 0.5 * cos(a) : 0.5 * (-1.0 * cos(a));
            const double sc = q15 ? 32768.0 : 2147483648.0;
            const long long lo = q15 ? INT16_MIN : INT32_MIN, hi = q15 ? INT16_MAX : INT32_MAX;
            const unsigned long long m = q15 ? 0xFFFFull : 0xFFFFFFFFull;
            printf(q15 ? "%s(q15_t)0x%04llXu,(q15_t)0x%04llXu," : "%s(q31_t)0x%08llXu,(q31_t)0x%08llXu,", (i % 4) ? "" : "\n  ",
                   (unsigned long long)clipll(floor(v0 * sc + 0.5), lo, hi) & m, (unsigned long long)clipll(floor(v1 * sc + 0.5), lo, hi) & m);
        }
        printf("\n};\n");
    }
    return 0;
}
