/*
 * oracle/orc_tables.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 *
 * Restated generation rules for the reference's constant tables
 * (Source/CommonTables/arm_common_tables.c).  The reference ships the tables
 * as literal data; the rules below were recovered by probing and are checked
 * entry-for-entry against the compiled reference in tests/test_oracle_vs_ref.py:
 *
 *  twiddleCoef_N          (:8523-16835)  (cos,+sin)(2*pi*i/N), i<N, written with
 *                                         9 decimals => float(rint(x*1e9)/1e9)
 *  twiddleCoef_N_q31      (:16860-21140) i<3N/4, floor(x*2^31 + 0.05) clipped to int32
 *  twiddleCoef_N_q15      (:21149-24400) i<3N/4, floor(x*2^15) clipped (== q31 >> 16)
 *  twiddleCoef_rfft_N     (:30820-34940) (sin,cos)(2*pi*i/N), i<N/2, 9 decimals
 *  armBitRevIndexTableN   (:25057-26040) ordered swap list realising the
 *                                         mixed-radix (r0,8,8,..) digit reversal,
 *                                         entries = complex index * 8
 *  armBitRevIndexTable_fixed_N (:26042-26700) swap list of the plain binary bit
 *                                         reversal, pairs (i,rev(i)), i<rev(i), ascending i
 */
#include "orc_fft.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NLEN 9
static const uint32_t k_len[ORC_NLEN] = {16, 32, 64, 128, 256, 512, 1024, 2048, 4096};

static float    *g_tw_f32[ORC_NLEN];
static int32_t  *g_tw_q31[ORC_NLEN];
static int16_t  *g_tw_q15[ORC_NLEN];
static float    *g_tw_rfft[ORC_NLEN];   /* indexed by real length N (32..4096); slot 0 unused */
static uint16_t *g_br_f32[ORC_NLEN];
static uint16_t  g_br_f32_len[ORC_NLEN];
static uint16_t *g_br_fix[ORC_NLEN];
static uint16_t  g_br_fix_len[ORC_NLEN];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static int len_index(uint32_t N)
{
    for (int i = 0; i < ORC_NLEN; i++)
        if (k_len[i] == N) return i;
    return -1;
}

static float dec9(double x) { return (float)(rint(x * 1e9) / 1e9); }

static int64_t clip64(double v, int64_t lo, int64_t hi)
{
    if (v < (double)lo) return lo;
    if (v > (double)hi) return hi;
    return (int64_t)v;
}

/* Mixed-radix digit reversal of the f32 CFFT: digits of k, least significant
 * first, have radices (r0, 8, 8, ...) with r0 = N / 8^floor(log8 N) in {1,2,4};
 * the scrambled position reads the same digits most significant first.
 * (arm_cfft_f32.c:1263-1280 picks radix8by2 / radix8by4 / radix8 accordingly.) */
static uint32_t digit_reverse_f32(uint32_t N, uint32_t k)
{
    uint32_t lg = 0;
    while ((1u << lg) < N) lg++;
    uint32_t r0 = 1u << (lg % 3u);
    uint32_t pos = 0;
    if (r0 > 1) { pos = k % r0; k /= r0; }
    for (uint32_t d = 0; d < lg / 3u; d++) { pos = pos * 8u + (k & 7u); k >>= 3; }
    return pos;
}

static uint32_t bit_reverse(uint32_t lg, uint32_t k)
{
    uint32_t r = 0;
    for (uint32_t b = 0; b < lg; b++) r |= ((k >> b) & 1u) << (lg - 1u - b);
    return r;
}

/* Decompose "out[t] = in[P[t]]" into an ordered list of swaps by walking t
 * upwards and pulling the wanted element into place. */
static uint16_t *swap_list(uint32_t N, const uint32_t *P, uint16_t *outLen)
{
    uint32_t *at = malloc(N * sizeof *at), *where = malloc(N * sizeof *where);
    uint16_t *tab = malloc(2u * N * sizeof *tab);
    uint32_t n = 0;
    for (uint32_t i = 0; i < N; i++) at[i] = where[i] = i;
    for (uint32_t t = 0; t < N; t++) {
        uint32_t pos = where[P[t]];
        if (pos == t) continue;
        tab[n++] = (uint16_t)(t * 8u);
        tab[n++] = (uint16_t)(pos * 8u);
        uint32_t x = at[t], y = at[pos];
        at[t] = y; at[pos] = x; where[y] = t; where[x] = pos;
    }
    free(at); free(where);
    *outLen = (uint16_t)n;
    return tab;
}

/* The reference tables for N=16,32,256,2048 list the same swaps with a few
 * neighbouring entries exchanged (entries 2i and 2i+1 never share an index,
 * which its two-swaps-per-iteration assembly needs).  Positions are swap
 * indices e: entries e and e+1 are exchanged. */
static void apply_exchanges(uint32_t N, uint16_t *tab)
{
    static const uint16_t e16[]   = {1, 3, 5, 7};
    static const uint16_t e32[]   = {9, 17, 21};
    static const uint16_t e256[]  = {63, 207, 213, 215, 217};
    static const uint16_t e2048[] = {1741, 1765, 1773, 1857, 1875, 1883, 1901};
    const uint16_t *e = NULL; uint32_t ne = 0;
    switch (N) {
    case 16:   e = e16;   ne = 4; break;
    case 32:   e = e32;   ne = 3; break;
    case 256:  e = e256;  ne = 5; break;
    case 2048: e = e2048; ne = 7; break;
    default: return;
    }
    for (uint32_t i = 0; i < ne; i++) {
        uint16_t *a = tab + 2u * e[i], *b = a + 2;
        uint16_t t0 = a[0], t1 = a[1];
        a[0] = b[0]; a[1] = b[1]; b[0] = t0; b[1] = t1;
    }
}

static void build_all(void)
{
    const double two_pi = 6.283185307179586476925286766559;
    for (int li = 0; li < ORC_NLEN; li++) {
        uint32_t N = k_len[li];
        uint32_t lg = 0;
        while ((1u << lg) < N) lg++;

        g_tw_f32[li] = malloc(2u * N * sizeof(float));
        for (uint32_t i = 0; i < N; i++) {
            double a = two_pi * (double)i / (double)N;
            g_tw_f32[li][2 * i]     = dec9(cos(a));
            g_tw_f32[li][2 * i + 1] = dec9(sin(a));
        }
        uint32_t nq = 3u * N / 4u;
        g_tw_q31[li] = malloc(2u * nq * sizeof(int32_t));
        g_tw_q15[li] = malloc(2u * nq * sizeof(int16_t));
        for (uint32_t i = 0; i < nq; i++) {
            double a = two_pi * (double)i / (double)N;
            double cs[2] = {cos(a), sin(a)};
            for (int c = 0; c < 2; c++) {
                g_tw_q31[li][2 * i + c] =
                    (int32_t)clip64(floor(cs[c] * 2147483648.0 + 0.05), INT32_MIN, INT32_MAX);
                g_tw_q15[li][2 * i + c] =
                    (int16_t)clip64(floor(cs[c] * 32768.0), INT16_MIN, INT16_MAX);
            }
        }
        if (N >= 32) {
            g_tw_rfft[li] = malloc(N * sizeof(float));
            for (uint32_t i = 0; i < N / 2u; i++) {
                double a = two_pi * (double)i / (double)N;
                g_tw_rfft[li][2 * i]     = dec9(sin(a));
                g_tw_rfft[li][2 * i + 1] = dec9(cos(a));
            }
        }
        uint32_t *P = malloc(N * sizeof *P);
        for (uint32_t k = 0; k < N; k++) P[k] = digit_reverse_f32(N, k);
        g_br_f32[li] = swap_list(N, P, &g_br_f32_len[li]);
        apply_exchanges(N, g_br_f32[li]);
        for (uint32_t k = 0; k < N; k++) P[k] = bit_reverse(lg, k);
        g_br_fix[li] = swap_list(N, P, &g_br_fix_len[li]);
        free(P);
    }
}

static int ready_index(uint32_t N)
{
    pthread_once(&g_once, build_all);
    return len_index(N);
}

const float *orc_twiddle_f32(uint32_t N)   { int i = ready_index(N); return i < 0 ? NULL : g_tw_f32[i]; }
const int32_t *orc_twiddle_q31(uint32_t N) { int i = ready_index(N); return i < 0 ? NULL : g_tw_q31[i]; }
const int16_t *orc_twiddle_q15(uint32_t N) { int i = ready_index(N); return i < 0 ? NULL : g_tw_q15[i]; }
const float *orc_twiddle_rfft_f32(uint32_t N)
{
    int i = ready_index(N);
    return (i < 1) ? NULL : g_tw_rfft[i];
}
const uint16_t *orc_bitrev_f32(uint32_t N, uint16_t *len)
{
    int i = ready_index(N);
    if (i < 0) return NULL;
    if (len) *len = g_br_f32_len[i];
    return g_br_f32[i];
}
const uint16_t *orc_bitrev_fixed(uint32_t N, uint16_t *len)
{
    int i = ready_index(N);
    if (i < 0) return NULL;
    if (len) *len = g_br_fix_len[i];
    return g_br_fix[i];
}

uint64_t orc_fnv1a64(const void *data, uint64_t nbytes)
{
    const uint8_t *p = data;
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < nbytes; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
