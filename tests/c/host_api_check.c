/*
 * host_api_check.c -- a plain C caller of the drop-in library, built with the sanitizers.
 *
 * What a CMSIS-DSP user's translation unit looks like after switching: #include "arm_math.h", the reference's init
 * functions and instance structs, the batched entry points.  tests/test_host_sanitizers.py compiles THIS file together
 * with the library's own host sources (the .c files of csrc/host) under -fsanitize=address,undefined and under -fsanitize=thread and
 * runs it
 *   - without a GPU (the build container): every batched call must answer ARM_MATH_CUDA_NO_DEVICE, leave its buffers
 *     untouched and leak / corrupt nothing -- there is no CPU fallback;
 *   - on the B200 box (-m gpu): the dispatcher (arm_cuda_engine.c) is driven with one device, with a repeated ordinal
 *     (three workers, three stream sets, one GPU), from four host threads at once, with small staging chunks; every
 *     variant must give the SAME BITS as the plain one-device call, the f32 round trips must close.
 * Parity against the oracle is the business of tests/test_gpu_*.py; this program is about the host side's memory and
 * thread safety.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arm_math.h"

static int g_fail;
#define CHECK(cond, ...)                                   \
    do {                                                   \
        if (!(cond)) {                                     \
            fprintf(stderr, "host_api_check: " __VA_ARGS__); \
            fprintf(stderr, "  [%s:%d]\n", __FILE__, __LINE__); \
            g_fail++;                                      \
        }                                                  \
    } while (0)

static uint32_t lcg(uint32_t *s) { return *s = *s * 1664525u + 1013904223u; }

static void fill_f32(float *p, size_t n, uint32_t seed)
{
    for (size_t k = 0; k < n; k++) p[k] = (float)((int32_t)lcg(&seed) >> 8) * (1.0f / 8388608.0f);
}
static void fill_q31(q31_t *p, size_t n, uint32_t seed)
{
    for (size_t k = 0; k < n; k++) p[k] = (q31_t)lcg(&seed);
}
static void fill_q15(q15_t *p, size_t n, uint32_t seed)
{
    for (size_t k = 0; k < n; k++) p[k] = (q15_t)(lcg(&seed) >> 16);
}

#define RN 1024u          /* arm_rfft_fast_f32 length */
#define RB 8192u          /* 32 MiB per direction: enough for three workers (8 MiB per worker minimum) */
#define QN 1024u          /* arm_cfft_q15 length */
#define QB 16384u         /* 64 MiB */
#define LN 512u           /* arm_cfft_q31 / arm_rfft_q31 length */
#define LB 8192u

static int g_gpu;          /* 1: a device answered; 0: ARM_MATH_CUDA_NO_DEVICE everywhere */

static arm_status expect(arm_status st, const char *what)
{
    if (g_gpu) CHECK(st == ARM_MATH_SUCCESS, "%s: status %d", what, (int)st);
    else       CHECK(st == ARM_MATH_CUDA_NO_DEVICE, "%s: status %d without a device (want ARM_MATH_CUDA_NO_DEVICE)", what, (int)st);
    return st;
}

/* one thread's private work: forward + inverse real FFT of its own frames, spectrum compared with the reference bits */
struct job {
    const arm_rfft_fast_instance_f32 *S;
    const float *in;
    const float *wantSpec;
    uint32_t frames;
    int id;
};
static void *thread_main(void *arg)
{
    struct job *j = (struct job *)arg;
    const size_t n = (size_t)j->frames * RN;
    float *in = malloc(n * sizeof(float)), *spec = malloc(n * sizeof(float)), *back = malloc(n * sizeof(float));
    memcpy(in, j->in, n * sizeof(float));
    expect(arm_rfft_fast_batch_f32(j->S, in, spec, j->frames, 0), "thread forward");
    expect(arm_rfft_fast_batch_f32(j->S, spec, back, j->frames, 1), "thread inverse");
    CHECK(memcmp(in, j->in, n * sizeof(float)) == 0, "thread %d: the batched call changed its input", j->id);
    if (g_gpu) {
        CHECK(memcmp(spec, j->wantSpec, n * sizeof(float)) == 0, "thread %d: spectrum differs from the main thread's", j->id);
        double e = 0.0, s = 0.0;
        for (size_t k = 0; k < n; k++) { e += (double)(back[k] - in[k]) * (back[k] - in[k]); s += (double)in[k] * in[k]; }
        CHECK(sqrt(e / s) < 1e-5, "thread %d: round trip %.3g", j->id, sqrt(e / s));
    }
    free(in); free(spec); free(back);
    arm_cuda_release();                      /* also runs from the thread-exit destructor; calling it twice must be harmless */
    return NULL;
}

int main(void)
{
    arm_rfft_fast_instance_f32 rf;
    arm_cfft_instance_f32 cf;
    arm_cfft_instance_q31 c31;
    arm_cfft_instance_q15 c15;
    arm_rfft_instance_q31 r31;
    CHECK(arm_rfft_fast_init_f32(&rf, RN) == ARM_MATH_SUCCESS, "arm_rfft_fast_init_f32");
    CHECK(arm_cfft_init_f32(&cf, 256) == ARM_MATH_SUCCESS, "arm_cfft_init_f32");
    CHECK(arm_cfft_init_q31(&c31, LN) == ARM_MATH_SUCCESS, "arm_cfft_init_q31");
    CHECK(arm_cfft_init_q15(&c15, QN) == ARM_MATH_SUCCESS, "arm_cfft_init_q15");
    CHECK(arm_rfft_init_q31(&r31, LN, 0, 1) == ARM_MATH_SUCCESS, "arm_rfft_init_q31");
    CHECK(arm_cfft_init_f32(&cf, 100) == ARM_MATH_ARGUMENT_ERROR, "arm_cfft_init_f32 accepted length 100");

    /* ---- probe: is there a device? ---- */
    const size_t rn = (size_t)RB * RN;
    float *rin = malloc(rn * sizeof(float)), *rkeep = malloc(rn * sizeof(float)), *spec0 = malloc(rn * sizeof(float)), *spec = malloc(rn * sizeof(float)),
          *back = malloc(rn * sizeof(float));
    fill_f32(rin, rn, 1u);
    memcpy(rkeep, rin, rn * sizeof(float));
    memset(spec0, 0x5a, rn * sizeof(float));
    const int32_t one[1] = {0}, three[3] = {0, 0, 0};
    arm_status st = arm_cuda_set_devices(one, 1);
    CHECK(st == ARM_MATH_SUCCESS || st == ARM_MATH_CUDA_NO_DEVICE, "arm_cuda_set_devices([0]): %d", (int)st);
    st = arm_rfft_fast_batch_f32(&rf, rin, spec0, RB, 0);
    g_gpu = (st == ARM_MATH_SUCCESS);
    CHECK(st == ARM_MATH_SUCCESS || st == ARM_MATH_CUDA_NO_DEVICE, "first batched call: status %d", (int)st);
    CHECK(memcmp(rin, rkeep, rn * sizeof(float)) == 0, "arm_rfft_fast_batch_f32 changed its input");
    if (!g_gpu) {
        const unsigned char *b = (const unsigned char *)spec0;
        size_t touched = 0;
        for (size_t k = 0; k < rn * sizeof(float); k++) touched += (b[k] != 0x5a);
        CHECK(touched == 0, "no device, yet %zu output bytes were written (a CPU fallback?)", touched);
    }

    /* ---- f32 real FFT: [0] vs [0, 0, 0] vs small staging chunks; round trip ---- */
    expect(arm_rfft_fast_batch_f32(&rf, spec0, back, RB, 1), "inverse, one device");
    if (g_gpu) {
        double e = 0.0, s = 0.0;
        for (size_t k = 0; k < rn; k++) { e += (double)(back[k] - rin[k]) * (back[k] - rin[k]); s += (double)rin[k] * rin[k]; }
        CHECK(sqrt(e / s) < 1e-5, "arm_rfft_fast_f32 round trip %.3g", sqrt(e / s));
    }
    arm_cuda_set_devices(three, 3);
    {
        int32_t got[8];
        const uint32_t n = arm_cuda_get_devices(got, 8);
        CHECK(!g_gpu || (n == 3 && got[0] == 0 && got[2] == 0), "arm_cuda_get_devices after set_devices([0,0,0]): %u", n);
    }
    expect(arm_rfft_fast_batch_f32(&rf, rin, spec, RB, 0), "forward, three workers on one device");
    CHECK(!g_gpu || memcmp(spec, spec0, rn * sizeof(float)) == 0, "three workers: spectrum differs from the one-device call");
    expect(arm_cuda_set_staging(4, 2) == ARM_MATH_SUCCESS ? arm_rfft_fast_batch_f32(&rf, rin, spec, RB, 0) : ARM_MATH_ARGUMENT_ERROR,
           "forward, 4 MiB chunks on 2 streams");
    CHECK(!g_gpu || memcmp(spec, spec0, rn * sizeof(float)) == 0, "4 MiB x 2 staging: spectrum differs");
    expect(arm_cuda_set_staging_ramp(1) == ARM_MATH_SUCCESS ? arm_rfft_fast_batch_f32(&rf, rin, spec, RB, 0) : ARM_MATH_ARGUMENT_ERROR,
           "forward, 4 MiB chunks on 2 streams, chunk sizes ramping from 1 MiB");
    CHECK(!g_gpu || memcmp(spec, spec0, rn * sizeof(float)) == 0, "ramped staging: spectrum differs");
    expect(arm_cuda_set_staging_ramp(0) == ARM_MATH_SUCCESS ? arm_rfft_fast_batch_f32(&rf, rin, spec, RB, 0) : ARM_MATH_ARGUMENT_ERROR,
           "forward, no ramp");
    CHECK(!g_gpu || memcmp(spec, spec0, rn * sizeof(float)) == 0, "unramped staging: spectrum differs");
    arm_cuda_set_staging(32, 3);
    arm_cuda_set_staging_ramp(4);

    /* ---- fixed point, in place: the fan-out must give the one-device bits ---- */
    {
        const size_t n = (size_t)QB * QN * 2;
        q15_t *a = malloc(n * sizeof(q15_t)), *b = malloc(n * sizeof(q15_t));
        fill_q15(a, n, 7u);
        memcpy(b, a, n * sizeof(q15_t));
        arm_cuda_set_devices(one, 1);
        expect(arm_cfft_batch_q15(&c15, a, QB, 0, 1), "arm_cfft_batch_q15, one device");
        arm_cuda_set_devices(three, 3);
        expect(arm_cfft_batch_q15(&c15, b, QB, 0, 1), "arm_cfft_batch_q15, three workers");
        CHECK(memcmp(a, b, n * sizeof(q15_t)) == 0, "arm_cfft_batch_q15: fan-out changes the result");
        free(a); free(b);
    }
    {
        const size_t n = (size_t)LB * LN * 2;
        q31_t *a = malloc(n * sizeof(q31_t)), *b = malloc(n * sizeof(q31_t));
        fill_q31(a, n, 9u);
        memcpy(b, a, n * sizeof(q31_t));
        arm_cuda_set_devices(one, 1);
        expect(arm_cfft_batch_q31(&c31, a, LB, 1, 1), "arm_cfft_batch_q31 (inverse), one device");
        arm_cuda_set_devices(three, 3);
        expect(arm_cfft_batch_q31(&c31, b, LB, 1, 1), "arm_cfft_batch_q31 (inverse), three workers");
        CHECK(memcmp(a, b, n * sizeof(q31_t)) == 0, "arm_cfft_batch_q31: fan-out changes the result");
        free(a); free(b);
    }
    {
        const size_t ni = (size_t)LB * 4 * LN, no = (size_t)LB * 4 * 2 * LN;      /* 4x the frames: 64 MiB of input */
        q31_t *src = malloc(ni * sizeof(q31_t)), *a = calloc(no, sizeof(q31_t)), *b = calloc(no, sizeof(q31_t));
        fill_q31(src, ni, 11u);
        arm_cuda_set_devices(one, 1);
        expect(arm_rfft_batch_q31(&r31, src, a, LB * 4), "arm_rfft_batch_q31, one device");
        arm_cuda_set_devices(three, 3);
        expect(arm_rfft_batch_q31(&r31, src, b, LB * 4), "arm_rfft_batch_q31, three workers");
        CHECK(memcmp(a, b, no * sizeof(q31_t)) == 0, "arm_rfft_batch_q31: fan-out changes the result");
        free(src); free(a); free(b);
    }
    /* ---- argument errors are reported, not executed ---- */
    CHECK(arm_rfft_fast_batch_f32(&rf, NULL, spec, 4, 0) == ARM_MATH_ARGUMENT_ERROR || !g_gpu, "NULL input accepted");
    CHECK(arm_rfft_fast_batch_f32(NULL, rin, spec, 4, 0) == ARM_MATH_ARGUMENT_ERROR, "NULL instance accepted");
    CHECK(arm_cfft_batch_f32(&cf, rin, 0, 0, 1) == ARM_MATH_SUCCESS || !g_gpu, "an empty batch is not an error");

    /* ---- four host threads at once, each with its own frames (a quarter of the batch) ---- */
    {
        pthread_t th[4];
        struct job jobs[4];
        arm_cuda_set_devices(one, 1);
        for (int t = 0; t < 4; t++) {
            jobs[t] = (struct job){&rf, rkeep + (size_t)t * (RB / 4) * RN, spec0 + (size_t)t * (RB / 4) * RN, RB / 4, t};
            CHECK(pthread_create(&th[t], NULL, thread_main, &jobs[t]) == 0, "pthread_create");
        }
        for (int t = 0; t < 4; t++) pthread_join(th[t], NULL);
    }
    arm_cuda_set_devices(NULL, 0);
    arm_cuda_release();
    arm_mfcc_release_plans();
    free(rin); free(rkeep); free(spec0); free(spec); free(back);
    if (g_fail) {
        fprintf(stderr, "host_api_check: %d check(s) failed\n", g_fail);
        return 1;
    }
    printf("host_api_check: ok (%s)\n", g_gpu ? "device 0: one worker, three workers, four host threads, small chunks -- same bits" : "no CUDA device: every call refused, nothing written");
    return 0;
}
