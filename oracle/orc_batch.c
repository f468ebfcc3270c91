/*
 * oracle/orc_batch.c -- TEST INFRASTRUCTURE (see orc_fft.h).
 * Batch drivers for the oracle port: frames are contiguous and split statically
 * over `nthreads` pthreads, each thread looping the single-frame call
 * (the CPU-baseline plan of BASELINE.md section 4).
 */
#include "orc_fft.h"
#include <pthread.h>
#include <stdlib.h>

typedef struct {
    int kind;              /* 0 cfft_f32, 1 cfft_q31, 2 cfft_q15, 3 rfft_fast_f32 */
    uint32_t N;
    void *p, *out;
    uint64_t f0, f1;
    int ifft, bitrev;
} job_t;

static void *worker(void *arg)
{
    job_t *j = arg;
    for (uint64_t f = j->f0; f < j->f1; f++) {
        switch (j->kind) {
        case 0: orc_cfft_f32(j->N, (float *)j->p + 2ull * j->N * f, j->ifft, j->bitrev); break;
        case 1: orc_cfft_q31(j->N, (int32_t *)j->p + 2ull * j->N * f, j->ifft, j->bitrev); break;
        case 2: orc_cfft_q15(j->N, (int16_t *)j->p + 2ull * j->N * f, j->ifft, j->bitrev); break;
        default:
            orc_rfft_fast_f32(j->N, (float *)j->p + (uint64_t)j->N * f,
                              (float *)j->out + (uint64_t)j->N * f, j->ifft);
            break;
        }
    }
    return NULL;
}

static void run(int kind, uint32_t N, void *p, void *out, uint64_t nFrames, int ifft, int bitrev, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nFrames) nthreads = nFrames ? (int)nFrames : 1;
    (void)orc_twiddle_f32(16);                /* build tables before spawning */
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

void orc_cfft_f32_batch(uint32_t N, float *p, uint64_t n, int ifft, int bitrev, int nt) { run(0, N, p, 0, n, ifft, bitrev, nt); }
void orc_cfft_q31_batch(uint32_t N, int32_t *p, uint64_t n, int ifft, int bitrev, int nt) { run(1, N, p, 0, n, ifft, bitrev, nt); }
void orc_cfft_q15_batch(uint32_t N, int16_t *p, uint64_t n, int ifft, int bitrev, int nt) { run(2, N, p, 0, n, ifft, bitrev, nt); }
void orc_rfft_fast_f32_batch(uint32_t N, float *p, float *out, uint64_t n, int ifft, int nt) { run(3, N, p, out, n, ifft, 0, nt); }
