/*
 * arm_mfcc.c -- arm_mfcc_init_f32 / arm_mfcc_f32 of the reference API plus the batched extension,
 * thin C over libcmsisdsp_cuda (cmsisdsp_cuda_mfcc_*).
 *   reference: Source/TransformFunctions/arm_mfcc_init_f32.c:91-121, arm_mfcc_f32.c:88-174
 *
 * The instance struct is plain data owned by the caller (the reference never allocates), so the
 * device-side plan -- copies of the coefficient arrays -- is cached here, per thread, keyed by the
 * instance's contents (array pointers and dimensions) and the current device.
 */
#include "arm_math_types.h"
#include "dsp/transform_functions.h"
#include "cmsisdsp_cuda.h"

#include <stdlib.h>
#include <string.h>

arm_status arm_mfcc_init_f32(arm_mfcc_instance_f32 *S, uint32_t fftLen, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs)
{
    S->fftLen = fftLen;
    S->nbMelFilters = nbMelFilters;
    S->nbDctOutputs = nbDctOutputs;
    S->dctCoefs = dctCoefs;
    S->filterPos = filterPos;
    S->filterLengths = filterLengths;
    S->filterCoefs = filterCoefs;
    S->windowCoefs = windowCoefs;
    return arm_rfft_fast_init_f32(&S->rfft, (uint16_t)fftLen);
}

#define MFCC_INIT(LEN)                                                                                          \
arm_status arm_mfcc_init_##LEN##_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,     \
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths, \
                             const float32_t *filterCoefs, const float32_t *windowCoefs)                         \
{ return arm_mfcc_init_f32(S, LEN, nbMelFilters, nbDctOutputs, dctCoefs, filterPos, filterLengths, filterCoefs, windowCoefs); }
MFCC_INIT(32) MFCC_INIT(64) MFCC_INIT(128) MFCC_INIT(256) MFCC_INIT(512) MFCC_INIT(1024) MFCC_INIT(2048) MFCC_INIT(4096)

/* ---- device plan cache ----
 * The instance is plain data owned by the caller, so the device-side plan (copies of the coefficient arrays) is kept
 * here, process-wide, keyed by the device and by the CONTENT of the arrays: a 64-bit FNV-1a hash over the dimensions
 * and every coefficient (about 9 KB for the 1024-point configuration, a few microseconds per call).  Keying on the
 * array addresses alone would hand stale coefficients to a caller that rewrote its tables in place or whose
 * allocator reused the addresses. */
#include <pthread.h>
#include "arm_cuda_engine.h"

#define NPLAN 16
typedef struct { uint64_t hash; int device; void *plan; unsigned age; } slot_t;
static slot_t g_slots[NPLAN];
static unsigned g_clock;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;

static uint64_t fnv(uint64_t h, const void *data, size_t bytes)
{
    const unsigned char *p = (const unsigned char *)data;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 1099511628211ull;
    }
    for (; i < bytes; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}
static uint64_t content_hash(const arm_mfcc_instance_f32 *S)
{
    uint64_t h = 1469598103934665603ull;
    const uint32_t dims[3] = { S->fftLen, S->nbMelFilters, S->nbDctOutputs };
    size_t taps = 0;
    for (uint32_t f = 0; f < S->nbMelFilters; f++) taps += S->filterLengths[f];
    h = fnv(h, dims, sizeof dims);
    h = fnv(h, S->filterPos, S->nbMelFilters * sizeof(uint32_t));
    h = fnv(h, S->filterLengths, S->nbMelFilters * sizeof(uint32_t));
    h = fnv(h, S->dctCoefs, (size_t)S->nbMelFilters * S->nbDctOutputs * sizeof(float32_t));
    h = fnv(h, S->filterCoefs, taps * sizeof(float32_t));
    h = fnv(h, S->windowCoefs, (size_t)S->fftLen * sizeof(float32_t));
    return h;
}

/* the plan of S on the CURRENT device (created on first use); the rfft tables the instance points at go first */
static int get_plan(const arm_mfcc_instance_f32 *S, uint64_t hash, void **plan)
{
    const int dev = cmsisdsp_cuda_get_device();
    if (dev < 0) return CMSISDSP_CUDA_ERR_NO_DEVICE;
    const arm_rfft_fast_instance_f32 *R = &S->rfft;
    if (!R->pTwiddleRFFT || !R->Sint.pTwiddle || R->fftLenRFFT != S->fftLen) return CMSISDSP_CUDA_ERR_ARGUMENT;
    int rc = cmsisdsp_cuda_plan_upload(CMSISDSP_CUDA_F32, S->fftLen / 2, R->Sint.pTwiddle, R->Sint.pBitRevTable, R->Sint.bitRevLength);
    if (!rc) rc = cmsisdsp_cuda_rfft_plan_upload(S->fftLen, R->pTwiddleRFFT);
    if (rc) return rc;
    pthread_mutex_lock(&g_mu);
    int victim = -1;
    for (int k = 0; k < NPLAN; k++)
        if (g_slots[k].plan && g_slots[k].device == dev && g_slots[k].hash == hash) {
            g_slots[k].age = ++g_clock;
            *plan = g_slots[k].plan;
            pthread_mutex_unlock(&g_mu);
            return 0;
        }
    for (int k = 0; k < NPLAN && victim < 0; k++)
        if (!g_slots[k].plan) victim = k;
    void *p = 0;
    rc = cmsisdsp_cuda_mfcc_plan_create(S->fftLen, S->nbMelFilters, S->nbDctOutputs, S->dctCoefs, S->filterPos,
                                        S->filterLengths, S->filterCoefs, S->windowCoefs, &p);
    if (rc) {
        pthread_mutex_unlock(&g_mu);
        return rc;
    }
    if (victim < 0) {
        /* every slot is taken: this plan is used for the call and stays uncached (a cached plan may be in use by
         * another thread's kernel, so none is evicted) */
        pthread_mutex_unlock(&g_mu);
        *plan = p;
        return 1;
    }
    g_slots[victim].hash = hash;
    g_slots[victim].device = dev;
    g_slots[victim].plan = p;
    g_slots[victim].age = ++g_clock;
    pthread_mutex_unlock(&g_mu);
    *plan = p;
    return 0;
}

/* drop every cached device plan (the caller guarantees that no MFCC call is in flight) */
void arm_mfcc_release_plans(void)
{
    pthread_mutex_lock(&g_mu);
    for (int k = 0; k < NPLAN; k++) {
        if (g_slots[k].plan) cmsisdsp_cuda_mfcc_plan_destroy(g_slots[k].plan);
        g_slots[k].plan = 0;
    }
    pthread_mutex_unlock(&g_mu);
}

typedef struct { const arm_mfcc_instance_f32 *S; uint64_t hash; uint32_t hop; } mfcc_args;
/* worker threads live for one call: the plan each one resolved in prepare() is found again at launch */
static __thread void *t_plan;
static __thread int t_plan_uncached;

static int mfcc_prepare(const arm_cuda_job *job)
{
    const mfcc_args *a = (const mfcc_args *)job->self;
    const int rc = get_plan(a->S, a->hash, &t_plan);
    t_plan_uncached = (rc == 1);
    return rc == 1 ? 0 : rc;
}
static int mfcc_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const mfcc_args *a = (const mfcc_args *)job->self;
    (void)doutB;
    return cmsisdsp_cuda_mfcc_f32(t_plan, din, a->hop, dout, n, stream);
}

arm_status arm_mfcc_batch_f32(const arm_mfcc_instance_f32 *S, const float32_t *pSrc, uint32_t hop,
                              float32_t *pDst, uint32_t nFrames)
{
    if (!S || !pSrc || !pDst || hop == 0 || (hop & 1u)) return ARM_MATH_ARGUMENT_ERROR;
    if (!S->dctCoefs || !S->filterCoefs || !S->windowCoefs || !S->filterPos || !S->filterLengths) return ARM_MATH_ARGUMENT_ERROR;
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    const mfcc_args a = { S, content_hash(S), hop };
    arm_cuda_job job = {0};
    job.inStride = (size_t)hop * sizeof(float32_t);
    job.inFrame = (size_t)S->fftLen * sizeof(float32_t);
    job.outStride = job.outFrame = (size_t)S->nbDctOutputs * sizeof(float32_t);
    job.prepare = mfcc_prepare;
    job.launch = mfcc_launch;
    job.self = &a;
    const arm_status st = arm_cuda_run(&job, pSrc, pDst, nFrames);
    if (t_plan_uncached && t_plan) {             /* single-device overflow plan of this thread (see get_plan) */
        cmsisdsp_cuda_mfcc_plan_destroy(t_plan);
        t_plan = 0;
        t_plan_uncached = 0;
    }
    return st;
}

void arm_mfcc_f32(const arm_mfcc_instance_f32 *S, float32_t *pSrc, float32_t *pDst, float32_t *pTmp)
{
    (void)pTmp;
    /* hop is irrelevant for one frame; any even value does */
    (void)arm_cuda_set_last_status(arm_mfcc_batch_f32(S, pSrc, 2, pDst, 1));
}
