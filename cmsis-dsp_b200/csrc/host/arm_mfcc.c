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

/* ---- device plan cache ---- */
#define NPLAN 8
typedef struct { arm_mfcc_instance_f32 key; int device; void *plan; unsigned age; } slot_t;
static __thread slot_t g_slots[NPLAN];
static __thread unsigned g_clock;
static __thread void *g_stream;
static __thread int g_stream_dev = -1;
static __thread void *g_din, *g_dout;
static __thread size_t g_din_cap, g_dout_cap;

static int same_key(const arm_mfcc_instance_f32 *a, const arm_mfcc_instance_f32 *b)
{
    return a->dctCoefs == b->dctCoefs && a->filterCoefs == b->filterCoefs && a->windowCoefs == b->windowCoefs &&
           a->filterPos == b->filterPos && a->filterLengths == b->filterLengths && a->fftLen == b->fftLen &&
           a->nbMelFilters == b->nbMelFilters && a->nbDctOutputs == b->nbDctOutputs;
}

static int get_plan(const arm_mfcc_instance_f32 *S, void **plan)
{
    const int dev = cmsisdsp_cuda_get_device();
    if (dev < 0) return -1;
    int victim = -1;
    for (int k = 0; k < NPLAN; k++) {
        if (g_slots[k].plan && g_slots[k].device == dev && same_key(&g_slots[k].key, S)) {
            g_slots[k].age = ++g_clock;
            *plan = g_slots[k].plan;
            return 0;
        }
    }
    for (int k = 0; k < NPLAN && victim < 0; k++)
        if (!g_slots[k].plan) victim = k;                    /* a free slot, else the least recently used */
    if (victim < 0) {
        victim = 0;
        for (int k = 1; k < NPLAN; k++)
            if (g_slots[k].age < g_slots[victim].age) victim = k;
    }
    /* the rfft plan of fftLen first (tables the instance points at) */
    const arm_rfft_fast_instance_f32 *R = &S->rfft;
    if (!R->pTwiddleRFFT || !R->Sint.pTwiddle || R->fftLenRFFT != S->fftLen) return -1;
    if (!cmsisdsp_cuda_plan_ready(CMSISDSP_CUDA_F32, S->fftLen / 2) &&
        cmsisdsp_cuda_plan_upload(CMSISDSP_CUDA_F32, S->fftLen / 2, R->Sint.pTwiddle, R->Sint.pBitRevTable, R->Sint.bitRevLength)) return -1;
    if (!cmsisdsp_cuda_rfft_plan_ready(S->fftLen) && cmsisdsp_cuda_rfft_plan_upload(S->fftLen, R->pTwiddleRFFT)) return -1;
    void *p = 0;
    if (cmsisdsp_cuda_mfcc_plan_create(S->fftLen, S->nbMelFilters, S->nbDctOutputs, S->dctCoefs, S->filterPos,
                                       S->filterLengths, S->filterCoefs, S->windowCoefs, &p)) return -1;
    if (g_slots[victim].plan) cmsisdsp_cuda_mfcc_plan_destroy(g_slots[victim].plan);
    g_slots[victim].key = *S;
    g_slots[victim].device = dev;
    g_slots[victim].plan = p;
    g_slots[victim].age = ++g_clock;
    *plan = p;
    return 0;
}

static int grow(void **buf, size_t *cap, size_t bytes)
{
    if (*cap >= bytes) return 0;
    if (*buf) cmsisdsp_cuda_free(*buf);
    *buf = 0; *cap = 0;
    if (cmsisdsp_cuda_malloc(buf, bytes)) return -1;
    *cap = bytes;
    return 0;
}

arm_status arm_mfcc_batch_f32(const arm_mfcc_instance_f32 *S, const float32_t *pSrc, uint32_t hop,
                              float32_t *pDst, uint32_t nFrames)
{
    if (!S || !pSrc || !pDst || hop == 0 || (hop & 1u)) return ARM_MATH_ARGUMENT_ERROR;
    if (!S->dctCoefs || !S->filterCoefs || !S->windowCoefs || !S->filterPos || !S->filterLengths) return ARM_MATH_ARGUMENT_ERROR;
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    void *plan = 0;
    if (get_plan(S, &plan)) return ARM_MATH_ARGUMENT_ERROR;
    const int dev = cmsisdsp_cuda_get_device();
    if (g_stream_dev != dev) {
        if (cmsisdsp_cuda_stream_create(&g_stream)) return ARM_MATH_ARGUMENT_ERROR;
        g_stream_dev = dev;
        g_din = g_dout = 0; g_din_cap = g_dout_cap = 0;
    }
    const int inDev = cmsisdsp_cuda_is_device_pointer(pSrc), outDev = cmsisdsp_cuda_is_device_pointer(pDst);
    if (inDev < 0 || outDev < 0 || inDev != outDev) return ARM_MATH_ARGUMENT_ERROR;
    if (inDev) {
        if (cmsisdsp_cuda_mfcc_f32(plan, pSrc, hop, pDst, nFrames, g_stream)) return ARM_MATH_ARGUMENT_ERROR;
        return cmsisdsp_cuda_stream_synchronize(g_stream) ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
    }
    /* host buffers: chunks of frames through a device staging pair */
    const uint64_t maxChunk = ((uint64_t)64 << 20) / (sizeof(float32_t) * (hop > S->fftLen ? hop : S->fftLen)) + 1;
    for (uint64_t f = 0; f < nFrames;) {
        const uint64_t n = (nFrames - f < maxChunk) ? nFrames - f : maxChunk;
        const size_t inFloats = (size_t)(n - 1) * hop + S->fftLen, outFloats = (size_t)n * S->nbDctOutputs;
        if (grow(&g_din, &g_din_cap, inFloats * sizeof(float32_t)) || grow(&g_dout, &g_dout_cap, outFloats * sizeof(float32_t)))
            return ARM_MATH_ARGUMENT_ERROR;
        if (cmsisdsp_cuda_memcpy_h2d(g_din, pSrc + f * hop, inFloats * sizeof(float32_t), g_stream) ||
            cmsisdsp_cuda_mfcc_f32(plan, g_din, hop, g_dout, n, g_stream) ||
            cmsisdsp_cuda_memcpy_d2h(pDst + f * S->nbDctOutputs, g_dout, outFloats * sizeof(float32_t), g_stream) ||
            cmsisdsp_cuda_stream_synchronize(g_stream))
            return ARM_MATH_ARGUMENT_ERROR;
        f += n;
    }
    return ARM_MATH_SUCCESS;
}

extern arm_status arm_cuda_set_last_status(arm_status s);

void arm_mfcc_f32(const arm_mfcc_instance_f32 *S, float32_t *pSrc, float32_t *pDst, float32_t *pTmp)
{
    (void)pTmp;
    /* hop is irrelevant for one frame; any even value does */
    (void)arm_cuda_set_last_status(arm_mfcc_batch_f32(S, pSrc, 2, pDst, 1));
}
