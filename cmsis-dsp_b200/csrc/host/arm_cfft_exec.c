/*
 * arm_cfft_exec.c -- exec functions of the FFT path: thin C over libcmsisdsp_cuda.
 *
 *   arm_cfft_f32 / q31 / q15   reference: arm_cfft_f32.c:1243-1298, arm_cfft_q31.c:704-755, arm_cfft_q15.c:671-722
 *   arm_cfft_f64               reference: arm_cfft_f64.c:262-312
 *   arm_rfft_fast_f32          reference: arm_rfft_fast_f32.c:675-699
 *   arm_rfft_q31 / q15         reference: arm_rfft_q31.c:145-181, arm_rfft_q15.c:148-182
 *   arm_*_batch_*              B200 extension (include/dsp/transform_functions.h)
 *
 * Data may live in host or device memory.  Host buffers are streamed through the device
 * in chunks on three streams (copy-in, kernel, copy-out overlap; buffers obtained from
 * cmsisdsp_cuda_host_alloc are pinned and get full-duplex PCIe).  Device buffers are
 * transformed in place on the library's per-thread stream.  No CPU fallback exists: when
 * the shim reports an error the batched call returns ARM_MATH_ARGUMENT_ERROR and the
 * legacy void call records it for arm_cuda_last_status().
 *
 * Like the reference, unsupported lengths in a hand-built instance make the legacy exec
 * functions a no-op (arm_cfft_f32.c:1263-1280 falls through its switch).
 */
#include "arm_math_types.h"
#include "dsp/transform_functions.h"
#include "cmsisdsp_cuda.h"

#include <stdlib.h>

#define NSTREAM 3
#define CHUNK_BYTES ((size_t)16 << 20)

typedef struct {
    int device;
    void *stream[NSTREAM];
    void *buf[NSTREAM][2];      /* staging: [0] data / rfft input, [1] rfft output */
    size_t cap[NSTREAM][2];
} tls_ctx;

static __thread tls_ctx g_ctx = { -1, {0}, {{0}}, {{0}} };
static __thread arm_status g_last = ARM_MATH_SUCCESS;

arm_status arm_cuda_last_status(void) { return g_last; }
/* used by the other exec files of this library (arm_mfcc.c) */
arm_status arm_cuda_set_last_status(arm_status s) { g_last = s; return s; }

static int ctx_ready(void)
{
    int dev = cmsisdsp_cuda_get_device();
    if (dev < 0) return CMSISDSP_CUDA_ERR_NO_DEVICE;
    if (g_ctx.device == dev) return 0;
    /* first use on this thread, or the caller switched device: (re)create streams lazily */
    for (int i = 0; i < NSTREAM; i++) {
        g_ctx.stream[i] = 0;
        g_ctx.buf[i][0] = g_ctx.buf[i][1] = 0;
        g_ctx.cap[i][0] = g_ctx.cap[i][1] = 0;
        int rc = cmsisdsp_cuda_stream_create(&g_ctx.stream[i]);
        if (rc) return rc;
    }
    g_ctx.device = dev;
    return 0;
}

static int staging(int s, int which, size_t bytes, void **out)
{
    if (g_ctx.cap[s][which] < bytes) {
        if (g_ctx.buf[s][which]) cmsisdsp_cuda_free(g_ctx.buf[s][which]);
        g_ctx.buf[s][which] = 0;
        g_ctx.cap[s][which] = 0;
        int rc = cmsisdsp_cuda_malloc(&g_ctx.buf[s][which], bytes);
        if (rc) return rc;
        g_ctx.cap[s][which] = bytes;
    }
    *out = g_ctx.buf[s][which];
    return 0;
}

static int ensure_plan(int type, uint32_t fftLen, const void *tw, const uint16_t *br, uint16_t brLen)
{
    if (cmsisdsp_cuda_plan_ready(type, fftLen)) return 0;
    return cmsisdsp_cuda_plan_upload(type, fftLen, tw, br, brLen);
}

typedef int (*cfft_fn)(void *, uint32_t, uint64_t, uint8_t, uint8_t, void *);

static int valid_len(uint32_t n) { return n >= 16 && n <= 4096 && (n & (n - 1)) == 0; }

static arm_status cfft_batch(int type, cfft_fn fn, size_t scalarBytes, uint32_t fftLen, const void *tw,
                             const uint16_t *br, uint16_t brLen, void *p, uint64_t nFrames,
                             uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!p || !tw) return ARM_MATH_ARGUMENT_ERROR;
    if (!valid_len(fftLen)) return ARM_MATH_ARGUMENT_ERROR;
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    if (ctx_ready()) return ARM_MATH_ARGUMENT_ERROR;
    if (ensure_plan(type, fftLen, tw, br, brLen)) return ARM_MATH_ARGUMENT_ERROR;

    const size_t frameBytes = (size_t)2 * fftLen * scalarBytes;
    const int onDevice = cmsisdsp_cuda_is_device_pointer(p);
    if (onDevice < 0) return ARM_MATH_ARGUMENT_ERROR;
    if (onDevice) {
        if (fn(p, fftLen, nFrames, ifftFlag, bitReverseFlag, g_ctx.stream[0])) return ARM_MATH_ARGUMENT_ERROR;
        return cmsisdsp_cuda_stream_synchronize(g_ctx.stream[0]) ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
    }
    uint64_t perChunk = CHUNK_BYTES / frameBytes;
    if (perChunk == 0) perChunk = 1;
    int rc = 0, s = 0;
    for (uint64_t f = 0; f < nFrames && !rc; f += perChunk, s = (s + 1) % NSTREAM) {
        const uint64_t n = (nFrames - f < perChunk) ? nFrames - f : perChunk;
        char *h = (char *)p + f * frameBytes;
        void *d;
        /* the stream serialises reuse of its staging buffer */
        if ((rc = staging(s, 0, (size_t)perChunk * frameBytes, &d))) break;
        if ((rc = cmsisdsp_cuda_memcpy_h2d(d, h, (size_t)n * frameBytes, g_ctx.stream[s]))) break;
        if ((rc = fn(d, fftLen, n, ifftFlag, bitReverseFlag, g_ctx.stream[s]))) break;
        rc = cmsisdsp_cuda_memcpy_d2h(h, d, (size_t)n * frameBytes, g_ctx.stream[s]);
    }
    for (int i = 0; i < NSTREAM; i++)
        if (cmsisdsp_cuda_stream_synchronize(g_ctx.stream[i])) rc = rc ? rc : -1;
    return rc ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
}

arm_status arm_cfft_batch_f32(const arm_cfft_instance_f32 *S, float32_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_F32, cmsisdsp_cuda_cfft_f32, sizeof(float32_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag);
}
arm_status arm_cfft_batch_q31(const arm_cfft_instance_q31 *S, q31_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_Q31, cmsisdsp_cuda_cfft_q31, sizeof(q31_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag);
}
arm_status arm_cfft_batch_q15(const arm_cfft_instance_q15 *S, q15_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_Q15, cmsisdsp_cuda_cfft_q15, sizeof(q15_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag);
}

arm_status arm_cfft_batch_f64(const arm_cfft_instance_f64 *S, float64_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_F64, cmsisdsp_cuda_cfft_f64, sizeof(float64_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag);
}

/* clobber != 0: also leave the N/2-point CFFT in p after a forward transform, the
 * side effect of the reference's in-place CFFT on the input buffer (rfft_fast_f32.c:694) */
static arm_status rfft_batch(const arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut,
                             uint64_t nFrames, uint8_t ifftFlag, int clobber)
{
    if (!S || !p || !pOut || p == pOut || !S->pTwiddleRFFT || !S->Sint.pTwiddle) return ARM_MATH_ARGUMENT_ERROR;
    const uint32_t N = S->fftLenRFFT;
    if (N < 32 || !valid_len(N) || S->Sint.fftLen != N / 2) return ARM_MATH_ARGUMENT_ERROR;
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    if (ctx_ready()) return ARM_MATH_ARGUMENT_ERROR;
    if (ensure_plan(CMSISDSP_CUDA_F32, N / 2, S->Sint.pTwiddle, S->Sint.pBitRevTable, S->Sint.bitRevLength))
        return ARM_MATH_ARGUMENT_ERROR;
    if (!cmsisdsp_cuda_rfft_plan_ready(N) && cmsisdsp_cuda_rfft_plan_upload(N, S->pTwiddleRFFT))
        return ARM_MATH_ARGUMENT_ERROR;

    const size_t frameBytes = (size_t)N * sizeof(float32_t);
    const int inDev = cmsisdsp_cuda_is_device_pointer(p), outDev = cmsisdsp_cuda_is_device_pointer(pOut);
    if (inDev < 0 || outDev < 0 || inDev != outDev) return ARM_MATH_ARGUMENT_ERROR;
    if (inDev) {
        if (cmsisdsp_cuda_rfft_fast_f32(p, pOut, N, nFrames, ifftFlag, g_ctx.stream[0])) return ARM_MATH_ARGUMENT_ERROR;
        if (clobber && !ifftFlag && cmsisdsp_cuda_cfft_f32(p, N / 2, nFrames, 0, 1, g_ctx.stream[0])) return ARM_MATH_ARGUMENT_ERROR;
        return cmsisdsp_cuda_stream_synchronize(g_ctx.stream[0]) ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
    }
    uint64_t perChunk = CHUNK_BYTES / frameBytes;
    if (perChunk == 0) perChunk = 1;
    int rc = 0, s = 0;
    for (uint64_t f = 0; f < nFrames && !rc; f += perChunk, s = (s + 1) % NSTREAM) {
        const uint64_t n = (nFrames - f < perChunk) ? nFrames - f : perChunk;
        char *hin = (char *)p + f * frameBytes, *hout = (char *)pOut + f * frameBytes;
        void *din, *dout;
        if ((rc = staging(s, 0, (size_t)perChunk * frameBytes, &din))) break;
        if ((rc = staging(s, 1, (size_t)perChunk * frameBytes, &dout))) break;
        if ((rc = cmsisdsp_cuda_memcpy_h2d(din, hin, (size_t)n * frameBytes, g_ctx.stream[s]))) break;
        if ((rc = cmsisdsp_cuda_rfft_fast_f32(din, dout, N, n, ifftFlag, g_ctx.stream[s]))) break;
        if ((rc = cmsisdsp_cuda_memcpy_d2h(hout, dout, (size_t)n * frameBytes, g_ctx.stream[s]))) break;
        if (clobber && !ifftFlag) {
            if ((rc = cmsisdsp_cuda_cfft_f32(din, N / 2, n, 0, 1, g_ctx.stream[s]))) break;
            rc = cmsisdsp_cuda_memcpy_d2h(hin, din, (size_t)n * frameBytes, g_ctx.stream[s]);
        }
    }
    for (int i = 0; i < NSTREAM; i++)
        if (cmsisdsp_cuda_stream_synchronize(g_ctx.stream[i])) rc = rc ? rc : -1;
    return rc ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
}

arm_status arm_rfft_fast_batch_f32(const arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut,
                                   uint32_t nFrames, uint8_t ifftFlag)
{
    return rfft_batch(S, p, pOut, nFrames, ifftFlag, 0);
}

/* arm_rfft_fast_f64: same contract as rfft_batch above, 8-byte scalars, the f64 shim entry */
static arm_status rfft64_batch(const arm_rfft_fast_instance_f64 *S, float64_t *p, float64_t *pOut,
                               uint64_t nFrames, uint8_t ifftFlag, int clobber)
{
    if (!S || !p || !pOut || p == pOut || !S->pTwiddleRFFT || !S->Sint.pTwiddle) return ARM_MATH_ARGUMENT_ERROR;
    const uint32_t N = S->fftLenRFFT;
    if (N < 32 || !valid_len(N)) return ARM_MATH_ARGUMENT_ERROR;
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    if (ctx_ready()) return ARM_MATH_ARGUMENT_ERROR;
    if (ensure_plan(CMSISDSP_CUDA_F64, N / 2, S->Sint.pTwiddle, S->Sint.pBitRevTable, S->Sint.bitRevLength))
        return ARM_MATH_ARGUMENT_ERROR;
    if (!cmsisdsp_cuda_rfft_f64_plan_ready(N) && cmsisdsp_cuda_rfft_f64_plan_upload(N, S->pTwiddleRFFT))
        return ARM_MATH_ARGUMENT_ERROR;

    const size_t frameBytes = (size_t)N * sizeof(float64_t);
    const int inDev = cmsisdsp_cuda_is_device_pointer(p), outDev = cmsisdsp_cuda_is_device_pointer(pOut);
    if (inDev < 0 || outDev < 0 || inDev != outDev) return ARM_MATH_ARGUMENT_ERROR;
    if (inDev) {
        if (cmsisdsp_cuda_rfft_fast_f64(p, pOut, N, nFrames, ifftFlag, g_ctx.stream[0])) return ARM_MATH_ARGUMENT_ERROR;
        if (clobber && !ifftFlag && cmsisdsp_cuda_cfft_f64(p, N / 2, nFrames, 0, 1, g_ctx.stream[0])) return ARM_MATH_ARGUMENT_ERROR;
        return cmsisdsp_cuda_stream_synchronize(g_ctx.stream[0]) ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
    }
    uint64_t perChunk = CHUNK_BYTES / frameBytes;
    if (perChunk == 0) perChunk = 1;
    int rc = 0, s = 0;
    for (uint64_t f = 0; f < nFrames && !rc; f += perChunk, s = (s + 1) % NSTREAM) {
        const uint64_t n = (nFrames - f < perChunk) ? nFrames - f : perChunk;
        char *hin = (char *)p + f * frameBytes, *hout = (char *)pOut + f * frameBytes;
        void *din, *dout;
        if ((rc = staging(s, 0, (size_t)perChunk * frameBytes, &din))) break;
        if ((rc = staging(s, 1, (size_t)perChunk * frameBytes, &dout))) break;
        if ((rc = cmsisdsp_cuda_memcpy_h2d(din, hin, (size_t)n * frameBytes, g_ctx.stream[s]))) break;
        if ((rc = cmsisdsp_cuda_rfft_fast_f64(din, dout, N, n, ifftFlag, g_ctx.stream[s]))) break;
        if ((rc = cmsisdsp_cuda_memcpy_d2h(hout, dout, (size_t)n * frameBytes, g_ctx.stream[s]))) break;
        if (clobber && !ifftFlag) {
            if ((rc = cmsisdsp_cuda_cfft_f64(din, N / 2, n, 0, 1, g_ctx.stream[s]))) break;
            rc = cmsisdsp_cuda_memcpy_d2h(hin, din, (size_t)n * frameBytes, g_ctx.stream[s]);
        }
    }
    for (int i = 0; i < NSTREAM; i++)
        if (cmsisdsp_cuda_stream_synchronize(g_ctx.stream[i])) rc = rc ? rc : -1;
    return rc ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
}

arm_status arm_rfft_fast_batch_f64(const arm_rfft_fast_instance_f64 *S, float64_t *p, float64_t *pOut,
                                   uint32_t nFrames, uint8_t ifftFlag)
{
    if (S && S->Sint.fftLen != S->fftLenRFFT / 2) return ARM_MATH_ARGUMENT_ERROR;
    return rfft64_batch(S, p, pOut, nFrames, ifftFlag, 0);
}

/* ---- legacy single-frame signatures ---- */

static int legacy_len_ok(uint32_t n) { return valid_len(n); }

void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!legacy_len_ok(S->fftLen)) { g_last = ARM_MATH_SUCCESS; return; }
    g_last = arm_cfft_batch_f32(S, p1, 1, ifftFlag, bitReverseFlag);
}
void arm_cfft_f64(const arm_cfft_instance_f64 *S, float64_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!legacy_len_ok(S->fftLen)) { g_last = ARM_MATH_SUCCESS; return; }
    g_last = arm_cfft_batch_f64(S, p1, 1, ifftFlag, bitReverseFlag);
}
/* arm_rfft_fast_f64.c:207-233: sets Sint.fftLen like the reference; the forward call leaves the N/2-point CFFT in p */
void arm_rfft_fast_f64(arm_rfft_fast_instance_f64 *S, float64_t *p, float64_t *pOut, uint8_t ifftFlag)
{
    S->Sint.fftLen = S->fftLenRFFT / 2;
    if (S->fftLenRFFT < 32 || !legacy_len_ok(S->fftLenRFFT)) { g_last = ARM_MATH_SUCCESS; return; }
    g_last = rfft64_batch(S, p, pOut, 1, ifftFlag, 1);
}
void arm_cfft_q31(const arm_cfft_instance_q31 *S, q31_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!legacy_len_ok(S->fftLen)) { g_last = ARM_MATH_SUCCESS; return; }
    g_last = arm_cfft_batch_q31(S, p1, 1, ifftFlag, bitReverseFlag);
}
void arm_cfft_q15(const arm_cfft_instance_q15 *S, q15_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!legacy_len_ok(S->fftLen)) { g_last = ARM_MATH_SUCCESS; return; }
    g_last = arm_cfft_batch_q15(S, p1, 1, ifftFlag, bitReverseFlag);
}
void arm_rfft_fast_f32(const arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut, uint8_t ifftFlag)
{
    g_last = rfft_batch(S, p, pOut, 1, ifftFlag, 1);
}

/* ---- fixed-point real FFT ---- */

typedef int (*rfix_fn)(const void *, void *, uint32_t, uint64_t, uint8_t, void *);

/* clobber != 0: a forward transform also leaves the fftLenReal/2-point CFFT in pSrc, the side effect of the
 * reference's in-place CFFT on the source buffer (arm_rfft_q31.c:174, arm_rfft_q15.c:176) */
static arm_status rfix_batch(int type, rfix_fn fn, cfft_fn cfn, size_t scalarBytes, uint32_t N, uint8_t ifftFlagR,
                             uint8_t bitReverseFlagR, uint32_t modifier, const void *coefA, const void *coefB,
                             const void *cfftTw, const uint16_t *br, uint16_t brLen, uint32_t cfftLen,
                             void *pSrc, void *pDst, uint64_t nFrames, int clobber)
{
    if (!pSrc || !pDst || pSrc == pDst || !coefA || !coefB || !cfftTw) return ARM_MATH_ARGUMENT_ERROR;
    if (N < 32 || N > 8192 || (N & (N - 1)) != 0 || cfftLen != N / 2 || bitReverseFlagR != 1) return ARM_MATH_ARGUMENT_ERROR;
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    if (ctx_ready()) return ARM_MATH_ARGUMENT_ERROR;
    if (ensure_plan(type, N / 2, cfftTw, br, brLen)) return ARM_MATH_ARGUMENT_ERROR;
    if (!cmsisdsp_cuda_rfft_fix_plan_ready(type, N) && cmsisdsp_cuda_rfft_fix_plan_upload(type, N, coefA, coefB, modifier))
        return ARM_MATH_ARGUMENT_ERROR;

    const size_t inBytes = (size_t)(ifftFlagR ? 2 * N : N) * scalarBytes, outBytes = (size_t)(ifftFlagR ? N : 2 * N) * scalarBytes;
    const int inDev = cmsisdsp_cuda_is_device_pointer(pSrc), outDev = cmsisdsp_cuda_is_device_pointer(pDst);
    if (inDev < 0 || outDev < 0 || inDev != outDev) return ARM_MATH_ARGUMENT_ERROR;
    if (inDev) {
        if (fn(pSrc, pDst, N, nFrames, ifftFlagR, g_ctx.stream[0])) return ARM_MATH_ARGUMENT_ERROR;
        if (clobber && !ifftFlagR && cfn(pSrc, N / 2, nFrames, 0, 1, g_ctx.stream[0])) return ARM_MATH_ARGUMENT_ERROR;
        return cmsisdsp_cuda_stream_synchronize(g_ctx.stream[0]) ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
    }
    uint64_t perChunk = CHUNK_BYTES / (inBytes > outBytes ? inBytes : outBytes);
    if (perChunk == 0) perChunk = 1;
    int rc = 0, s = 0;
    for (uint64_t f = 0; f < nFrames && !rc; f += perChunk, s = (s + 1) % NSTREAM) {
        const uint64_t n = (nFrames - f < perChunk) ? nFrames - f : perChunk;
        char *hin = (char *)pSrc + f * inBytes, *hout = (char *)pDst + f * outBytes;
        void *din, *dout;
        if ((rc = staging(s, 0, (size_t)perChunk * inBytes, &din))) break;
        if ((rc = staging(s, 1, (size_t)perChunk * outBytes, &dout))) break;
        if ((rc = cmsisdsp_cuda_memcpy_h2d(din, hin, (size_t)n * inBytes, g_ctx.stream[s]))) break;
        if ((rc = fn(din, dout, N, n, ifftFlagR, g_ctx.stream[s]))) break;
        if ((rc = cmsisdsp_cuda_memcpy_d2h(hout, dout, (size_t)n * outBytes, g_ctx.stream[s]))) break;
        if (clobber && !ifftFlagR) {
            if ((rc = cfn(din, N / 2, n, 0, 1, g_ctx.stream[s]))) break;
            rc = cmsisdsp_cuda_memcpy_d2h(hin, din, (size_t)n * inBytes, g_ctx.stream[s]);
        }
    }
    for (int i = 0; i < NSTREAM; i++)
        if (cmsisdsp_cuda_stream_synchronize(g_ctx.stream[i])) rc = rc ? rc : -1;
    return rc ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
}

static arm_status rfix_q31(const arm_rfft_instance_q31 *S, q31_t *pSrc, q31_t *pDst, uint64_t nFrames, int clobber)
{
    if (!S || !S->pCfft) return ARM_MATH_ARGUMENT_ERROR;
    return rfix_batch(CMSISDSP_CUDA_Q31, cmsisdsp_cuda_rfft_q31, cmsisdsp_cuda_cfft_q31, sizeof(q31_t), S->fftLenReal, S->ifftFlagR,
                      S->bitReverseFlagR, S->twidCoefRModifier, S->pTwiddleAReal, S->pTwiddleBReal, S->pCfft->pTwiddle,
                      S->pCfft->pBitRevTable, S->pCfft->bitRevLength, S->pCfft->fftLen, pSrc, pDst, nFrames, clobber);
}
static arm_status rfix_q15(const arm_rfft_instance_q15 *S, q15_t *pSrc, q15_t *pDst, uint64_t nFrames, int clobber)
{
    if (!S || !S->pCfft) return ARM_MATH_ARGUMENT_ERROR;
    return rfix_batch(CMSISDSP_CUDA_Q15, cmsisdsp_cuda_rfft_q15, cmsisdsp_cuda_cfft_q15, sizeof(q15_t), S->fftLenReal, S->ifftFlagR,
                      S->bitReverseFlagR, S->twidCoefRModifier, S->pTwiddleAReal, S->pTwiddleBReal, S->pCfft->pTwiddle,
                      S->pCfft->pBitRevTable, S->pCfft->bitRevLength, S->pCfft->fftLen, pSrc, pDst, nFrames, clobber);
}
arm_status arm_rfft_batch_q31(const arm_rfft_instance_q31 *S, const q31_t *pSrc, q31_t *pDst, uint32_t nFrames)
{
    return rfix_q31(S, (q31_t *)pSrc, pDst, nFrames, 0);
}
arm_status arm_rfft_batch_q15(const arm_rfft_instance_q15 *S, const q15_t *pSrc, q15_t *pDst, uint32_t nFrames)
{
    return rfix_q15(S, (q15_t *)pSrc, pDst, nFrames, 0);
}
void arm_rfft_q31(const arm_rfft_instance_q31 *S, q31_t *pSrc, q31_t *pDst) { g_last = rfix_q31(S, pSrc, pDst, 1, 1); }
void arm_rfft_q15(const arm_rfft_instance_q15 *S, q15_t *pSrc, q15_t *pDst) { g_last = rfix_q15(S, pSrc, pDst, 1, 1); }

/* ---- arm_cfft_f32 fused with arm_cmplx_mag[_squared]_f32 (mode 0 / 1) or with arm_cmplx_mag_f32 + arm_max_f32 (mode 2) ---- */
static arm_status spectrum_batch(const arm_cfft_instance_f32 *S, const float32_t *pSrc, void *pOut, uint32_t *pIndex,
                                 uint64_t nFrames, uint8_t ifftFlag, int mode)
{
    if (!S || !pSrc || !pOut || (mode == 2 && !pIndex) || !S->pTwiddle || (const void *)pSrc == pOut) return ARM_MATH_ARGUMENT_ERROR;
    const uint32_t N = S->fftLen;
    if (!valid_len(N)) return ARM_MATH_ARGUMENT_ERROR;
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    if (ctx_ready()) return ARM_MATH_ARGUMENT_ERROR;
    if (ensure_plan(CMSISDSP_CUDA_F32, N, S->pTwiddle, S->pBitRevTable, S->bitRevLength)) return ARM_MATH_ARGUMENT_ERROR;
    const size_t inBytes = (size_t)2 * N * sizeof(float32_t), outBytes = (mode == 2) ? sizeof(float32_t) : (size_t)N * sizeof(float32_t);
    const int inDev = cmsisdsp_cuda_is_device_pointer(pSrc), outDev = cmsisdsp_cuda_is_device_pointer(pOut);
    const int idxDev = (mode == 2) ? cmsisdsp_cuda_is_device_pointer(pIndex) : outDev;
    if (inDev < 0 || outDev < 0 || inDev != outDev || idxDev != outDev) return ARM_MATH_ARGUMENT_ERROR;
    if (inDev) {
        int rc = (mode == 2) ? cmsisdsp_cuda_cfft_peak_f32(pSrc, pOut, pIndex, N, nFrames, ifftFlag, g_ctx.stream[0])
                             : cmsisdsp_cuda_cfft_mag_f32(pSrc, pOut, N, nFrames, ifftFlag, (uint8_t)mode, g_ctx.stream[0]);
        if (rc) return ARM_MATH_ARGUMENT_ERROR;
        return cmsisdsp_cuda_stream_synchronize(g_ctx.stream[0]) ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
    }
    uint64_t perChunk = CHUNK_BYTES / inBytes;
    if (perChunk == 0) perChunk = 1;
    int rc = 0, s = 0;
    for (uint64_t f = 0; f < nFrames && !rc; f += perChunk, s = (s + 1) % NSTREAM) {
        const uint64_t n = (nFrames - f < perChunk) ? nFrames - f : perChunk;
        void *din, *dout;
        /* staging [1] holds the magnitudes, or the peak values followed by the peak indices */
        if ((rc = staging(s, 0, (size_t)perChunk * inBytes, &din))) break;
        if ((rc = staging(s, 1, (size_t)perChunk * (mode == 2 ? 8 : outBytes), &dout))) break;
        if ((rc = cmsisdsp_cuda_memcpy_h2d(din, (const char *)pSrc + f * inBytes, (size_t)n * inBytes, g_ctx.stream[s]))) break;
        if (mode == 2) {
            void *didx = (char *)dout + (size_t)perChunk * 4;
            if ((rc = cmsisdsp_cuda_cfft_peak_f32(din, dout, didx, N, n, ifftFlag, g_ctx.stream[s]))) break;
            if ((rc = cmsisdsp_cuda_memcpy_d2h((float32_t *)pOut + f, dout, (size_t)n * 4, g_ctx.stream[s]))) break;
            rc = cmsisdsp_cuda_memcpy_d2h(pIndex + f, didx, (size_t)n * 4, g_ctx.stream[s]);
        } else {
            if ((rc = cmsisdsp_cuda_cfft_mag_f32(din, dout, N, n, ifftFlag, (uint8_t)mode, g_ctx.stream[s]))) break;
            rc = cmsisdsp_cuda_memcpy_d2h((char *)pOut + f * outBytes, dout, (size_t)n * outBytes, g_ctx.stream[s]);
        }
    }
    for (int i = 0; i < NSTREAM; i++)
        if (cmsisdsp_cuda_stream_synchronize(g_ctx.stream[i])) rc = rc ? rc : -1;
    return rc ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_SUCCESS;
}
arm_status arm_cfft_mag_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pSrc, float32_t *pMag, uint32_t nFrames, uint8_t ifftFlag)
{
    return spectrum_batch(S, pSrc, pMag, 0, nFrames, ifftFlag, 0);
}
arm_status arm_cfft_mag_squared_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pSrc, float32_t *pMag, uint32_t nFrames, uint8_t ifftFlag)
{
    return spectrum_batch(S, pSrc, pMag, 0, nFrames, ifftFlag, 1);
}
arm_status arm_cfft_peak_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pSrc, float32_t *pResult, uint32_t *pIndex,
                                   uint32_t nFrames, uint8_t ifftFlag)
{
    return spectrum_batch(S, pSrc, pResult, pIndex, nFrames, ifftFlag, 2);
}
