/*
 * arm_cfft_exec.c -- exec functions of the FFT path: thin C over libcmsisdsp_cuda.
 *
 *   arm_cfft_f32 / q31 / q15   reference: arm_cfft_f32.c:1243-1298, arm_cfft_q31.c:704-755, arm_cfft_q15.c:671-722
 *   arm_cfft_f64               reference: arm_cfft_f64.c:262-312
 *   arm_rfft_fast_f32 / f64    reference: arm_rfft_fast_f32.c:675-699, arm_rfft_fast_f64.c:207-233
 *   arm_rfft_q31 / q15         reference: arm_rfft_q31.c:145-181, arm_rfft_q15.c:148-182
 *   arm_*_batch_*              B200 extension (include/dsp/transform_functions.h)
 *
 * Each function checks its arguments like the reference does (or does not), describes the work as a job
 * (arm_cuda_engine.h) and lets the engine run it: device buffers in place on their own device, host buffers fanned
 * out over the device list and streamed through staging buffers.  No CPU fallback exists: a shim failure comes back as
 * an arm_status (ARM_MATH_CUDA_* for device-side causes) and the legacy void functions record it for
 * arm_cuda_last_status().
 *
 * Like the reference, unsupported lengths in a hand-built instance make the legacy exec functions a no-op
 * (arm_cfft_f32.c:1263-1280 falls through its switch).
 */
#include "arm_math_types.h"
#include "dsp/transform_functions.h"
#include "cmsisdsp_cuda.h"
#include "arm_cuda_engine.h"

#include <stdlib.h>

static int valid_len(uint32_t n) { return n >= 16 && n <= 4096 && (n & (n - 1)) == 0; }

/* ------------------------------------------------------------------ complex FFT, in place */

typedef int (*cfft_fn)(void *, uint32_t, uint64_t, uint8_t, uint8_t, void *);
typedef struct {
    int type;
    cfft_fn fn;
    uint32_t fftLen;
    const void *tw;
    const uint16_t *br;
    uint16_t brLen;
    uint8_t ifftFlag, bitReverseFlag;
    int bitrevOrder;                 /* f32 only: leave the result in plain bit-reversed order (deprecated radix API) */
} cfft_args;

static int cfft_prepare(const arm_cuda_job *job)
{
    const cfft_args *a = (const cfft_args *)job->self;
    return cmsisdsp_cuda_plan_upload(a->type, a->fftLen, a->tw, a->br, a->brLen);
}
static int cfft_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const cfft_args *a = (const cfft_args *)job->self;
    (void)din; (void)doutB;
    if (a->bitrevOrder) return cmsisdsp_cuda_cfft_f32_bitrev_order(dout, a->fftLen, n, a->ifftFlag, stream);
    return a->fn(dout, a->fftLen, n, a->ifftFlag, a->bitReverseFlag, stream);
}

static arm_status cfft_batch(int type, cfft_fn fn, size_t scalarBytes, uint32_t fftLen, const void *tw,
                             const uint16_t *br, uint16_t brLen, void *p, uint64_t nFrames,
                             uint8_t ifftFlag, uint8_t bitReverseFlag, int bitrevOrder)
{
    if (!p || !tw) return ARM_MATH_ARGUMENT_ERROR;
    if (!valid_len(fftLen)) return ARM_MATH_ARGUMENT_ERROR;
    const cfft_args a = { type, fn, fftLen, tw, br, brLen, ifftFlag, bitReverseFlag, bitrevOrder };
    arm_cuda_job job = {0};
    job.inStride = job.inFrame = job.outStride = job.outFrame = (size_t)2 * fftLen * scalarBytes;
    job.inPlace = 1;
    job.prepare = cfft_prepare;
    job.launch = cfft_launch;
    job.self = &a;
    return arm_cuda_run(&job, p, p, nFrames);
}

arm_status arm_cfft_batch_f32(const arm_cfft_instance_f32 *S, float32_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_F32, cmsisdsp_cuda_cfft_f32, sizeof(float32_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag, 0);
}
/* used by arm_cfft_deprecated.c: arm_cfft_radix{4,2}_f32 with bitReverseFlag = 0 */
arm_status arm_cfft_batch_bitrev_order_f32(const arm_cfft_instance_f32 *S, float32_t *p, uint32_t nFrames, uint8_t ifftFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_F32, cmsisdsp_cuda_cfft_f32, sizeof(float32_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, 0, 1);
}
arm_status arm_cfft_batch_q31(const arm_cfft_instance_q31 *S, q31_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_Q31, cmsisdsp_cuda_cfft_q31, sizeof(q31_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag, 0);
}
arm_status arm_cfft_batch_q15(const arm_cfft_instance_q15 *S, q15_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_Q15, cmsisdsp_cuda_cfft_q15, sizeof(q15_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag, 0);
}
arm_status arm_cfft_batch_f64(const arm_cfft_instance_f64 *S, float64_t *p, uint32_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return cfft_batch(CMSISDSP_CUDA_F64, cmsisdsp_cuda_cfft_f64, sizeof(float64_t), S->fftLen, S->pTwiddle,
                      S->pBitRevTable, S->bitRevLength, p, nFrames, ifftFlag, bitReverseFlag, 0);
}

/* arm_cfft_f32 on frames multiplied by a real window first (re and im of sample n times pWindow[n]): the multiply of
 * arm_cmplx_mult_real_f32 fused into the transform's load */
typedef struct { cfft_args c; const float32_t *win; } cwin_args;
static int cwin_prepare(const arm_cuda_job *job)
{
    const cwin_args *a = (const cwin_args *)job->self;
    int rc = cmsisdsp_cuda_plan_upload(a->c.type, a->c.fftLen, a->c.tw, a->c.br, a->c.brLen);
    return rc ? rc : cmsisdsp_cuda_window_upload(a->c.fftLen, a->win);
}
static int cwin_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const cwin_args *a = (const cwin_args *)job->self;
    (void)din; (void)doutB;
    return cmsisdsp_cuda_cfft_window_f32(dout, a->c.fftLen, n, a->c.ifftFlag, stream);
}
arm_status arm_cfft_window_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pWindow, float32_t *p, uint32_t nFrames, uint8_t ifftFlag)
{
    if (!S || !pWindow || !p || !S->pTwiddle || !valid_len(S->fftLen)) return ARM_MATH_ARGUMENT_ERROR;
    const cwin_args a = { { CMSISDSP_CUDA_F32, cmsisdsp_cuda_cfft_f32, S->fftLen, S->pTwiddle, S->pBitRevTable, S->bitRevLength, ifftFlag, 1, 0 }, pWindow };
    arm_cuda_job job = {0};
    job.inStride = job.inFrame = job.outStride = job.outFrame = (size_t)2 * S->fftLen * sizeof(float32_t);
    job.inPlace = 1;
    job.prepare = cwin_prepare;
    job.launch = cwin_launch;
    job.self = &a;
    return arm_cuda_run(&job, p, p, nFrames);
}

/* ------------------------------------------------------------------ arm_rfft_fast_f32 / _f64 */

typedef struct {
    int f64;
    uint32_t N;                      /* real length */
    const void *cfftTw;
    const uint16_t *br;
    uint16_t brLen;
    const void *twr;
    uint8_t ifftFlag;
} rfft_args;

static int rfft_prepare(const arm_cuda_job *job)
{
    const rfft_args *a = (const rfft_args *)job->self;
    int rc = cmsisdsp_cuda_plan_upload(a->f64 ? CMSISDSP_CUDA_F64 : CMSISDSP_CUDA_F32, a->N / 2, a->cfftTw, a->br, a->brLen);
    if (rc) return rc;
    return a->f64 ? cmsisdsp_cuda_rfft_f64_plan_upload(a->N, (const double *)a->twr)
                  : cmsisdsp_cuda_rfft_plan_upload(a->N, (const float *)a->twr);
}
static int rfft_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const rfft_args *a = (const rfft_args *)job->self;
    (void)doutB;
    return a->f64 ? cmsisdsp_cuda_rfft_fast_f64(din, dout, a->N, n, a->ifftFlag, stream)
                  : cmsisdsp_cuda_rfft_fast_f32(din, dout, a->N, n, a->ifftFlag, stream);
}
/* the side effect of the reference's in-place CFFT on the input buffer (rfft_fast_f32.c:694, rfft_fast_f64.c:228) */
static int rfft_post(const arm_cuda_job *job, void *din, uint64_t n, void *stream)
{
    const rfft_args *a = (const rfft_args *)job->self;
    return a->f64 ? cmsisdsp_cuda_cfft_f64(din, a->N / 2, n, 0, 1, stream) : cmsisdsp_cuda_cfft_f32(din, a->N / 2, n, 0, 1, stream);
}

/* clobber != 0: also leave the N/2-point CFFT in p after a forward transform */
static arm_status rfft_any(int f64, uint32_t N, uint32_t cfftLen, const void *cfftTw, const uint16_t *br, uint16_t brLen,
                           const void *twr, void *p, void *pOut, uint64_t nFrames, uint8_t ifftFlag, int clobber)
{
    if (!p || !pOut || p == pOut || !twr || !cfftTw) return ARM_MATH_ARGUMENT_ERROR;
    if (N < 32 || !valid_len(N) || cfftLen != N / 2) return ARM_MATH_ARGUMENT_ERROR;
    const rfft_args a = { f64, N, cfftTw, br, brLen, twr, ifftFlag };
    const size_t frame = (size_t)N * (f64 ? sizeof(float64_t) : sizeof(float32_t));
    arm_cuda_job job = {0};
    job.inStride = job.inFrame = job.outStride = job.outFrame = frame;
    job.prepare = rfft_prepare;
    job.launch = rfft_launch;
    job.self = &a;
    if (clobber && !ifftFlag) {
        job.out2 = (char *)p;
        job.out2Stride = job.out2Frame = frame;
        job.post = rfft_post;
    }
    return arm_cuda_run(&job, p, pOut, nFrames);
}

arm_status arm_rfft_fast_batch_f32(const arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut,
                                   uint32_t nFrames, uint8_t ifftFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return rfft_any(0, S->fftLenRFFT, S->Sint.fftLen, S->Sint.pTwiddle, S->Sint.pBitRevTable, S->Sint.bitRevLength,
                    S->pTwiddleRFFT, p, pOut, nFrames, ifftFlag, 0);
}
/* forward arm_rfft_fast_f32 of frames multiplied by pWindow first: arm_mult_f32 + arm_rfft_fast_f32 as in
 * arm_mfcc_f32.c:112,137, one pass over memory; p is left untouched */
typedef struct { rfft_args r; const float32_t *win; } rwin_args;
static int rwin_prepare(const arm_cuda_job *job)
{
    const rwin_args *a = (const rwin_args *)job->self;
    arm_cuda_job inner = *job;
    inner.self = &a->r;
    int rc = rfft_prepare(&inner);
    return rc ? rc : cmsisdsp_cuda_window_upload(a->r.N, a->win);
}
static int rwin_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const rwin_args *a = (const rwin_args *)job->self;
    (void)doutB;
    return cmsisdsp_cuda_rfft_fast_window_f32(din, dout, a->r.N, n, stream);
}
arm_status arm_rfft_fast_window_batch_f32(const arm_rfft_fast_instance_f32 *S, const float32_t *pWindow, const float32_t *p,
                                          float32_t *pOut, uint32_t nFrames)
{
    if (!S || !pWindow || !p || !pOut || (const void *)p == (const void *)pOut || !S->pTwiddleRFFT || !S->Sint.pTwiddle) return ARM_MATH_ARGUMENT_ERROR;
    const uint32_t N = S->fftLenRFFT;
    if (N < 32 || !valid_len(N) || S->Sint.fftLen != N / 2) return ARM_MATH_ARGUMENT_ERROR;
    const rwin_args a = { { 0, N, S->Sint.pTwiddle, S->Sint.pBitRevTable, S->Sint.bitRevLength, S->pTwiddleRFFT, 0 }, pWindow };
    arm_cuda_job job = {0};
    job.inStride = job.inFrame = job.outStride = job.outFrame = (size_t)N * sizeof(float32_t);
    job.prepare = rwin_prepare;
    job.launch = rwin_launch;
    job.self = &a;
    return arm_cuda_run(&job, p, pOut, nFrames);
}

arm_status arm_rfft_fast_batch_f64(const arm_rfft_fast_instance_f64 *S, float64_t *p, float64_t *pOut,
                                   uint32_t nFrames, uint8_t ifftFlag)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return rfft_any(1, S->fftLenRFFT, S->Sint.fftLen, S->Sint.pTwiddle, S->Sint.pBitRevTable, S->Sint.bitRevLength,
                    S->pTwiddleRFFT, p, pOut, nFrames, ifftFlag, 0);
}

/* ---- legacy single-frame signatures ---- */

void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!valid_len(S->fftLen)) { arm_cuda_set_last_status(ARM_MATH_SUCCESS); return; }
    arm_cuda_set_last_status(arm_cfft_batch_f32(S, p1, 1, ifftFlag, bitReverseFlag));
}
void arm_cfft_f64(const arm_cfft_instance_f64 *S, float64_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!valid_len(S->fftLen)) { arm_cuda_set_last_status(ARM_MATH_SUCCESS); return; }
    arm_cuda_set_last_status(arm_cfft_batch_f64(S, p1, 1, ifftFlag, bitReverseFlag));
}
void arm_cfft_q31(const arm_cfft_instance_q31 *S, q31_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!valid_len(S->fftLen)) { arm_cuda_set_last_status(ARM_MATH_SUCCESS); return; }
    arm_cuda_set_last_status(arm_cfft_batch_q31(S, p1, 1, ifftFlag, bitReverseFlag));
}
void arm_cfft_q15(const arm_cfft_instance_q15 *S, q15_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    if (!valid_len(S->fftLen)) { arm_cuda_set_last_status(ARM_MATH_SUCCESS); return; }
    arm_cuda_set_last_status(arm_cfft_batch_q15(S, p1, 1, ifftFlag, bitReverseFlag));
}
void arm_rfft_fast_f32(const arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut, uint8_t ifftFlag)
{
    arm_cuda_set_last_status(rfft_any(0, S->fftLenRFFT, S->Sint.fftLen, S->Sint.pTwiddle, S->Sint.pBitRevTable, S->Sint.bitRevLength,
                                      S->pTwiddleRFFT, p, pOut, 1, ifftFlag, 1));
}
/* arm_rfft_fast_f64.c:207-233: sets Sint.fftLen like the reference; the forward call leaves the N/2-point CFFT in p */
void arm_rfft_fast_f64(arm_rfft_fast_instance_f64 *S, float64_t *p, float64_t *pOut, uint8_t ifftFlag)
{
    S->Sint.fftLen = S->fftLenRFFT / 2;
    if (S->fftLenRFFT < 32 || !valid_len(S->fftLenRFFT)) { arm_cuda_set_last_status(ARM_MATH_SUCCESS); return; }
    arm_cuda_set_last_status(rfft_any(1, S->fftLenRFFT, S->Sint.fftLen, S->Sint.pTwiddle, S->Sint.pBitRevTable, S->Sint.bitRevLength,
                                      S->pTwiddleRFFT, p, pOut, 1, ifftFlag, 1));
}

/* ------------------------------------------------------------------ fixed-point real FFT */

typedef struct {
    int type;
    uint32_t N;
    uint8_t ifftFlagR, bitReverseFlagR;
    uint32_t modifier;
    const void *coefA, *coefB, *cfftTw;
    const uint16_t *br;
    uint16_t brLen;
} rfix_args;

static int rfix_prepare(const arm_cuda_job *job)
{
    const rfix_args *a = (const rfix_args *)job->self;
    int rc = cmsisdsp_cuda_plan_upload(a->type, a->N / 2, a->cfftTw, a->br, a->brLen);
    if (rc) return rc;
    return cmsisdsp_cuda_rfft_fix_plan_upload(a->type, a->N, a->coefA, a->coefB, a->modifier);
}
static int rfix_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const rfix_args *a = (const rfix_args *)job->self;
    (void)doutB;
    return a->type == CMSISDSP_CUDA_Q31 ? cmsisdsp_cuda_rfft_q31(din, dout, a->N, n, a->ifftFlagR, a->bitReverseFlagR, stream)
                                        : cmsisdsp_cuda_rfft_q15(din, dout, a->N, n, a->ifftFlagR, a->bitReverseFlagR, stream);
}
/* the side effect of the reference's in-place CFFT on the source buffer (arm_rfft_q31.c:174, arm_rfft_q15.c:176) */
static int rfix_post(const arm_cuda_job *job, void *din, uint64_t n, void *stream)
{
    const rfix_args *a = (const rfix_args *)job->self;
    return a->type == CMSISDSP_CUDA_Q31 ? cmsisdsp_cuda_cfft_q31(din, a->N / 2, n, 0, a->bitReverseFlagR, stream)
                                        : cmsisdsp_cuda_cfft_q15(din, a->N / 2, n, 0, a->bitReverseFlagR, stream);
}

static arm_status rfix_batch(int type, size_t scalarBytes, uint32_t N, uint8_t ifftFlagR,
                             uint8_t bitReverseFlagR, uint32_t modifier, const void *coefA, const void *coefB,
                             const void *cfftTw, const uint16_t *br, uint16_t brLen, uint32_t cfftLen,
                             void *pSrc, void *pDst, uint64_t nFrames, int clobber)
{
    if (!pSrc || !pDst || pSrc == pDst || !coefA || !coefB || !cfftTw) return ARM_MATH_ARGUMENT_ERROR;
    if (N < 32 || N > 8192 || (N & (N - 1)) != 0 || cfftLen != N / 2) return ARM_MATH_ARGUMENT_ERROR;
    const rfix_args a = { type, N, ifftFlagR, (uint8_t)(bitReverseFlagR ? 1 : 0), modifier, coefA, coefB, cfftTw, br, brLen };
    arm_cuda_job job = {0};
    if (ifftFlagR) {
        /* frames 2N scalars apart, of which bins 0..N/2 (N + 2 scalars) are read: arm_rifft_input_buffer_size */
        job.inStride = (size_t)2 * N * scalarBytes;
        job.inFrame = (size_t)(N + 2) * scalarBytes;
        job.outStride = job.outFrame = (size_t)N * scalarBytes;
    } else {
        job.inStride = job.inFrame = (size_t)N * scalarBytes;
        job.outStride = job.outFrame = (size_t)2 * N * scalarBytes;
    }
    job.prepare = rfix_prepare;
    job.launch = rfix_launch;
    job.self = &a;
    if (clobber && !ifftFlagR) {
        job.out2 = (char *)pSrc;
        job.out2Stride = job.out2Frame = job.inStride;
        job.post = rfix_post;
    }
    return arm_cuda_run(&job, pSrc, pDst, nFrames);
}

static arm_status rfix_q31(const arm_rfft_instance_q31 *S, q31_t *pSrc, q31_t *pDst, uint64_t nFrames, int clobber)
{
    if (!S || !S->pCfft) return ARM_MATH_ARGUMENT_ERROR;
    return rfix_batch(CMSISDSP_CUDA_Q31, sizeof(q31_t), S->fftLenReal, S->ifftFlagR,
                      S->bitReverseFlagR, S->twidCoefRModifier, S->pTwiddleAReal, S->pTwiddleBReal, S->pCfft->pTwiddle,
                      S->pCfft->pBitRevTable, S->pCfft->bitRevLength, S->pCfft->fftLen, pSrc, pDst, nFrames, clobber);
}
static arm_status rfix_q15(const arm_rfft_instance_q15 *S, q15_t *pSrc, q15_t *pDst, uint64_t nFrames, int clobber)
{
    if (!S || !S->pCfft) return ARM_MATH_ARGUMENT_ERROR;
    return rfix_batch(CMSISDSP_CUDA_Q15, sizeof(q15_t), S->fftLenReal, S->ifftFlagR,
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
void arm_rfft_q31(const arm_rfft_instance_q31 *S, q31_t *pSrc, q31_t *pDst) { arm_cuda_set_last_status(rfix_q31(S, pSrc, pDst, 1, 1)); }
void arm_rfft_q15(const arm_rfft_instance_q15 *S, q15_t *pSrc, q15_t *pDst) { arm_cuda_set_last_status(rfix_q15(S, pSrc, pDst, 1, 1)); }

/* ---- arm_cfft_f32 fused with arm_cmplx_mag[_squared]_f32 (mode 0 / 1) or with arm_cmplx_mag_f32 + arm_max_f32 (mode 2) ---- */

typedef struct {
    uint32_t N;
    const void *tw;
    const uint16_t *br;
    uint16_t brLen;
    uint8_t ifftFlag;
    int mode;
} spec_args;

static int spec_prepare(const arm_cuda_job *job)
{
    const spec_args *a = (const spec_args *)job->self;
    return cmsisdsp_cuda_plan_upload(CMSISDSP_CUDA_F32, a->N, a->tw, a->br, a->brLen);
}
static int spec_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const spec_args *a = (const spec_args *)job->self;
    if (a->mode == 2) return cmsisdsp_cuda_cfft_peak_f32(din, dout, doutB, a->N, n, a->ifftFlag, stream);
    return cmsisdsp_cuda_cfft_mag_f32(din, dout, a->N, n, a->ifftFlag, (uint8_t)a->mode, stream);
}

static arm_status spectrum_batch(const arm_cfft_instance_f32 *S, const float32_t *pSrc, void *pOut, uint32_t *pIndex,
                                 uint64_t nFrames, uint8_t ifftFlag, int mode)
{
    if (!S || !pSrc || !pOut || (mode == 2 && !pIndex) || !S->pTwiddle || (const void *)pSrc == pOut) return ARM_MATH_ARGUMENT_ERROR;
    const uint32_t N = S->fftLen;
    if (!valid_len(N)) return ARM_MATH_ARGUMENT_ERROR;
    const spec_args a = { N, S->pTwiddle, S->pBitRevTable, S->bitRevLength, ifftFlag, mode };
    arm_cuda_job job = {0};
    job.inStride = job.inFrame = (size_t)2 * N * sizeof(float32_t);
    job.outStride = job.outFrame = (mode == 2) ? sizeof(float32_t) : (size_t)N * sizeof(float32_t);
    if (mode == 2) {
        job.outB = (char *)pIndex;
        job.outBStride = job.outBFrame = sizeof(uint32_t);
    }
    job.prepare = spec_prepare;
    job.launch = spec_launch;
    job.self = &a;
    return arm_cuda_run(&job, pSrc, pOut, nFrames);
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
