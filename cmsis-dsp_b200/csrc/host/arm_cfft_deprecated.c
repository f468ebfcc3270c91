/*
 * arm_cfft_deprecated.c -- the deprecated radix-4 / radix-2 instance API as adapters over the batched CFFT
 * (SURVEY 8(f) rank 4).
 *   arm_cfft_radix4_init_{f32,q31,q15}   Source/TransformFunctions/arm_cfft_radix4_init_f32.c:64-160, _q31.c, _q15.c
 *   arm_cfft_radix2_init_f32             Source/TransformFunctions/arm_cfft_radix2_init_f32.c:64-190
 *   arm_cfft_radix4_* / arm_cfft_radix2_f32 exec   arm_cfft_radix4_f32.c, arm_cfft_radix4_q31.c, arm_cfft_radix4_q15.c,
 *                                                  arm_cfft_radix2_f32.c
 * The instance is filled exactly like the reference's (4096-point twiddles read with a stride, one shared bit
 * reversal table).  The exec functions hand the frame(s) to arm_cfft_batch_* with the preset of that length: for
 * fixed point and fftLen = 4^m that IS the same computation (arm_cfft_q31.c:712-750 calls the very same
 * arm_radix4_butterfly_q31 with the per-length table, whose entries are the strided entries of the 4096-point one).
 */
#include "arm_const_structs.h"
#include "cmsisdsp_cuda.h"
#include "arm_cuda_engine.h"

arm_status arm_cuda_set_last_status(arm_status s);

static int radix4_len(uint16_t n) { return n == 16 || n == 64 || n == 256 || n == 1024 || n == 4096; }
static int radix2_len(uint16_t n) { return n >= 16 && n <= 4096 && (n & (n - 1)) == 0; }

#define DEPR_INIT(NAME, INST, TW, LENOK, ONEBY)                                                         \
    arm_status NAME(INST *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag)                 \
    {                                                                                                   \
        S->fftLen = fftLen;                                                                             \
        S->pTwiddle = TW;                                                                               \
        S->ifftFlag = ifftFlag;                                                                         \
        S->bitReverseFlag = bitReverseFlag;                                                             \
        if (!LENOK(fftLen)) return ARM_MATH_ARGUMENT_ERROR;                                             \
        S->twidCoefModifier = (uint16_t)(4096U / fftLen);                                               \
        S->bitRevFactor = (uint16_t)(4096U / fftLen);                                                   \
        S->pBitRevTable = &armBitRevTable[4096U / fftLen - 1U];                                         \
        ONEBY;                                                                                          \
        return ARM_MATH_SUCCESS;                                                                        \
    }
DEPR_INIT(arm_cfft_radix4_init_q15, arm_cfft_radix4_instance_q15, twiddleCoef_4096_q15, radix4_len, (void)0)
DEPR_INIT(arm_cfft_radix4_init_q31, arm_cfft_radix4_instance_q31, twiddleCoef_4096_q31, radix4_len, (void)0)
DEPR_INIT(arm_cfft_radix4_init_f32, arm_cfft_radix4_instance_f32, twiddleCoef, radix4_len, S->onebyfftLen = 1.0f / (float32_t)fftLen)
DEPR_INIT(arm_cfft_radix2_init_f32, arm_cfft_radix2_instance_f32, twiddleCoef, radix2_len, S->onebyfftLen = 1.0f / (float32_t)fftLen)
DEPR_INIT(arm_cfft_radix2_init_q31, arm_cfft_radix2_instance_q31, twiddleCoef_4096_q31, radix2_len, (void)0)
DEPR_INIT(arm_cfft_radix2_init_q15, arm_cfft_radix2_instance_q15, twiddleCoef_4096_q15, radix2_len, (void)0)

arm_status arm_cfft_radix4_batch_q15(const arm_cfft_radix4_instance_q15 *S, q15_t *p, uint32_t nFrames)
{
    arm_cfft_instance_q15 C;
    if (!S || !radix4_len(S->fftLen) || arm_cfft_init_q15(&C, S->fftLen) != ARM_MATH_SUCCESS) return ARM_MATH_ARGUMENT_ERROR;
    return arm_cfft_batch_q15(&C, p, nFrames, S->ifftFlag, S->bitReverseFlag);
}
arm_status arm_cfft_radix4_batch_q31(const arm_cfft_radix4_instance_q31 *S, q31_t *p, uint32_t nFrames)
{
    arm_cfft_instance_q31 C;
    if (!S || !radix4_len(S->fftLen) || arm_cfft_init_q31(&C, S->fftLen) != ARM_MATH_SUCCESS) return ARM_MATH_ARGUMENT_ERROR;
    return arm_cfft_batch_q31(&C, p, nFrames, S->ifftFlag, S->bitReverseFlag);
}
arm_status arm_cfft_batch_bitrev_order_f32(const arm_cfft_instance_f32 *S, float32_t *p, uint32_t nFrames, uint8_t ifftFlag);
/* bitReverseFlag = 0: both functions leave the spectrum in plain binary bit-reversed order (the radix-4 stages write
 * their outputs in the order a, c, b, d: arm_cfft_radix4_f32.c:261-330), measured on the compiled reference */
static arm_status depr_f32(const arm_cfft_radix4_instance_f32 *S, float32_t *p, uint32_t nFrames, int radix4)
{
    arm_cfft_instance_f32 C;
    if (!S || !(radix4 ? radix4_len(S->fftLen) : radix2_len(S->fftLen))) return ARM_MATH_ARGUMENT_ERROR;
    if (arm_cfft_init_f32(&C, S->fftLen) != ARM_MATH_SUCCESS) return ARM_MATH_ARGUMENT_ERROR;
    if (!S->bitReverseFlag) return arm_cfft_batch_bitrev_order_f32(&C, p, nFrames, S->ifftFlag);
    return arm_cfft_batch_f32(&C, p, nFrames, S->ifftFlag, 1);
}
arm_status arm_cfft_radix4_batch_f32(const arm_cfft_radix4_instance_f32 *S, float32_t *p, uint32_t nFrames) { return depr_f32(S, p, nFrames, 1); }
arm_status arm_cfft_radix2_batch_f32(const arm_cfft_radix2_instance_f32 *S, float32_t *p, uint32_t nFrames) { return depr_f32(S, p, nFrames, 0); }

void arm_cfft_radix4_q15(const arm_cfft_radix4_instance_q15 *S, q15_t *pSrc) { arm_cuda_set_last_status(arm_cfft_radix4_batch_q15(S, pSrc, 1)); }
void arm_cfft_radix4_q31(const arm_cfft_radix4_instance_q31 *S, q31_t *pSrc) { arm_cuda_set_last_status(arm_cfft_radix4_batch_q31(S, pSrc, 1)); }
void arm_cfft_radix4_f32(const arm_cfft_radix4_instance_f32 *S, float32_t *pSrc) { arm_cuda_set_last_status(arm_cfft_radix4_batch_f32(S, pSrc, 1)); }
void arm_cfft_radix2_f32(const arm_cfft_radix2_instance_f32 *S, float32_t *pSrc) { arm_cuda_set_last_status(arm_cfft_radix2_batch_f32(S, pSrc, 1)); }

/* ---- arm_cfft_radix2_q31 / _q15: log2(N) radix-2 stages with their own scaling (not the arm_cfft_q31 computation), own kernel ---- */
typedef struct { int type; uint32_t N; const void *tw; uint32_t modifier; uint8_t ifftFlag; } r2_args;
static int r2_prepare(const arm_cuda_job *job)
{
    const r2_args *a = (const r2_args *)job->self;
    return cmsisdsp_cuda_radix2_plan_upload(a->type, a->N, a->tw, a->modifier);
}
static int r2_launch(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream)
{
    const r2_args *a = (const r2_args *)job->self;
    (void)din; (void)doutB;
    return a->type == CMSISDSP_CUDA_Q31 ? cmsisdsp_cuda_cfft_radix2_q31(dout, a->N, n, a->ifftFlag, stream)
                                        : cmsisdsp_cuda_cfft_radix2_q15(dout, a->N, n, a->ifftFlag, stream);
}
static arm_status r2_batch(int type, size_t scalarBytes, uint16_t fftLen, const void *tw, uint16_t modifier, uint8_t ifftFlag, void *p, uint32_t nFrames)
{
    if (!p || !tw || !radix2_len(fftLen) || modifier == 0 || (uint32_t)modifier * fftLen > 4096U) return ARM_MATH_ARGUMENT_ERROR;
    const r2_args a = { type, fftLen, tw, modifier, ifftFlag };
    arm_cuda_job job = {0};
    job.inStride = job.inFrame = job.outStride = job.outFrame = (size_t)2 * fftLen * scalarBytes;
    job.inPlace = 1;
    job.prepare = r2_prepare;
    job.launch = r2_launch;
    job.self = &a;
    return arm_cuda_run(&job, p, p, nFrames);
}
arm_status arm_cfft_radix2_batch_q31(const arm_cfft_radix2_instance_q31 *S, q31_t *p, uint32_t nFrames)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return r2_batch(CMSISDSP_CUDA_Q31, sizeof(q31_t), S->fftLen, S->pTwiddle, S->twidCoefModifier, S->ifftFlag, p, nFrames);
}
arm_status arm_cfft_radix2_batch_q15(const arm_cfft_radix2_instance_q15 *S, q15_t *p, uint32_t nFrames)
{
    if (!S) return ARM_MATH_ARGUMENT_ERROR;
    return r2_batch(CMSISDSP_CUDA_Q15, sizeof(q15_t), S->fftLen, S->pTwiddle, S->twidCoefModifier, S->ifftFlag, p, nFrames);
}
void arm_cfft_radix2_q31(const arm_cfft_radix2_instance_q31 *S, q31_t *pSrc) { arm_cuda_set_last_status(arm_cfft_radix2_batch_q31(S, pSrc, 1)); }
void arm_cfft_radix2_q15(const arm_cfft_radix2_instance_q15 *S, q15_t *pSrc) { arm_cuda_set_last_status(arm_cfft_radix2_batch_q15(S, pSrc, 1)); }
