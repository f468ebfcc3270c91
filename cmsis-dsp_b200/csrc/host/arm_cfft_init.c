/*
 * arm_cfft_init.c -- instance initialisation, same contract as the reference:
 * fill the caller's struct from the preset of that length, ARM_MATH_ARGUMENT_ERROR for
 * any other length; no allocation, no device work (the device plan is created lazily by
 * the first exec call on each device).
 *   arm_cfft_init_f32 / _N_f32     Source/TransformFunctions/arm_cfft_init_f32.c:116-136,291-354
 *   arm_cfft_init_q31 / q15        Source/TransformFunctions/arm_cfft_init_q31.c, arm_cfft_init_q15.c:116-138,283-345
 *   arm_cfft_init_f64 / _N_f64     Source/TransformFunctions/arm_cfft_init_f64.c:58-72,205-277
 *   arm_rfft_fast_init_f32 / _N    Source/TransformFunctions/arm_rfft_fast_init_f32.c:83-99,331-371
 *   arm_rfft_init_q31 / q15 / _N   Source/TransformFunctions/arm_rfft_init_q31.c:97-127,395-470, arm_rfft_init_q15.c
 */
#include "arm_const_structs.h"

#define INIT_N(EXT, N)                                                              \
    arm_status arm_cfft_init_##N##_##EXT(arm_cfft_instance_##EXT *S)                \
    {                                                                               \
        S->fftLen = N;                                                              \
        S->pTwiddle = arm_cfft_sR_##EXT##_len##N.pTwiddle;                          \
        S->pBitRevTable = arm_cfft_sR_##EXT##_len##N.pBitRevTable;                  \
        S->bitRevLength = arm_cfft_sR_##EXT##_len##N.bitRevLength;                  \
        return ARM_MATH_SUCCESS;                                                    \
    }
#define INIT_F32(N) INIT_N(f32, N)
#define INIT_Q31(N) INIT_N(q31, N)
#define INIT_Q15(N) INIT_N(q15, N)
#define INIT_F64(N) INIT_N(f64, N)
CMSISDSP_B200_FOR_EACH_LEN(INIT_F32)
CMSISDSP_B200_FOR_EACH_LEN(INIT_Q31)
CMSISDSP_B200_FOR_EACH_LEN(INIT_Q15)
CMSISDSP_B200_FOR_EACH_LEN(INIT_F64)

#define CASE_N(EXT, N) case N##U: return arm_cfft_init_##N##_##EXT(S);
#define INIT_ANY(EXT)                                                               \
    arm_status arm_cfft_init_##EXT(arm_cfft_instance_##EXT *S, uint16_t fftLen)     \
    {                                                                               \
        switch (fftLen) {                                                           \
            CASE_N(EXT, 16) CASE_N(EXT, 32) CASE_N(EXT, 64) CASE_N(EXT, 128)        \
            CASE_N(EXT, 256) CASE_N(EXT, 512) CASE_N(EXT, 1024) CASE_N(EXT, 2048)   \
            CASE_N(EXT, 4096)                                                       \
        default: return ARM_MATH_ARGUMENT_ERROR;                                    \
        }                                                                           \
    }
INIT_ANY(f32)
INIT_ANY(q31)
INIT_ANY(q15)
INIT_ANY(f64)

#define RINIT_N(N, H)                                                               \
    arm_status arm_rfft_fast_init_##N##_f32(arm_rfft_fast_instance_f32 *S)          \
    {                                                                               \
        arm_status status;                                                          \
        if (!S) return ARM_MATH_ARGUMENT_ERROR;                                     \
        status = arm_cfft_init_##H##_f32(&(S->Sint));                               \
        if (status != ARM_MATH_SUCCESS) return status;                              \
        S->fftLenRFFT = N##U;                                                       \
        S->pTwiddleRFFT = twiddleCoef_rfft_##N;                                     \
        return ARM_MATH_SUCCESS;                                                    \
    }
RINIT_N(32, 16)
RINIT_N(64, 32)
RINIT_N(128, 64)
RINIT_N(256, 128)
RINIT_N(512, 256)
RINIT_N(1024, 512)
RINIT_N(2048, 1024)
RINIT_N(4096, 2048)

/* arm_rfft_fast_init_f64.c:44-70 (per length), :282-321 */
#define RINIT64_N(N, H)                                                             \
    arm_status arm_rfft_fast_init_##N##_f64(arm_rfft_fast_instance_f64 *S)          \
    {                                                                               \
        arm_status status;                                                          \
        if (!S) return ARM_MATH_ARGUMENT_ERROR;                                     \
        status = arm_cfft_init_##H##_f64(&(S->Sint));                               \
        if (status != ARM_MATH_SUCCESS) return status;                              \
        S->fftLenRFFT = N##U;                                                       \
        S->pTwiddleRFFT = (const float64_t *)twiddleCoefF64_rfft_##N;               \
        return ARM_MATH_SUCCESS;                                                    \
    }
RINIT64_N(32, 16)
RINIT64_N(64, 32)
RINIT64_N(128, 64)
RINIT64_N(256, 128)
RINIT64_N(512, 256)
RINIT64_N(1024, 512)
RINIT64_N(2048, 1024)
RINIT64_N(4096, 2048)

arm_status arm_rfft_fast_init_f64(arm_rfft_fast_instance_f64 *S, uint16_t fftLen)
{
    switch (fftLen) {
    case 4096U: return arm_rfft_fast_init_4096_f64(S);
    case 2048U: return arm_rfft_fast_init_2048_f64(S);
    case 1024U: return arm_rfft_fast_init_1024_f64(S);
    case 512U:  return arm_rfft_fast_init_512_f64(S);
    case 256U:  return arm_rfft_fast_init_256_f64(S);
    case 128U:  return arm_rfft_fast_init_128_f64(S);
    case 64U:   return arm_rfft_fast_init_64_f64(S);
    case 32U:   return arm_rfft_fast_init_32_f64(S);
    default:    return ARM_MATH_ARGUMENT_ERROR;
    }
}

arm_status arm_rfft_fast_init_f32(arm_rfft_fast_instance_f32 *S, uint16_t fftLen)
{
    switch (fftLen) {
    case 4096U: return arm_rfft_fast_init_4096_f32(S);
    case 2048U: return arm_rfft_fast_init_2048_f32(S);
    case 1024U: return arm_rfft_fast_init_1024_f32(S);
    case 512U:  return arm_rfft_fast_init_512_f32(S);
    case 256U:  return arm_rfft_fast_init_256_f32(S);
    case 128U:  return arm_rfft_fast_init_128_f32(S);
    case 64U:   return arm_rfft_fast_init_64_f32(S);
    case 32U:   return arm_rfft_fast_init_32_f32(S);
    default:    return ARM_MATH_ARGUMENT_ERROR;
    }
}

/* fixed-point real FFT: fftLenReal, its complex half, the stride through the 8192-entry realCoef tables */
#define RFIX_INIT_N(EXT, TAB, N, H, MOD)                                                                        \
    arm_status arm_rfft_init_##N##_##EXT(arm_rfft_instance_##EXT *S, uint32_t ifftFlagR, uint32_t bitReverseFlag) \
    {                                                                                                           \
        S->fftLenReal = (uint16_t)N;                                                                            \
        S->pTwiddleAReal = realCoefA##TAB;                                                                      \
        S->pTwiddleBReal = realCoefB##TAB;                                                                      \
        S->ifftFlagR = (uint8_t)ifftFlagR;                                                                      \
        S->bitReverseFlagR = (uint8_t)bitReverseFlag;                                                           \
        S->twidCoefRModifier = MOD##U;                                                                          \
        S->pCfft = &arm_cfft_sR_##EXT##_len##H;                                                                 \
        return ARM_MATH_SUCCESS;                                                                                \
    }
#define RFIX_ALL(EXT, TAB)                                                                                      \
    RFIX_INIT_N(EXT, TAB, 8192, 4096, 1) RFIX_INIT_N(EXT, TAB, 4096, 2048, 2) RFIX_INIT_N(EXT, TAB, 2048, 1024, 4) \
    RFIX_INIT_N(EXT, TAB, 1024, 512, 8) RFIX_INIT_N(EXT, TAB, 512, 256, 16) RFIX_INIT_N(EXT, TAB, 256, 128, 32)  \
    RFIX_INIT_N(EXT, TAB, 128, 64, 64) RFIX_INIT_N(EXT, TAB, 64, 32, 128) RFIX_INIT_N(EXT, TAB, 32, 16, 256)     \
    arm_status arm_rfft_init_##EXT(arm_rfft_instance_##EXT *S, uint32_t fftLenReal, uint32_t ifftFlagR, uint32_t bitReverseFlag) \
    {                                                                                                           \
        switch (fftLenReal) {                                                                                   \
        case 8192U: return arm_rfft_init_8192_##EXT(S, ifftFlagR, bitReverseFlag);                              \
        case 4096U: return arm_rfft_init_4096_##EXT(S, ifftFlagR, bitReverseFlag);                              \
        case 2048U: return arm_rfft_init_2048_##EXT(S, ifftFlagR, bitReverseFlag);                              \
        case 1024U: return arm_rfft_init_1024_##EXT(S, ifftFlagR, bitReverseFlag);                              \
        case 512U:  return arm_rfft_init_512_##EXT(S, ifftFlagR, bitReverseFlag);                               \
        case 256U:  return arm_rfft_init_256_##EXT(S, ifftFlagR, bitReverseFlag);                               \
        case 128U:  return arm_rfft_init_128_##EXT(S, ifftFlagR, bitReverseFlag);                               \
        case 64U:   return arm_rfft_init_64_##EXT(S, ifftFlagR, bitReverseFlag);                                \
        case 32U:   return arm_rfft_init_32_##EXT(S, ifftFlagR, bitReverseFlag);                                \
        default:    return ARM_MATH_ARGUMENT_ERROR;                                                             \
        }                                                                                                       \
    }
RFIX_ALL(q31, Q31)
RFIX_ALL(q15, Q15)
