/*
 * arm_const_structs.c -- constant preset instances (reference:
 * Source/CommonTables/arm_const_structs.c:39-73 f64, :79-114 f32, :132-166 q31, :172-206 q15, :265-311 rfft).
 * Table data comes from the build-time generator (csrc/tables/gen_tables.c).
 */
#include "arm_const_structs.h"

#define SR_F32(N) \
    const arm_cfft_instance_f32 arm_cfft_sR_f32_len##N = { N, twiddleCoef_##N, armBitRevIndexTable##N, ARMBITREVINDEXTABLE_##N##_TABLE_LENGTH };
#define SR_Q31(N) \
    const arm_cfft_instance_q31 arm_cfft_sR_q31_len##N = { N, twiddleCoef_##N##_q31, armBitRevIndexTable_fixed_##N, ARMBITREVINDEXTABLE_FIXED_##N##_TABLE_LENGTH };
#define SR_Q15(N) \
    const arm_cfft_instance_q15 arm_cfft_sR_q15_len##N = { N, twiddleCoef_##N##_q15, armBitRevIndexTable_fixed_##N, ARMBITREVINDEXTABLE_FIXED_##N##_TABLE_LENGTH };

#define SR_F64(N) \
    const arm_cfft_instance_f64 arm_cfft_sR_f64_len##N = { N, (const float64_t *)twiddleCoefF64_##N, armBitRevIndexTableF64_##N, ARMBITREVINDEXTABLEF64_##N##_TABLE_LENGTH };

CMSISDSP_B200_FOR_EACH_LEN(SR_F32)
CMSISDSP_B200_FOR_EACH_LEN(SR_F64)
CMSISDSP_B200_FOR_EACH_LEN(SR_Q31)
CMSISDSP_B200_FOR_EACH_LEN(SR_Q15)

#define SR_RFFT(N, H) \
    const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len##N = { \
        { H, twiddleCoef_##H, armBitRevIndexTable##H, ARMBITREVINDEXTABLE_##H##_TABLE_LENGTH }, N, twiddleCoef_rfft_##N };
SR_RFFT(32, 16)
SR_RFFT(64, 32)
SR_RFFT(128, 64)
SR_RFFT(256, 128)
SR_RFFT(512, 256)
SR_RFFT(1024, 512)
SR_RFFT(2048, 1024)
SR_RFFT(4096, 2048)
