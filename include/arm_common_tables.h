/*
 * arm_common_tables.h -- declarations of the constant tables of the FFT path, under the
 * reference's symbol names and lengths (Include/arm_common_tables.h:39-236).  The data is
 * generated at build time by cmsis-dsp_b200/csrc/tables/gen_tables.c and is value-identical
 * to Source/CommonTables/arm_common_tables.c (checked by tests/test_tables.py).
 */
#ifndef ARM_COMMON_TABLES_H
#define ARM_COMMON_TABLES_H

#include "arm_math_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CMSISDSP_B200_FOR_EACH_LEN(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)

#define CMSISDSP_B200_DECL_TW(N)                       \
    extern const float32_t twiddleCoef_##N[(N) * 2];   \
    extern const q31_t twiddleCoef_##N##_q31[(N) * 3 / 2]; \
    extern const q15_t twiddleCoef_##N##_q15[(N) * 3 / 2];
CMSISDSP_B200_FOR_EACH_LEN(CMSISDSP_B200_DECL_TW)
#undef CMSISDSP_B200_DECL_TW

/* f64 complex FFT (Include/arm_common_tables.h:43-59,153-178): twiddles stored as IEEE-754 bit patterns like the
 * reference's; generated values are within 1 ulp of the reference's literals (see gen_tables.c).  The bit reversal
 * tables are the fixed-point (binary) swap lists, as in the reference. */
#define CMSISDSP_B200_DECL_TW64(N)                       \
    extern const uint64_t twiddleCoefF64_##N[(N) * 2];   \
    extern const uint16_t armBitRevIndexTableF64_##N[];
CMSISDSP_B200_FOR_EACH_LEN(CMSISDSP_B200_DECL_TW64)
#undef CMSISDSP_B200_DECL_TW64
extern const uint64_t twiddleCoefF64_rfft_32[32];
extern const uint64_t twiddleCoefF64_rfft_64[64];
extern const uint64_t twiddleCoefF64_rfft_128[128];
extern const uint64_t twiddleCoefF64_rfft_256[256];
extern const uint64_t twiddleCoefF64_rfft_512[512];
extern const uint64_t twiddleCoefF64_rfft_1024[1024];
extern const uint64_t twiddleCoefF64_rfft_2048[2048];
extern const uint64_t twiddleCoefF64_rfft_4096[4096];
#define ARMBITREVINDEXTABLEF64_16_TABLE_LENGTH   ((uint16_t)12)
#define ARMBITREVINDEXTABLEF64_32_TABLE_LENGTH   ((uint16_t)24)
#define ARMBITREVINDEXTABLEF64_64_TABLE_LENGTH   ((uint16_t)56)
#define ARMBITREVINDEXTABLEF64_128_TABLE_LENGTH  ((uint16_t)112)
#define ARMBITREVINDEXTABLEF64_256_TABLE_LENGTH  ((uint16_t)240)
#define ARMBITREVINDEXTABLEF64_512_TABLE_LENGTH  ((uint16_t)480)
#define ARMBITREVINDEXTABLEF64_1024_TABLE_LENGTH ((uint16_t)992)
#define ARMBITREVINDEXTABLEF64_2048_TABLE_LENGTH ((uint16_t)1984)
#define ARMBITREVINDEXTABLEF64_4096_TABLE_LENGTH ((uint16_t)4032)

extern const float32_t twiddleCoef_rfft_32[32];
extern const float32_t twiddleCoef_rfft_64[64];
extern const float32_t twiddleCoef_rfft_128[128];
extern const float32_t twiddleCoef_rfft_256[256];
extern const float32_t twiddleCoef_rfft_512[512];
extern const float32_t twiddleCoef_rfft_1024[1024];
extern const float32_t twiddleCoef_rfft_2048[2048];
extern const float32_t twiddleCoef_rfft_4096[4096];

/* floating-point bit reversal tables (mixed radix-8/4/2 digit reversal) */
#define ARMBITREVINDEXTABLE_16_TABLE_LENGTH   ((uint16_t)20)
#define ARMBITREVINDEXTABLE_32_TABLE_LENGTH   ((uint16_t)48)
#define ARMBITREVINDEXTABLE_64_TABLE_LENGTH   ((uint16_t)56)
#define ARMBITREVINDEXTABLE_128_TABLE_LENGTH  ((uint16_t)208)
#define ARMBITREVINDEXTABLE_256_TABLE_LENGTH  ((uint16_t)440)
#define ARMBITREVINDEXTABLE_512_TABLE_LENGTH  ((uint16_t)448)
#define ARMBITREVINDEXTABLE_1024_TABLE_LENGTH ((uint16_t)1800)
#define ARMBITREVINDEXTABLE_2048_TABLE_LENGTH ((uint16_t)3808)
#define ARMBITREVINDEXTABLE_4096_TABLE_LENGTH ((uint16_t)4032)
extern const uint16_t armBitRevIndexTable16[ARMBITREVINDEXTABLE_16_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable32[ARMBITREVINDEXTABLE_32_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable64[ARMBITREVINDEXTABLE_64_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable128[ARMBITREVINDEXTABLE_128_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable256[ARMBITREVINDEXTABLE_256_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable512[ARMBITREVINDEXTABLE_512_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable1024[ARMBITREVINDEXTABLE_1024_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable2048[ARMBITREVINDEXTABLE_2048_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable4096[ARMBITREVINDEXTABLE_4096_TABLE_LENGTH];

/* fixed-point bit reversal tables (binary bit reversal) */
#define ARMBITREVINDEXTABLE_FIXED_16_TABLE_LENGTH   ((uint16_t)12)
#define ARMBITREVINDEXTABLE_FIXED_32_TABLE_LENGTH   ((uint16_t)24)
#define ARMBITREVINDEXTABLE_FIXED_64_TABLE_LENGTH   ((uint16_t)56)
#define ARMBITREVINDEXTABLE_FIXED_128_TABLE_LENGTH  ((uint16_t)112)
#define ARMBITREVINDEXTABLE_FIXED_256_TABLE_LENGTH  ((uint16_t)240)
#define ARMBITREVINDEXTABLE_FIXED_512_TABLE_LENGTH  ((uint16_t)480)
#define ARMBITREVINDEXTABLE_FIXED_1024_TABLE_LENGTH ((uint16_t)992)
#define ARMBITREVINDEXTABLE_FIXED_2048_TABLE_LENGTH ((uint16_t)1984)
#define ARMBITREVINDEXTABLE_FIXED_4096_TABLE_LENGTH ((uint16_t)4032)
extern const uint16_t armBitRevIndexTable_fixed_16[ARMBITREVINDEXTABLE_FIXED_16_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_32[ARMBITREVINDEXTABLE_FIXED_32_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_64[ARMBITREVINDEXTABLE_FIXED_64_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_128[ARMBITREVINDEXTABLE_FIXED_128_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_256[ARMBITREVINDEXTABLE_FIXED_256_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_512[ARMBITREVINDEXTABLE_FIXED_512_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_1024[ARMBITREVINDEXTABLE_FIXED_1024_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_2048[ARMBITREVINDEXTABLE_FIXED_2048_TABLE_LENGTH];
extern const uint16_t armBitRevIndexTable_fixed_4096[ARMBITREVINDEXTABLE_FIXED_4096_TABLE_LENGTH];

/* deprecated radix-2 / radix-4 instance API: one bit reversal table for every length, the 4096-point twiddles
 * read with a stride (Include/arm_common_tables.h:52,78) */
extern const uint16_t armBitRevTable[1024];
#define twiddleCoef twiddleCoef_4096

/* split-stage coefficients of arm_rfft_q31 / arm_rfft_q15 (Include/arm_common_tables.h:241-245) */
extern const q31_t realCoefAQ31[8192];
extern const q31_t realCoefBQ31[8192];
extern const q15_t realCoefAQ15[8192];
extern const q15_t realCoefBQ15[8192];

#ifdef __cplusplus
}
#endif
#endif /* ARM_COMMON_TABLES_H */
