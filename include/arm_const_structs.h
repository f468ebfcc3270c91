/*
 * arm_const_structs.h -- ready-made constant CFFT instances, usable without calling the
 * init functions (reference: Include/arm_const_structs.h:41-78,
 * Source/CommonTables/arm_const_structs.c:79-114,132-206,265-311).
 */
#ifndef ARM_CONST_STRUCTS_H
#define ARM_CONST_STRUCTS_H

#include "arm_math_types.h"
#include "arm_common_tables.h"
#include "dsp/transform_functions.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CMSISDSP_B200_DECL_SR(N)                                  \
    extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len##N;    \
    extern const arm_cfft_instance_q31 arm_cfft_sR_q31_len##N;    \
    extern const arm_cfft_instance_q15 arm_cfft_sR_q15_len##N;    \
    extern const arm_cfft_instance_f64 arm_cfft_sR_f64_len##N;
CMSISDSP_B200_FOR_EACH_LEN(CMSISDSP_B200_DECL_SR)
#undef CMSISDSP_B200_DECL_SR

/* defined (but not declared) by the reference; declared here for convenience */
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len32;
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len64;
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len128;
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len256;
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len512;
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len1024;
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len2048;
extern const arm_rfft_fast_instance_f32 arm_rfft_fast_sR_f32_len4096;

#ifdef __cplusplus
}
#endif
#endif /* ARM_CONST_STRUCTS_H */
