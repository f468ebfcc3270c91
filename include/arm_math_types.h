/*
 * arm_math_types.h -- scalar types and status codes of the CMSIS-DSP C API, restated for
 * the B200 build of the FFT path.  Same names, widths and values as the reference
 * (Include/arm_math_types.h:324-345 typedefs, :603-613 arm_status), so code written
 * against CMSIS-DSP compiles unchanged.
 */
#ifndef ARM_MATH_TYPES_H_
#define ARM_MATH_TYPES_H_

#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ARM_DSP_ATTRIBUTE
#define ARM_DSP_ATTRIBUTE
#endif
#ifndef ARM_DSP_TABLE_ATTRIBUTE
#define ARM_DSP_TABLE_ATTRIBUTE
#endif

typedef int8_t  q7_t;       /* 8-bit fractional data type in 1.7 format   */
typedef int16_t q15_t;      /* 16-bit fractional data type in 1.15 format */
typedef int32_t q31_t;      /* 32-bit fractional data type in 1.31 format */
typedef int64_t q63_t;      /* 64-bit fractional data type in 1.63 format */
typedef float   float32_t;
typedef double  float64_t;

typedef enum
{
    ARM_MATH_SUCCESS                 =  0,
    ARM_MATH_ARGUMENT_ERROR          = -1,
    ARM_MATH_LENGTH_ERROR            = -2,
    ARM_MATH_SIZE_MISMATCH           = -3,
    ARM_MATH_NANINF                  = -4,
    ARM_MATH_SINGULAR                = -5,
    ARM_MATH_TEST_FAILURE            = -6,
    ARM_MATH_DECOMPOSITION_FAILURE   = -7,
    /* B200 build only (the reference's values above are unchanged): why a batched / legacy exec call did not run */
    ARM_MATH_CUDA_NO_DEVICE          = -101,   /* no usable CUDA device (there is no CPU fallback) */
    ARM_MATH_CUDA_NO_PLAN            = -102,   /* the instance's tables are not resident on the device */
    ARM_MATH_CUDA_RUNTIME_ERROR      = -103    /* a CUDA call failed: text in cmsisdsp_cuda_last_error() */
} arm_status;

#ifdef __cplusplus
}
#endif
#endif /* ARM_MATH_TYPES_H_ */
