/* arm_math.h -- umbrella header of the B200 build: the FFT path of CMSIS-DSP only. */
#ifndef ARM_MATH_H_
#define ARM_MATH_H_
#include "arm_math_types.h"
#include "dsp/transform_functions.h"
#endif
