/*
 * dsp/transform_functions.h -- the FFT-path slice of the CMSIS-DSP transform API, B200 build.
 *
 * Instance structs, init and exec prototypes are those of the reference's generic
 * (non-Neon, non-MVE) branch, field for field and argument for argument
 * (Include/dsp/transform_functions.h:282-331 q15, :347-394 q31, :410-460 f32, :813-849 rfft_fast,
 * :508-617 rfft_q15, :655-760 rfft_q31),
 * so existing callers re-link without source changes.  The `*_batch_*` functions at the end
 * are the B200 extension: the same transform over nFrames contiguous frames in one call.
 *
 * Execution model: the exec functions run on the current CUDA device through
 * libcmsisdsp_cuda (include/cmsisdsp_cuda.h).  Data pointers may be host pointers (the
 * call stages them through device memory and returns when the result is back in the
 * buffer) or device pointers (the call enqueues on the library's stream and returns after
 * completion).  There is no CPU fallback: without a CUDA device the batched functions
 * return ARM_MATH_CUDA_NO_DEVICE and the legacy void functions latch the error in
 * arm_cuda_last_status().
 */
#ifndef TRANSFORM_FUNCTIONS_H_
#define TRANSFORM_FUNCTIONS_H_

#include "arm_math_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- q15 CFFT */
typedef struct
{
          uint16_t  fftLen;          /* length of the FFT */
    const q15_t    *pTwiddle;        /* points to the twiddle factor table */
    const uint16_t *pBitRevTable;    /* points to the bit reversal table */
          uint16_t  bitRevLength;    /* bit reversal table length */
} arm_cfft_instance_q15;

arm_status arm_cfft_init_4096_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_2048_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_1024_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_512_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_256_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_128_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_64_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_32_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_16_q15(arm_cfft_instance_q15 *S);
arm_status arm_cfft_init_q15(arm_cfft_instance_q15 *S, uint16_t fftLen);
void arm_cfft_q15(const arm_cfft_instance_q15 *S, q15_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag);

/* ---------------------------------------------------------------- q31 CFFT */
typedef struct
{
          uint16_t  fftLen;
    const q31_t    *pTwiddle;
    const uint16_t *pBitRevTable;
          uint16_t  bitRevLength;
} arm_cfft_instance_q31;

arm_status arm_cfft_init_4096_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_2048_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_1024_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_512_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_256_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_128_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_64_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_32_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_16_q31(arm_cfft_instance_q31 *S);
arm_status arm_cfft_init_q31(arm_cfft_instance_q31 *S, uint16_t fftLen);
void arm_cfft_q31(const arm_cfft_instance_q31 *S, q31_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag);

/* ---------------------------------------------------------------- f32 CFFT */
typedef struct
{
          uint16_t   fftLen;
    const float32_t *pTwiddle;
    const uint16_t  *pBitRevTable;
          uint16_t   bitRevLength;
} arm_cfft_instance_f32;

arm_status arm_cfft_init_4096_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_2048_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_1024_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_512_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_256_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_128_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_64_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_32_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_16_f32(arm_cfft_instance_f32 *S);
arm_status arm_cfft_init_f32(arm_cfft_instance_f32 *S, uint16_t fftLen);
void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag);

/* ---------------------------------------------------------------- f64 CFFT
 * (Include/dsp/transform_functions.h:466-493; arm_cfft_f64.c:262-312) */
typedef struct
{
          uint16_t   fftLen;
    const float64_t *pTwiddle;
    const uint16_t  *pBitRevTable;
          uint16_t   bitRevLength;
} arm_cfft_instance_f64;

arm_status arm_cfft_init_4096_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_2048_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_1024_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_512_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_256_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_128_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_64_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_32_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_16_f64(arm_cfft_instance_f64 *S);
arm_status arm_cfft_init_f64(arm_cfft_instance_f64 *S, uint16_t fftLen);
void arm_cfft_f64(const arm_cfft_instance_f64 *S, float64_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag);

/* ---------------------------------------------------------------- f64 fast RFFT
 * (Include/dsp/transform_functions.h:770-794; arm_rfft_fast_f64.c:207-233).  S is not const, as in the reference. */
typedef struct
{
          arm_cfft_instance_f64 Sint;     /* internal CFFT structure (length fftLenRFFT/2) */
          uint16_t   fftLenRFFT;          /* length of the real sequence */
    const float64_t *pTwiddleRFFT;        /* twiddle factors of the real stage */
} arm_rfft_fast_instance_f64;

arm_status arm_rfft_fast_init_32_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_64_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_128_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_256_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_512_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_1024_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_2048_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_4096_f64(arm_rfft_fast_instance_f64 *S);
arm_status arm_rfft_fast_init_f64(arm_rfft_fast_instance_f64 *S, uint16_t fftLen);
void arm_rfft_fast_f64(arm_rfft_fast_instance_f64 *S, float64_t *p, float64_t *pOut, uint8_t ifftFlag);

/* ---------------------------------------------------------------- f32 fast RFFT */
typedef struct
{
          arm_cfft_instance_f32 Sint;     /* internal CFFT structure (length fftLenRFFT/2) */
          uint16_t   fftLenRFFT;          /* length of the real sequence */
    const float32_t *pTwiddleRFFT;        /* twiddle factors of the real stage */
} arm_rfft_fast_instance_f32;

arm_status arm_rfft_fast_init_32_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_64_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_128_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_256_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_512_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_1024_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_2048_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_4096_f32(arm_rfft_fast_instance_f32 *S);
arm_status arm_rfft_fast_init_f32(arm_rfft_fast_instance_f32 *S, uint16_t fftLen);
void arm_rfft_fast_f32(const arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut, uint8_t ifftFlag);

/* ---------------------------------------------------------------- deprecated radix-4 / radix-2 instance API
 * (Include/dsp/transform_functions.h:107-266; direction and bit-reversal flag live in the instance).  Thin adapters
 * over the same kernels: the reference's arm_cfft_radix4_q31 / _q15 ARE arm_cfft_q31 / _q15 for fftLen = 4^m (same
 * butterflies, same table values read with a stride: bit-identical, checked against the compiled reference), and the
 * f32 variants compute the same DFT (relative RMS vs the reference's radix-4 / radix-2 code <= 2e-6).
 * Lengths: radix-4 16, 64, 256, 1024, 4096; radix-2 f32 16..4096.  The f32 adapters need bitReverseFlag = 1 (the raw
 * order of a radix-4 / radix-2 pass differs from arm_cfft_f32's).  The fixed-point radix-2 functions are a different
 * algorithm (other per-stage scaling) and are not provided. */
typedef struct
{
          uint16_t  fftLen;
          uint8_t   ifftFlag;
          uint8_t   bitReverseFlag;
    const q15_t    *pTwiddle;            /* twiddleCoef_4096_q15, read with stride twidCoefModifier */
    const uint16_t *pBitRevTable;        /* &armBitRevTable[bitRevFactor - 1] */
          uint16_t  twidCoefModifier;
          uint16_t  bitRevFactor;
} arm_cfft_radix4_instance_q15;
typedef struct
{
          uint16_t  fftLen;
          uint8_t   ifftFlag;
          uint8_t   bitReverseFlag;
    const q31_t    *pTwiddle;
    const uint16_t *pBitRevTable;
          uint16_t  twidCoefModifier;
          uint16_t  bitRevFactor;
} arm_cfft_radix4_instance_q31;
typedef struct
{
          uint16_t   fftLen;
          uint8_t    ifftFlag;
          uint8_t    bitReverseFlag;
    const float32_t *pTwiddle;
    const uint16_t  *pBitRevTable;
          uint16_t   twidCoefModifier;
          uint16_t   bitRevFactor;
          float32_t  onebyfftLen;
} arm_cfft_radix4_instance_f32;
typedef arm_cfft_radix4_instance_f32 arm_cfft_radix2_instance_f32;     /* same fields (transform_functions.h:215-225) */
typedef arm_cfft_radix4_instance_q31 arm_cfft_radix2_instance_q31;     /* same fields (transform_functions.h:163-172) */
typedef arm_cfft_radix4_instance_q15 arm_cfft_radix2_instance_q15;     /* same fields (transform_functions.h:110-119) */

arm_status arm_cfft_radix4_init_q15(arm_cfft_radix4_instance_q15 *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag);
arm_status arm_cfft_radix4_init_q31(arm_cfft_radix4_instance_q31 *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag);
arm_status arm_cfft_radix4_init_f32(arm_cfft_radix4_instance_f32 *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag);
arm_status arm_cfft_radix2_init_f32(arm_cfft_radix2_instance_f32 *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag);
void arm_cfft_radix4_q15(const arm_cfft_radix4_instance_q15 *S, q15_t *pSrc);
void arm_cfft_radix4_q31(const arm_cfft_radix4_instance_q31 *S, q31_t *pSrc);
void arm_cfft_radix4_f32(const arm_cfft_radix4_instance_f32 *S, float32_t *pSrc);
void arm_cfft_radix2_f32(const arm_cfft_radix2_instance_f32 *S, float32_t *pSrc);
/* fixed-point radix-2 (arm_cfft_radix2_q31.c:62-81, arm_cfft_radix2_q15.c:62-78): their own algorithm and scaling, own
 * kernel; fftLen 16..4096, every power of two; the result is always in natural order (the reference bit-reverses
 * whatever bitReverseFlag says) */
arm_status arm_cfft_radix2_init_q31(arm_cfft_radix2_instance_q31 *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag);
arm_status arm_cfft_radix2_init_q15(arm_cfft_radix2_instance_q15 *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag);
void arm_cfft_radix2_q31(const arm_cfft_radix2_instance_q31 *S, q31_t *pSrc);
void arm_cfft_radix2_q15(const arm_cfft_radix2_instance_q15 *S, q15_t *pSrc);
/* B200 extension: nFrames contiguous frames */
arm_status arm_cfft_radix4_batch_q15(const arm_cfft_radix4_instance_q15 *S, q15_t *p, uint32_t nFrames);
arm_status arm_cfft_radix4_batch_q31(const arm_cfft_radix4_instance_q31 *S, q31_t *p, uint32_t nFrames);
arm_status arm_cfft_radix4_batch_f32(const arm_cfft_radix4_instance_f32 *S, float32_t *p, uint32_t nFrames);
arm_status arm_cfft_radix2_batch_f32(const arm_cfft_radix2_instance_f32 *S, float32_t *p, uint32_t nFrames);
arm_status arm_cfft_radix2_batch_q31(const arm_cfft_radix2_instance_q31 *S, q31_t *p, uint32_t nFrames);
arm_status arm_cfft_radix2_batch_q15(const arm_cfft_radix2_instance_q15 *S, q15_t *p, uint32_t nFrames);

/* ---------------------------------------------------------------- q15 / q31 RFFT
 * Instance, init and exec are the reference's generic (non-Neon, non-MVE) branch
 * (Include/dsp/transform_functions.h:508-521,566-617 q15, :655-668,694-745 q31).  fftLenReal in {32..8192}.
 * Buffers as in the reference: forward reads fftLenReal scalars and writes 2*fftLenReal (all fftLenReal
 * complex bins, conjugate half included); inverse reads bins 0..fftLenReal/2 and writes fftLenReal scalars.
 * bitReverseFlagR is handed to the complex transform inside, as in the reference (arm_rfft_q31.c:164,173): with 0 that
 * transform's result stays in bit-reversed order. */
typedef struct
{
          uint32_t fftLenReal;                /* length of the real FFT */
          uint8_t  ifftFlagR;                 /* 0: forward, 1: inverse */
          uint8_t  bitReverseFlagR;           /* 1: output in natural order */
          uint32_t twidCoefRModifier;         /* stride through the 8192-entry coefficient tables */
    const q15_t   *pTwiddleAReal;             /* realCoefAQ15 */
    const q15_t   *pTwiddleBReal;             /* realCoefBQ15 */
    const arm_cfft_instance_q15 *pCfft;       /* complex FFT instance of length fftLenReal/2 */
} arm_rfft_instance_q15;

typedef struct
{
          uint32_t fftLenReal;
          uint8_t  ifftFlagR;
          uint8_t  bitReverseFlagR;
          uint32_t twidCoefRModifier;
    const q31_t   *pTwiddleAReal;             /* realCoefAQ31 */
    const q31_t   *pTwiddleBReal;             /* realCoefBQ31 */
    const arm_cfft_instance_q31 *pCfft;
} arm_rfft_instance_q31;

#define CMSISDSP_B200_DECL_RFFT_FIX(N)                                                                             \
    arm_status arm_rfft_init_##N##_q15(arm_rfft_instance_q15 *S, uint32_t ifftFlagR, uint32_t bitReverseFlag);   \
    arm_status arm_rfft_init_##N##_q31(arm_rfft_instance_q31 *S, uint32_t ifftFlagR, uint32_t bitReverseFlag);
CMSISDSP_B200_DECL_RFFT_FIX(32) CMSISDSP_B200_DECL_RFFT_FIX(64) CMSISDSP_B200_DECL_RFFT_FIX(128)
CMSISDSP_B200_DECL_RFFT_FIX(256) CMSISDSP_B200_DECL_RFFT_FIX(512) CMSISDSP_B200_DECL_RFFT_FIX(1024)
CMSISDSP_B200_DECL_RFFT_FIX(2048) CMSISDSP_B200_DECL_RFFT_FIX(4096) CMSISDSP_B200_DECL_RFFT_FIX(8192)
#undef CMSISDSP_B200_DECL_RFFT_FIX
arm_status arm_rfft_init_q15(arm_rfft_instance_q15 *S, uint32_t fftLenReal, uint32_t ifftFlagR, uint32_t bitReverseFlag);
arm_status arm_rfft_init_q31(arm_rfft_instance_q31 *S, uint32_t fftLenReal, uint32_t ifftFlagR, uint32_t bitReverseFlag);
/* pSrc is modified by the forward transform like in the reference (it then holds the fftLenReal/2-point CFFT) */
void arm_rfft_q15(const arm_rfft_instance_q15 *S, q15_t *pSrc, q15_t *pDst);
void arm_rfft_q31(const arm_rfft_instance_q31 *S, q31_t *pSrc, q31_t *pDst);

/* ---------------------------------------------------------------- f32 MFCC (RFFT based)
 * Instance and init are the reference's (Include/dsp/transform_functions.h:856-998, generic branch);
 * supported fftLen here: 256, 512, 1024, 2048, 4096. */
typedef struct
{
     const float32_t *dctCoefs;       /* DCT matrix, nbDctOutputs x nbMelFilters */
     const float32_t *filterCoefs;    /* packed mel filter taps */
     const float32_t *windowCoefs;    /* fftLen window coefficients */
     const uint32_t  *filterPos;      /* first spectrum bin of each mel filter */
     const uint32_t  *filterLengths;  /* number of taps of each mel filter */
     uint32_t fftLen;
     uint32_t nbMelFilters;
     uint32_t nbDctOutputs;
     arm_rfft_fast_instance_f32 rfft;
} arm_mfcc_instance_f32;

arm_status arm_mfcc_init_f32(arm_mfcc_instance_f32 *S, uint32_t fftLen, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_32_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_64_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_128_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_256_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_512_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_1024_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_2048_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
arm_status arm_mfcc_init_4096_f32(arm_mfcc_instance_f32 *S, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                             const float32_t *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                             const float32_t *filterCoefs, const float32_t *windowCoefs);
/* pSrc (fftLen samples) is destroyed like in the reference; pTmp (2*fftLen floats there) is not used */
void arm_mfcc_f32(const arm_mfcc_instance_f32 *S, float32_t *pSrc, float32_t *pDst, float32_t *pTmp);
/* B200 extension: nFrames frames taken every `hop` samples from pSrc (hop = fftLen: back to back; smaller:
 * overlapping; must be even), nbDctOutputs coefficients per frame written to pDst.  pSrc is left untouched.
 * Host or device pointers. */
arm_status arm_mfcc_batch_f32(const arm_mfcc_instance_f32 *S, const float32_t *pSrc, uint32_t hop,
                              float32_t *pDst, uint32_t nFrames);

/* ---------------------------------------------------------------- B200 extension: batches
 *
 * nFrames frames stored back to back (frame stride 2*fftLen scalars for CFFT, fftLenRFFT
 * floats for RFFT), transformed exactly as nFrames calls of the single-frame function.
 * Returns ARM_MATH_SUCCESS; ARM_MATH_ARGUMENT_ERROR for a NULL / unsupported instance or buffer; ARM_MATH_CUDA_NO_DEVICE,
 * ARM_MATH_CUDA_NO_PLAN or ARM_MATH_CUDA_RUNTIME_ERROR (arm_math_types.h) when the device side failed (text in
 * cmsisdsp_cuda_last_error()).
 * arm_rfft_fast_batch_f32 leaves p untouched in both directions (the single-frame forward
 * call keeps the reference's documented side effect: p then holds the N/2-point CFFT). */
arm_status arm_cfft_batch_f32(const arm_cfft_instance_f32 *S, float32_t *p, uint32_t nFrames,
                              uint8_t ifftFlag, uint8_t bitReverseFlag);
arm_status arm_cfft_batch_q31(const arm_cfft_instance_q31 *S, q31_t *p, uint32_t nFrames,
                              uint8_t ifftFlag, uint8_t bitReverseFlag);
arm_status arm_cfft_batch_q15(const arm_cfft_instance_q15 *S, q15_t *p, uint32_t nFrames,
                              uint8_t ifftFlag, uint8_t bitReverseFlag);
/* device pointers must be 16-byte aligned (host buffers are staged) */
arm_status arm_cfft_batch_f64(const arm_cfft_instance_f64 *S, float64_t *p, uint32_t nFrames,
                              uint8_t ifftFlag, uint8_t bitReverseFlag);
arm_status arm_rfft_fast_batch_f32(const arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut,
                                   uint32_t nFrames, uint8_t ifftFlag);
/* leaves p untouched in both directions, like arm_rfft_fast_batch_f32 (the single-frame forward call keeps the
 * reference's side effect: p then holds the N/2-point CFFT) */
arm_status arm_rfft_fast_batch_f64(const arm_rfft_fast_instance_f64 *S, float64_t *p, float64_t *pOut,
                                   uint32_t nFrames, uint8_t ifftFlag);
/* arm_cfft_f32(S, p, ifftFlag, 1) fused with its usual consumers; pSrc (nFrames * 2*fftLen floats) is left untouched and
 * the spectrum itself is never written:
 *   _mag_ / _mag_squared_   + arm_cmplx_mag_f32 / arm_cmplx_mag_squared_f32: pMag receives fftLen floats per frame
 *   _peak_                  + arm_cmplx_mag_f32 + arm_max_f32: one (value, index) per frame, the first maximum wins
 * (Examples/ARM/arm_fft_bin_example/arm_fft_bin_example_f32.c:141-149).  Host or device pointers. */
arm_status arm_cfft_mag_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pSrc, float32_t *pMag,
                                  uint32_t nFrames, uint8_t ifftFlag);
arm_status arm_cfft_mag_squared_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pSrc, float32_t *pMag,
                                          uint32_t nFrames, uint8_t ifftFlag);
arm_status arm_cfft_peak_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pSrc, float32_t *pResult,
                                   uint32_t *pIndex, uint32_t nFrames, uint8_t ifftFlag);
/* Pre-FFT window multiply fused into the transform's load (the windowed frames never exist in memory):
 *   arm_rfft_fast_window_batch_f32   = arm_mult_f32(p, pWindow, tmp, fftLenRFFT) + arm_rfft_fast_f32(S, tmp, pOut, 0) per
 *       frame, the front of arm_mfcc_f32 (arm_mfcc_f32.c:112,137); pWindow: fftLenRFFT values (arm_hamming_f32 ...);
 *       p is left untouched
 *   arm_cfft_window_batch_f32        = re and im of sample n times pWindow[n] (arm_cmplx_mult_real_f32), then
 *       arm_cfft_f32(S, p, ifftFlag, 1) in place; pWindow: fftLen values
 * Host or device data pointers; pWindow is a host array (kept resident on the device, keyed by its content). */
arm_status arm_rfft_fast_window_batch_f32(const arm_rfft_fast_instance_f32 *S, const float32_t *pWindow, const float32_t *p,
                                          float32_t *pOut, uint32_t nFrames);
arm_status arm_cfft_window_batch_f32(const arm_cfft_instance_f32 *S, const float32_t *pWindow, float32_t *p, uint32_t nFrames,
                                     uint8_t ifftFlag);
/* Fixed-point real FFT over nFrames frames; direction = S->ifftFlagR.  forward: pSrc frames fftLenReal
 * scalars apart, pDst frames 2*fftLenReal apart; inverse: pSrc frames 2*fftLenReal apart (bins
 * 0..fftLenReal/2 are read), pDst frames fftLenReal apart.  pSrc is left untouched. */
arm_status arm_rfft_batch_q31(const arm_rfft_instance_q31 *S, const q31_t *pSrc, q31_t *pDst, uint32_t nFrames);
arm_status arm_rfft_batch_q15(const arm_rfft_instance_q15 *S, const q15_t *pSrc, q15_t *pDst, uint32_t nFrames);
/* status of the most recent legacy (void) exec call on this thread */
arm_status arm_cuda_last_status(void);

/* ---------------------------------------------------------------- B200 extension: devices
 *
 * Device buffers: a call runs on the device that owns the buffers, enqueued on the legacy default stream (so it is
 * ordered after the caller's earlier work on that stream or on any blocking stream) and returns when the result is
 * there.  Host buffers: the frame range of a batched call
 * is block-partitioned over the device list -- device g of G gets frames [g*ceil(B/G), min(B, (g+1)*ceil(B/G))) --
 * one host thread, stream set and table cache per device, nothing exchanged between devices (frames are independent).
 * The list is, in this order: arm_cuda_set_devices(); the environment variable CMSISDSP_CUDA_DEVICES ("all" or a comma
 * list of ordinals); every visible device, the calling thread's current one first.  Small calls (under 8 MiB per
 * device) use fewer devices; a single-frame legacy call always runs on the first one. */
arm_status arm_cuda_set_devices(const int32_t *devices, uint32_t nDevices);      /* nDevices = 0: back to the default */
uint32_t   arm_cuda_get_devices(int32_t *devices, uint32_t maxDevices);          /* returns the list's length */
/* host buffers travel in chunks of chunkMiB MiB over nStreams streams per device (0 = leave unchanged; defaults 32 and
 * 3, or CMSISDSP_CUDA_CHUNK_MIB / CMSISDSP_CUDA_NSTREAMS) */
arm_status arm_cuda_set_staging(uint32_t chunkMiB, uint32_t nStreams);
/* a call's first chunk has firstChunkMiB MiB, the next ones double up to chunkMiB, the last ones halve down again
 * (default 4, or CMSISDSP_CUDA_RAMP_MIB; 0 = every chunk full-sized): the first copy in and the last copy out are the
 * only transfers nothing overlaps */
arm_status arm_cuda_set_staging_ramp(uint32_t firstChunkMiB);
/* frees the calling thread's streams and staging buffers (also done when the thread exits) */
void       arm_cuda_release(void);
/* frees the cached device copies of MFCC coefficient sets (no MFCC call may be in flight) */
void       arm_mfcc_release_plans(void);

#ifdef __cplusplus
}
#endif
#endif /* TRANSFORM_FUNCTIONS_H_ */
