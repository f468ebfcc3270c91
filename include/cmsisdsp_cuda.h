/*
 * cmsisdsp_cuda.h -- C ABI of libcmsisdsp_cuda.so, the thin CUDA shim under the CMSIS-DSP
 * front library (libcmsisdsp_b200.so).  Plain pointers and sizes only; no CUDA, C++ or
 * torch types appear in any signature, so it can be bound from C, ctypes, cgo, JNI ...
 *
 * Device data must be aligned to one complex element (8 bytes f32 / q31, 4 bytes q15, 16 bytes f64); other
 * pointers are refused with CMSISDSP_CUDA_ERR_ARGUMENT.
 * Every entry point works on the CURRENT device of the calling thread
 * (cmsisdsp_cuda_set_device) and enqueues on the given stream (NULL = default stream,
 * otherwise a cudaStream_t / CUstream handle cast to void*).  Transform entry points take
 * DEVICE pointers; frames are contiguous.  There is no CPU fallback: without a CUDA device
 * every call fails with CMSISDSP_CUDA_ERR_NO_DEVICE.
 *
 * What each entry point replaces in the reference (paths relative to the CMSIS-DSP tree):
 *   cmsisdsp_cuda_cfft_f32        Source/TransformFunctions/arm_cfft_f32.c:1243-1298 (+ arm_cfft_radix8_f32.c:51-291,
 *                                 arm_bitreversal2.c:84-108), looped over nFrames frames
 *   cmsisdsp_cuda_cfft_q31        Source/TransformFunctions/arm_cfft_q31.c:704-755 (+ arm_cfft_radix4_q31.c:153-834)
 *   cmsisdsp_cuda_cfft_q15        Source/TransformFunctions/arm_cfft_q15.c:671-722 (+ arm_cfft_radix4_q15.c:572-970,1434-1813)
 *   cmsisdsp_cuda_cfft_f64        Source/TransformFunctions/arm_cfft_f64.c:262-312 (+ :58-239, arm_bitreversal2.c:45-70)
 *   cmsisdsp_cuda_rfft_fast_f32   Source/TransformFunctions/arm_rfft_fast_f32.c:675-699 (+ :316-462)
 *   cmsisdsp_cuda_plan_upload     the residency of Source/CommonTables/arm_common_tables.c twiddle / bit-reversal
 *                                 tables (:8523-26700) reached through arm_cfft_instance_* (transform_functions.h:282-424)
 *   cmsisdsp_cuda_rfft_plan_upload  twiddleCoef_rfft_N (arm_common_tables.c:30820-34940) via arm_rfft_fast_instance_f32
 */
#ifndef CMSISDSP_CUDA_H
#define CMSISDSP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    CMSISDSP_CUDA_F32 = 0,
    CMSISDSP_CUDA_Q31 = 1,
    CMSISDSP_CUDA_Q15 = 2,
    CMSISDSP_CUDA_F64 = 3
};

enum {
    CMSISDSP_CUDA_OK = 0,
    CMSISDSP_CUDA_ERR_ARGUMENT = -1,      /* unsupported length / type / null pointer */
    CMSISDSP_CUDA_ERR_NO_PLAN = -2,       /* tables for (type, length) not uploaded on this device */
    CMSISDSP_CUDA_ERR_NO_DEVICE = -3,     /* no usable CUDA device */
    CMSISDSP_CUDA_ERR_RUNTIME = -4        /* a CUDA runtime call failed; see cmsisdsp_cuda_last_error() */
};

/* ---- device / memory / stream plumbing for pure-C callers ---- */
int  cmsisdsp_cuda_device_count(void);
int  cmsisdsp_cuda_set_device(int device);
int  cmsisdsp_cuda_get_device(void);
int  cmsisdsp_cuda_malloc(void **devPtr, size_t bytes);
int  cmsisdsp_cuda_free(void *devPtr);
int  cmsisdsp_cuda_host_alloc(void **hostPtr, size_t bytes);          /* pinned host memory */
int  cmsisdsp_cuda_host_free(void *hostPtr);
int  cmsisdsp_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int  cmsisdsp_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int  cmsisdsp_cuda_stream_create(void **stream);
int  cmsisdsp_cuda_stream_destroy(void *stream);
int  cmsisdsp_cuda_stream_synchronize(void *stream);
/* 1 if ptr is device (or managed) memory usable by kernels on the current device, 0 if host, <0 on error */
int  cmsisdsp_cuda_is_device_pointer(const void *ptr);
/* ordinal of the device that owns ptr, -1 for host memory, < -1 on error */
int  cmsisdsp_cuda_pointer_device(const void *ptr);
/* CUDA-event timing on `stream`: begin/end return an opaque timer; elapsed in milliseconds */
int  cmsisdsp_cuda_timer_begin(void **timer, void *stream);
int  cmsisdsp_cuda_timer_end(void *timer, void *stream, float *elapsedMs);

/* ---- plans: device-resident tables, keyed by (device, type, fftLen) and by the CONTENT of the host tables ----
 * Every *_plan_upload call is cheap when the tables are known (pointer compare + 32-sample fingerprint; a 64-bit hash
 * of the whole table when the pointer is new) and selects, for the calling thread, the device copy the following
 * transform calls of that (type, length) use.  Up to 4 distinct tables per (device, type, length). */
/* pTwiddle: f32 -> 2*fftLen floats (cos,+sin); f64 -> 2*fftLen doubles; q31/q15 -> 3*fftLen/2 values (3N/4 pairs).
 * pBitRevTable/bitRevLength: the ordered swap list of the instance struct; it is expanded to the
 * output permutation used when bitReverseFlag == 0.  Idempotent. */
int  cmsisdsp_cuda_plan_upload(int type, uint32_t fftLen, const void *pTwiddle,
                               const uint16_t *pBitRevTable, uint16_t bitRevLength);
/* pTwiddleRFFT: fftLenReal floats = fftLenReal/2 (sin,cos) pairs; needs the f32 plan of fftLenReal/2 too */
int  cmsisdsp_cuda_rfft_plan_upload(uint32_t fftLenReal, const float *pTwiddleRFFT);
int  cmsisdsp_cuda_plan_ready(int type, uint32_t fftLen);              /* 1 / 0 */
int  cmsisdsp_cuda_rfft_plan_ready(uint32_t fftLenReal);

/* ---- transforms on device-resident batches (in place for cfft) ---- */
int  cmsisdsp_cuda_cfft_f32(void *d_p, uint32_t fftLen, uint64_t nFrames,
                            uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream);
int  cmsisdsp_cuda_cfft_q31(void *d_p, uint32_t fftLen, uint64_t nFrames,
                            uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream);
int  cmsisdsp_cuda_cfft_q15(void *d_p, uint32_t fftLen, uint64_t nFrames,
                            uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream);
/* arm_cfft_f32 whose result is left in plain binary bit-reversed order: what the deprecated arm_cfft_radix4_f32 /
 * arm_cfft_radix2_f32 produce with bitReverseFlag = 0 (arm_cfft_radix4_f32.c:81-118, arm_cfft_radix2_f32.c) */
int  cmsisdsp_cuda_cfft_f32_bitrev_order(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream);
/* d_p: nFrames*2*fftLen doubles, 16-byte aligned */
int  cmsisdsp_cuda_cfft_f64(void *d_p, uint32_t fftLen, uint64_t nFrames,
                            uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream);
/* arm_rfft_fast_f64 (Source/TransformFunctions/arm_rfft_fast_f64.c:207-233): one fused kernel per direction (the
 * N/2-point f64 transform with the split stage as its epilogue / the merge stage as the load of the inverse).
 * Needs the f64 plan of fftLenReal/2 and pTwiddleRFFT = fftLenReal doubles ((sin,cos) pairs).  d_p is left untouched;
 * d_p and d_out: nFrames*fftLenReal doubles, 16-byte aligned, not aliased. */
int  cmsisdsp_cuda_rfft_f64_plan_upload(uint32_t fftLenReal, const double *pTwiddleRFFT);
int  cmsisdsp_cuda_rfft_f64_plan_ready(uint32_t fftLenReal);
int  cmsisdsp_cuda_rfft_fast_f64(const void *d_p, void *d_out, uint32_t fftLenReal, uint64_t nFrames,
                                 uint8_t ifftFlag, void *stream);
/* d_p: nFrames*fftLenReal floats, left untouched (the reference clobbers it);
 * d_out: nFrames*fftLenReal floats, packed {DC, Nyquist, Re1, Im1, ...}.  d_p and d_out must not alias. */
int  cmsisdsp_cuda_rfft_fast_f32(const void *d_p, void *d_out, uint32_t fftLenReal, uint64_t nFrames,
                                 uint8_t ifftFlag, void *stream);

/* ---- arm_cfft_f32 with a fused spectrum epilogue (the spectrum itself is never written) ----
 * = arm_cfft_f32(.., ifftFlag, 1) + arm_cmplx_mag_f32 / arm_cmplx_mag_squared_f32 (ComplexMathFunctions/
 * arm_cmplx_mag_f32.c:252-264) [+ arm_max_f32, the first maximum wins]: Examples/ARM/arm_fft_bin_example/
 * arm_fft_bin_example_f32.c:141-149.  d_src: nFrames*2*fftLen floats, left untouched; d_mag: nFrames*fftLen floats;
 * d_val / d_idx: nFrames floats / uint32.  Needs the cfft plan of (f32, fftLen). */
int  cmsisdsp_cuda_cfft_mag_f32(const void *d_src, void *d_mag, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag,
                                uint8_t squared, void *stream);
int  cmsisdsp_cuda_cfft_peak_f32(const void *d_src, void *d_val, void *d_idx, uint32_t fftLen, uint64_t nFrames,
                                 uint8_t ifftFlag, void *stream);

/* ---- arm_rfft_q31 / arm_rfft_q15 (Source/TransformFunctions/arm_rfft_q31.c:145-181, arm_rfft_q15.c:148-182) ----
 * fftLenReal in {32..8192}.  bitReverseFlagR is handed to the complex transform inside, as the reference does
 * (arm_rfft_q31.c:164,173): 1 = natural order; 0 = that transform's result stays in bit-reversed order (forward: the
 * split stage then reads it as if it were natural order).  The plan is the cfft plan of the
 * same type and length fftLenReal/2 (cmsisdsp_cuda_plan_upload) plus the split-stage coefficients: the
 * instance's pTwiddleAReal / pTwiddleBReal (realCoefA/B, read at twidCoefRModifier), compacted on upload.
 * forward (ifftFlagR = 0): d_src nFrames*fftLenReal scalars (left untouched), d_dst nFrames*2*fftLenReal
 *     scalars = fftLenReal complex bins per frame, conjugate mirror included, as the reference writes them;
 * inverse (ifftFlagR = 1): d_src frames 2*fftLenReal scalars apart, bins 0..fftLenReal/2 are read;
 *     d_dst nFrames*fftLenReal scalars.  d_src and d_dst must not alias. */
int  cmsisdsp_cuda_rfft_fix_plan_upload(int type, uint32_t fftLenReal, const void *pTwiddleAReal,
                                        const void *pTwiddleBReal, uint32_t twidCoefRModifier);
int  cmsisdsp_cuda_rfft_fix_plan_ready(int type, uint32_t fftLenReal);
int  cmsisdsp_cuda_rfft_q31(const void *d_src, void *d_dst, uint32_t fftLenReal, uint64_t nFrames,
                            uint8_t ifftFlagR, uint8_t bitReverseFlagR, void *stream);
int  cmsisdsp_cuda_rfft_q15(const void *d_src, void *d_dst, uint32_t fftLenReal, uint64_t nFrames,
                            uint8_t ifftFlagR, uint8_t bitReverseFlagR, void *stream);

/* ---- window multiply fused into the load of arm_cfft_f32 / arm_rfft_fast_f32 (forward) ----
 * = arm_mult_f32(p, window, p, fftLen) + the transform, the pre-FFT step of Source/TransformFunctions/arm_mfcc_f32.c:112
 * (windows: Source/WindowFunctions/arm_hamming_f32.c:72 ...), without the round trip of the windowed frame through
 * memory.  cmsisdsp_cuda_window_upload makes `length` window values resident and selects them for the calling thread
 * (content-keyed like the plans); the complex transform multiplies re and im of sample n by window[n]; natural-order
 * output. */
int  cmsisdsp_cuda_window_upload(uint32_t length, const float *pWindow);
int  cmsisdsp_cuda_cfft_window_f32(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream);
int  cmsisdsp_cuda_rfft_fast_window_f32(const void *d_p, void *d_out, uint32_t fftLenReal, uint64_t nFrames, void *stream);

/* ---- deprecated arm_cfft_radix2_q31 / arm_cfft_radix2_q15 (Source/TransformFunctions/arm_cfft_radix2_q31.c:62-318,
 * arm_cfft_radix2_q15.c:62-78,275-386,577-681): log2(fftLen) radix-2 stages with their own per-stage scaling, then the
 * bit reversal (always: natural-order result).  pCoef / twidCoefModifier as in the instance (the 4096-point table read
 * with a stride); in place on nFrames frames of 2*fftLen scalars. ---- */
int  cmsisdsp_cuda_radix2_plan_upload(int type, uint32_t fftLen, const void *pCoef, uint32_t twidCoefModifier);
int  cmsisdsp_cuda_cfft_radix2_q31(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream);
int  cmsisdsp_cuda_cfft_radix2_q15(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream);

/* ---- arm_mfcc_f32 front end (Source/TransformFunctions/arm_mfcc_f32.c:88-174, arm_mfcc_init_f32.c:91-121) ----
 * One fused kernel per frame batch: normalise, window, rfft, magnitude, mel filter bank, log, DCT.
 * The plan keeps device copies of the caller's coefficient arrays (the arguments of arm_mfcc_init_f32);
 * it needs the rfft plan of fftLen uploaded first (cmsisdsp_cuda_plan_upload for fftLen/2 +
 * cmsisdsp_cuda_rfft_plan_upload).  fftLen in {256, 512, 1024, 2048, 4096}. */
int  cmsisdsp_cuda_mfcc_plan_create(uint32_t fftLen, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                                    const float *dctCoefs, const uint32_t *filterPos, const uint32_t *filterLengths,
                                    const float *filterCoefs, const float *windowCoefs, void **plan);
int  cmsisdsp_cuda_mfcc_plan_destroy(void *plan);
/* frame f = fftLen floats starting at d_src + f*strideFloats (strideFloats = fftLen: back to back;
 * smaller: overlapping frames; must be even); d_dst: nFrames * nbDctOutputs floats.  d_src is not modified. */
int  cmsisdsp_cuda_mfcc_f32(const void *plan, const void *d_src, uint64_t strideFloats, void *d_dst,
                            uint64_t nFrames, void *stream);

/* ---- tuning / diagnostics ---- */
/* Kernel flavour used by the transform entry points: -1 = the measured default of each (op, length),
 * 0 = direct (one CTA per frame group, loads into registers), 1 = persistent TMA-fed kernel where the
 * pair has one.  Both compute identical results; the switch exists for A/B measurements and tests
 * (the environment variable CMSISDSP_CUDA_KERNEL=direct|pipe sets the initial value). */
int  cmsisdsp_cuda_set_kernel_flavour(int flavour);
const char *cmsisdsp_cuda_last_error(void);        /* thread-local, never NULL */
uint64_t    cmsisdsp_cuda_launch_count(void);      /* kernels launched by this library so far */
/* static facts about the kernel chosen for (op, fftLen): op 0 cfft_f32, 1 cfft_q31, 2 cfft_q15,
 * 3 rfft forward, 4 rfft inverse, 5/6 rfft_q31 forward/inverse, 7/8 rfft_q15 forward/inverse (fftLen = real
 * length for 3..8), 9 cfft_f32 + magnitude epilogue, 10 cfft_f64, 11/12 rfft_fast_f64 forward/inverse (real length).  Any out pointer may be NULL. */
int  cmsisdsp_cuda_kernel_info(int op, uint32_t fftLen, int *threadsPerCta, int *framesPerCta,
                               int *smemBytes, int *regsPerThread, int *ctasPerSm);

#ifdef __cplusplus
}
#endif
#endif
