/*
 * oracle/orc_fft.h -- CPU ORACLE for the CMSIS-DSP FFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * generic-C (non-DSP, non-Neon, non-MVE) algorithm for arm_cfft_{f32,q31,q15},
 * arm_rfft_fast_f32, arm_rfft_{q31,q15} and arm_mfcc_f32.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (libcmsisdsp_cuda / libcmsisdsp_b200) never links or calls anything here.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py compares every function
 * below bit-for-bit (f32 included, -ffp-contract=off) with the reference's own
 * sources compiled from /root/reference into oracle/_ref/libcmsisdsp_ref.so,
 * and tests/test_oracle_golden.py checks it against the reference's
 * Testing/Patterns golden vectors committed under tests/golden/.
 *
 * Reference citations are relative to /root/reference/.
 */
#ifndef ORC_FFT_H
#define ORC_FFT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- tables (restated generation rules; see orc_tables.c) ---- */
/* N in {16..4096}; returns NULL for unsupported N */
const float    *orc_twiddle_f32(uint32_t N);               /* N (cos,+sin) pairs      */
const int32_t  *orc_twiddle_q31(uint32_t N);               /* 3N/4 (cos,+sin) pairs   */
const int16_t  *orc_twiddle_q15(uint32_t N);               /* 3N/4 (cos,+sin) pairs   */
const float    *orc_twiddle_rfft_f32(uint32_t N);          /* N/2 (sin,cos) pairs, N = real length 32..4096 */
const uint16_t *orc_bitrev_f32(uint32_t N, uint16_t *len); /* ordered swap list (byte offsets / complex idx*8) */
const uint16_t *orc_bitrev_fixed(uint32_t N, uint16_t *len);

/* ---- transforms, in place, one frame ---- */
/* Source/TransformFunctions/arm_cfft_f32.c:1243-1298 */
void orc_cfft_f32(uint32_t N, float *p, int ifftFlag, int bitReverseFlag);
/* Source/TransformFunctions/arm_cfft_q31.c:704-755 */
void orc_cfft_q31(uint32_t N, int32_t *p, int ifftFlag, int bitReverseFlag);
/* Source/TransformFunctions/arm_cfft_q15.c:671-722 */
void orc_cfft_q15(uint32_t N, int16_t *p, int ifftFlag, int bitReverseFlag);
/* Source/TransformFunctions/arm_rfft_fast_f32.c:675-699 ; N = real length.
 * forward destroys p (it holds the N/2-point CFFT afterwards), like the reference. */
void orc_rfft_fast_f32(uint32_t N, float *p, float *pOut, int ifftFlag);

/* ---- batch drivers (frames contiguous), nthreads pthreads, static split ---- */
void orc_cfft_f32_batch(uint32_t N, float *p, uint64_t nFrames, int ifft, int bitrev, int nthreads);
void orc_cfft_q31_batch(uint32_t N, int32_t *p, uint64_t nFrames, int ifft, int bitrev, int nthreads);
void orc_cfft_q15_batch(uint32_t N, int16_t *p, uint64_t nFrames, int ifft, int bitrev, int nthreads);
void orc_rfft_fast_f32_batch(uint32_t N, float *p, float *pOut, uint64_t nFrames, int ifft, int nthreads);

/* ---- spectrum epilogues: arm_cfft_f32 + arm_cmplx_mag[_squared]_f32 (+ arm_max_f32), orc_cfft_f32.c ---- */
void orc_cfft_mag_f32(uint32_t N, float *p, float *mag, int ifftFlag, int squared);
void orc_max_f32(const float *src, uint32_t n, float *val, uint32_t *idx);
void orc_cfft_mag_f32_batch(uint32_t N, const float *src, float *mag, float *val, uint32_t *idx, uint64_t nFrames,
                            int ifftFlag, int squared);

/* ---- fixed-point real FFT (orc_rfft_fix.c): arm_rfft_q31.c:145-181, arm_rfft_q15.c:148-182 ----
 * N = real length 32..8192.  forward: pSrc N scalars (destroyed), pDst 2N scalars; inverse: pSrc bins
 * 0..N/2 read (frames 2N scalars apart in the batch drivers), pDst N scalars. */
const int32_t *orc_real_coef_q31(int b);                   /* realCoefAQ31 (b = 0) / realCoefBQ31 (b = 1), 8192 entries */
const int16_t *orc_real_coef_q15(int b);
void orc_rfft_q31(uint32_t N, int32_t *pSrc, int32_t *pDst, int ifftFlagR, int bitReverseFlagR);
void orc_rfft_q15(uint32_t N, int16_t *pSrc, int16_t *pDst, int ifftFlagR, int bitReverseFlagR);
void orc_rfft_q31_batch(uint32_t N, const int32_t *src, int32_t *dst, uint64_t nFrames, int ifft, int bitrev, int nthreads);
void orc_rfft_q15_batch(uint32_t N, const int16_t *src, int16_t *dst, uint64_t nFrames, int ifft, int bitrev, int nthreads);

/* ---- arm_mfcc_f32 (Source/TransformFunctions/arm_mfcc_f32.c:88-174), RFFT based ---- */
void orc_mfcc_f32(uint32_t fftLen, uint32_t nbMel, uint32_t nbDct, const float *dct, const uint32_t *pos,
                  const uint32_t *len, const float *coefs, const float *window, float *pSrc, float *pDst, float *pTmp);
void orc_mfcc_f32_batch(uint32_t fftLen, uint32_t nbMel, uint32_t nbDct, const float *dct, const uint32_t *pos,
                        const uint32_t *len, const float *coefs, const float *window, const float *src,
                        uint64_t stride, float *dst, uint64_t nFrames, int nthreads);

/* table digests used by the golden checks (FNV-1a 64 over the raw bytes) */
uint64_t orc_fnv1a64(const void *data, uint64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
