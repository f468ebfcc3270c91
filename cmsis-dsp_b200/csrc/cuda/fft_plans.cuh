/*
 * fft_plans.cuh -- the compile-time plan table: for every supported (type, length) the
 * thread geometry, radix passes and exchange-buffer padding.  Shared by the CUDA kernels
 * and by the CPU emulator used in the non-GPU tests.
 *
 * f32 lengths (complex): 16..4096     reference dispatch arm_cfft_f32.c:1263-1280
 * q31/q15 lengths:       16..4096     reference dispatch arm_cfft_q31.c:712-750, arm_cfft_q15.c:679-717
 *   N = 4^m   -> stages FIRST4, MID4.., LAST4            (arm_cfft_radix4_q31.c:153-473)
 *   N = 2*4^m -> PRE2 then the same on both halves, <<1  (arm_cfft_q31.c:763-822)
 * rfft_fast lengths (real): 32..4096  -> complex length N/2 with a Mirror8 pass next to
 *   the real side (fft_body.cuh).
 *
 * Plan<Arith, N, T, F, PADA, PADB, passes...>: T threads per frame (16 points per thread
 * wherever N >= 16), F frames per CTA, PADB padding elements after every 2^PADA elements
 * of the exchange buffer.  Padding choices are validated by tests/test_emulator.py
 * (bank-conflict counts of every exchange, measured by the emulator's smem trace).
 */
#pragma once
#include "fft_body.cuh"

namespace b200fft {

typedef PassF32<2> F2;
typedef PassF32<4> F4;
typedef PassF32<8> F8;
typedef PassF32<16> F16;
typedef PassF32<32> F32;
typedef PassF32<64> F64;
typedef PassF32Mirror<8> M8;
typedef PassF32Mirror<16> M16;
typedef PassF32Mirror<32> M32;

/* ---- f32 complex ----
 * N >= 512: at most TWO passes (one shared-memory exchange): 32 or 64 points per thread, so a
 * frame belongs to one warp (two for N = 4096) and the exchange needs only __syncwarp(). */
template <int N> struct PlanCfftF32;
template <> struct PlanCfftF32<16>   { typedef Plan<ArithF32, 16, 1, 128, 0, 0, F16> type; };
template <> struct PlanCfftF32<32>   { typedef Plan<ArithF32, 32, 1, 128, 0, 0, F32> type; };
template <> struct PlanCfftF32<64>   { typedef Plan<ArithF32, 64, 1, 128, 0, 0, F64> type; };
template <> struct PlanCfftF32<128>  { typedef Plan<ArithF32, 128, 8, 16, 4, 1, F16, F8> type; };
template <> struct PlanCfftF32<256>  { typedef Plan<ArithF32, 256, 16, 8, 4, 1, F16, F16> type; };
template <> struct PlanCfftF32<512>  { typedef Plan<ArithF32, 512, 16, 8, 5, 1, F32, F16> type; };
template <> struct PlanCfftF32<1024> { typedef Plan<ArithF32, 1024, 32, 4, 5, 1, F32, F32> type; };
template <> struct PlanCfftF32<2048> { typedef Plan<ArithF32, 2048, 32, 2, 6, 1, F64, F32> type; };
template <> struct PlanCfftF32<4096> { typedef Plan<ArithF32, 4096, 64, 1, 6, 1, F64, F64> type; };

/* ---- rfft_fast_f32: NC = complex length = real length / 2 ---- */
template <int NC> struct PlanRfftFwd;   /* trailing Mirror pass */
template <> struct PlanRfftFwd<16>   { typedef Plan<ArithF32, 16, 1, 128, 0, 0, F2, M8> type; };
template <> struct PlanRfftFwd<32>   { typedef Plan<ArithF32, 32, 2, 64, 3, 1, F4, M8> type; };
template <> struct PlanRfftFwd<64>   { typedef Plan<ArithF32, 64, 4, 32, 3, 1, F8, M8> type; };
template <> struct PlanRfftFwd<128>  { typedef Plan<ArithF32, 128, 8, 16, 4, 1, F16, M8> type; };
template <> struct PlanRfftFwd<256>  { typedef Plan<ArithF32, 256, 8, 16, 4, 1, F16, M16> type; };
template <> struct PlanRfftFwd<512>  { typedef Plan<ArithF32, 512, 16, 8, 5, 1, F32, M16> type; };
template <> struct PlanRfftFwd<1024> { typedef Plan<ArithF32, 1024, 16, 4, 5, 1, F32, M32> type; };
template <> struct PlanRfftFwd<2048> { typedef Plan<ArithF32, 2048, 32, 2, 6, 1, F64, M32> type; };

template <int NC> struct PlanRfftInv;   /* leading Mirror pass */
template <> struct PlanRfftInv<16>   { typedef Plan<ArithF32, 16, 1, 128, 0, 0, M8, F2> type; };
template <> struct PlanRfftInv<32>   { typedef Plan<ArithF32, 32, 2, 64, 3, 1, M8, F4> type; };
template <> struct PlanRfftInv<64>   { typedef Plan<ArithF32, 64, 4, 32, 3, 1, M8, F8> type; };
template <> struct PlanRfftInv<128>  { typedef Plan<ArithF32, 128, 8, 16, 3, 1, M8, F16> type; };   /* padding per 8: see the fixed-point N = 128 plan */
template <> struct PlanRfftInv<256>  { typedef Plan<ArithF32, 256, 8, 16, 4, 1, M16, F16> type; };
template <> struct PlanRfftInv<512>  { typedef Plan<ArithF32, 512, 16, 8, 4, 1, M16, F32> type; };
template <> struct PlanRfftInv<1024> { typedef Plan<ArithF32, 1024, 16, 4, 5, 1, M32, F32> type; };
template <> struct PlanRfftInv<2048> { typedef Plan<ArithF32, 2048, 32, 2, 5, 1, M32, F64> type; };

/* ---- q31 / q15 ---- */
template <class AR, int N> struct PlanCfftFix;
#define FIXPLAN(N_, T_, F_, PA_, PB_, ...)                                                   \
    template <class AR> struct PlanCfftFix<AR, N_> {                                         \
        /* exchange elements are 8 bytes for both types (q15 travels sign-extended) */       \
        typedef Plan<AR, N_, T_, F_, PA_, PB_, __VA_ARGS__> type;                            \
    };
#define PF(...) PassFix<AR, __VA_ARGS__>
FIXPLAN(16,   1,   128, 0, 0, PF(ST_FIRST4, ST_LAST4))
FIXPLAN(32,   1,   128, 0, 0, PF(ST_PRE2, ST_FIRST4, ST_LAST4))
FIXPLAN(64,   1,   128, 0, 0, PF(ST_FIRST4, ST_MID4, ST_LAST4))
/* N = 128: eight threads per frame, so a half-warp holds TWO frames.  A thread stores its radix-8 outputs at 8 i + e; with one padding
 * element per 16 (PADA = 4) the stores of a frame cover the 8-byte bank slots {e .. e+3, e+8 .. e+11} and the neighbouring frame, placed
 * 8 slots further for the sake of the loads, covers the same set: two wavefronts per store instead of one, seen as 17 % excess
 * wavefronts in the (L1-bound) real FFTs of N = 256.  One padding element per 8 interleaves the two frames: conflict-free both ways
 * (tests/test_emulator.py::test_exchange_bank_conflicts from N = 128; profiles/r2_bd_fix128_padding.txt). */
FIXPLAN(128,  8,   16,  3, 1, PF(ST_PRE2, ST_FIRST4), PF(ST_MID4, ST_LAST4))
FIXPLAN(256,  16,  8,   4, 1, PF(ST_FIRST4, ST_MID4), PF(ST_MID4, ST_LAST4))
/* N >= 512: three passes of 16 points per thread.  Two passes of 32/64 points per thread (the f32
 * recipe) were measured 15-45 % SLOWER here: a fixed-point pass cannot factor its twiddles (the
 * reference multiplies by specific table entries, truncating), so a radix-64 pass loads 63 of them
 * per thread and the low-occupancy kernel waits on those loads.  Three passes with 32 points per thread (half the
 * threads and barriers: T = 32 / 64 / 128 for N = 1024 / 2048 / 4096) were also measured slower, by 8-25 %
 * (profiles/r1_e_notes.md): occupancy, not barrier count, is what these kernels live on */
FIXPLAN(512,  32,  4,   4, 1, PF(ST_PRE2, ST_FIRST4), PF(ST_MID4, ST_MID4), PF(ST_LAST4))
#ifndef FFT_FIX1024_F
#define FFT_FIX1024_F 2
#endif
#ifndef FFT_FIX_LONG_T          /* A/B builds: threads per frame of the three long plans as a divisor (1: 16 points per thread, 2: 32) */
#define FFT_FIX_LONG_T 1
#endif
FIXPLAN(1024, 64 / FFT_FIX_LONG_T,  FFT_FIX1024_F * FFT_FIX_LONG_T,   4, 1, PF(ST_FIRST4, ST_MID4), PF(ST_MID4, ST_MID4), PF(ST_LAST4))
FIXPLAN(2048, 128 / FFT_FIX_LONG_T, FFT_FIX_LONG_T,   4, 1, PF(ST_PRE2, ST_FIRST4), PF(ST_MID4, ST_MID4), PF(ST_MID4, ST_LAST4))
FIXPLAN(4096, 256 / FFT_FIX_LONG_T, 1,   4, 1, PF(ST_FIRST4, ST_MID4), PF(ST_MID4, ST_MID4), PF(ST_MID4, ST_LAST4))
#undef PF
#undef FIXPLAN
/* q31, N = 64: four threads per frame and two passes (direct kernel, 82 registers / 5 CTAs) instead of one thread per frame (tiny
 * kernel, 183 registers / 2 CTAs): cfft_q31 74.0 -> 86.5 % of the HBM peak, rfft_q31 real N = 128 forward 69.3 -> 66.8 %, inverse
 * 55.2 -> 56.9 %.  The same plan for q15 is slower everywhere (cfft_q15 51.3 -> 46.6 %, rfft_q15 58 / 40 -> 37 / 30 %) and is not
 * used (profiles/r2_aj_fix64.txt).  FFT_FIX64_ONE_THREAD restores the one-thread plan for A/B builds. */
#if !defined(FFT_FIX64_ONE_THREAD)
template <> struct PlanCfftFix<ArithQ31, 64> {
    typedef Plan<ArithQ31, 64, 4, 32, 4, 1, PassFix<ArithQ31, ST_FIRST4, ST_MID4>, PassFix<ArithQ31, ST_LAST4>> type;
};
#endif

/* ---- f64 complex (arm_cfft_f64.c: the fixed-point stage structure, 16-byte points) ----
 * 16 points (64 registers) per thread from N = 64 up; the short lengths use 4 threads per frame so that a
 * frame's loads stay 64-byte pieces (one thread per frame would read 16-byte pieces 16 N bytes apart).
 * Exchange padding: one element after every 2^PADA, chosen so that every exchange is conflict-free for 16-byte
 * elements (tests/test_emulator.py::test_f64_exchange_bank_conflicts): 8 where the first pass is radix 8. */
template <int N> struct PlanCfftF64;
#define PD(...) PassFix<ArithF64, __VA_ARGS__>
template <> struct PlanCfftF64<16>   { typedef Plan<ArithF64, 16,   4,   32, 2, 1, PD(ST_FIRST4), PD(ST_LAST4)> type; };
template <> struct PlanCfftF64<32>   { typedef Plan<ArithF64, 32,   4,   32, 3, 1, PD(ST_PRE2, ST_FIRST4), PD(ST_LAST4)> type; };
template <> struct PlanCfftF64<64>   { typedef Plan<ArithF64, 64,   4,   32, 4, 1, PD(ST_FIRST4, ST_MID4), PD(ST_LAST4)> type; };
template <> struct PlanCfftF64<128>  { typedef Plan<ArithF64, 128,  8,   16, 3, 1, PD(ST_PRE2, ST_FIRST4), PD(ST_MID4, ST_LAST4)> type; };
template <> struct PlanCfftF64<256>  { typedef Plan<ArithF64, 256,  16,  8,  4, 1, PD(ST_FIRST4, ST_MID4), PD(ST_MID4, ST_LAST4)> type; };
template <> struct PlanCfftF64<512>  { typedef Plan<ArithF64, 512,  32,  4,  3, 1, PD(ST_PRE2, ST_FIRST4), PD(ST_MID4, ST_MID4), PD(ST_LAST4)> type; };
template <> struct PlanCfftF64<1024> { typedef Plan<ArithF64, 1024, 64,  2,  4, 1, PD(ST_FIRST4, ST_MID4), PD(ST_MID4, ST_MID4), PD(ST_LAST4)> type; };
#define PDD(...) PassFix<ArithF64D, __VA_ARGS__>       /* twiddles from the reference-layout table (ArithF64D) */
template <> struct PlanCfftF64<2048> { typedef Plan<ArithF64D, 2048, 128, 1, 3, 1, PDD(ST_PRE2, ST_FIRST4), PDD(ST_MID4, ST_MID4), PDD(ST_MID4, ST_LAST4)> type; };
template <> struct PlanCfftF64<4096> { typedef Plan<ArithF64D, 4096, 256, 1, 4, 1, PDD(ST_FIRST4, ST_MID4), PDD(ST_MID4, ST_MID4), PDD(ST_MID4, ST_LAST4)> type; };
#undef PDD
#undef PD

}  // namespace b200fft
