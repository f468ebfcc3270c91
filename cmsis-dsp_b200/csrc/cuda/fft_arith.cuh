/*
 * fft_arith.cuh -- per-thread butterfly arithmetic for the batched CMSIS-DSP FFT kernels.
 *
 * Everything here is register-level math on small arrays with compile-time indices,
 * written so that the same code compiles for the device (nvcc, sm_100a) and for the
 * host-side kernel emulator used by the CPU tests (tests/emu).
 *
 *  - f32: plain radix-2/4/8/16 DFT butterflies (the f32 path only has to match the
 *    reference to 2e-6 relative RMS, so the factorisation is free:
 *    reference = Source/TransformFunctions/arm_cfft_radix8_f32.c:51-291).
 *  - q31/q15: the reference's radix-4 DIF stage arithmetic restated operation for
 *    operation (shifts, truncation, __SSAT, wrap-around), because those outputs must be
 *    bit-exact: Source/TransformFunctions/arm_cfft_radix4_q31.c:153-473,524-834,
 *    arm_cfft_radix4_q15.c:572-970,1434-1813, arm_cfft_q31.c:763-881, arm_cfft_q15.c:782-927,
 *    Include/dsp/none.h:78-94,185-194.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FFT_HD __host__ __device__ __forceinline__
#else
#define FFT_HD inline __attribute__((always_inline))
#endif

namespace b200fft {

struct alignas(8) cf32 { float x, y; };
struct alignas(8) ci32 { int32_t x, y; };
struct alignas(4) ci16 { int16_t x, y; };
struct alignas(16) cf64 { double x, y; };
/* split-stage coefficients of one bin of the fixed-point real FFT: realCoefA[2k], realCoefA[2k+1],
 * realCoefB[2k], realCoefB[2k+1] at the instance's twidCoefRModifier, 32-bit (q15 sign-extended) */
struct alignas(16) ci32x4 { int32_t a0, a1, b0, b1; };

/* ------------------------------------------------------------------ f32
 *
 * Complex values live in aligned register pairs, and on the device every complex add,
 * subtract and multiply is issued as a PACKED fp32x2 instruction (FADD2 / FMUL2 / FFMA2,
 * new in sm_100): one issue slot for the real and the imaginary lane.  ptxas folds the
 * lane swap (.LO_HI), the half negation (.NP) and scalar broadcasts (.F32) into operand
 * modifiers, so a radix-4 butterfly is 8 instructions and a complex multiply is 2. */

#if defined(__CUDA_ARCH__)
FFT_HD float2 f2(cf32 a) { return make_float2(a.x, a.y); }
FFT_HD cf32 c2(float2 a) { return {a.x, a.y}; }
FFT_HD cf32 cadd(cf32 a, cf32 b) { return c2(__fadd2_rn(f2(a), f2(b))); }
FFT_HD cf32 csub(cf32 a, cf32 b) { return c2(__fadd2_rn(f2(a), make_float2(-b.x, -b.y))); }
/* a + (-i)*b  and  a + (+i)*b */
FFT_HD cf32 cadd_mi(cf32 a, cf32 b) { return c2(__fadd2_rn(f2(a), make_float2(b.y, -b.x))); }
FFT_HD cf32 cadd_pi(cf32 a, cf32 b) { return c2(__fadd2_rn(f2(a), make_float2(-b.y, b.x))); }
/* a * conj(w), w = (cos, +sin) as stored in the reference's twiddle tables */
FFT_HD cf32 mul_conj(cf32 a, cf32 w)
{
    return c2(__ffma2_rn(f2(a), make_float2(w.x, w.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(w.y, -w.y))));
}
/* a * w (no conjugation) */
FFT_HD cf32 mul_cplx(cf32 a, cf32 w)
{
    return c2(__ffma2_rn(f2(a), make_float2(w.x, w.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y))));
}
FFT_HD cf32 cscale(cf32 a, float s) { return c2(__fmul2_rn(f2(a), make_float2(s, s))); }
/* a*s + b */
FFT_HD cf32 caxpy(cf32 a, float s, cf32 b) { return c2(__ffma2_rn(f2(a), make_float2(s, s), f2(b))); }
#else
FFT_HD cf32 cadd(cf32 a, cf32 b) { return {a.x + b.x, a.y + b.y}; }
FFT_HD cf32 csub(cf32 a, cf32 b) { return {a.x - b.x, a.y - b.y}; }
FFT_HD cf32 cadd_mi(cf32 a, cf32 b) { return {a.x + b.y, a.y - b.x}; }
FFT_HD cf32 cadd_pi(cf32 a, cf32 b) { return {a.x - b.y, a.y + b.x}; }
FFT_HD cf32 mul_conj(cf32 a, cf32 w) { return {a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y}; }
FFT_HD cf32 mul_cplx(cf32 a, cf32 w) { return {a.x * w.x - a.y * w.y, a.y * w.x + a.x * w.y}; }
FFT_HD cf32 cscale(cf32 a, float s) { return {a.x * s, a.y * s}; }
FFT_HD cf32 caxpy(cf32 a, float s, cf32 b) { return {a.x * s + b.x, a.y * s + b.y}; }
#endif
/* a * (-i) and a * (+i) */
FFT_HD cf32 mul_mi(cf32 a) { return {a.y, -a.x}; }
FFT_HD cf32 mul_pi(cf32 a) { return {-a.y, a.x}; }

/* cos(2*pi*k/128), k = 0..32; the other three quadrants follow by symmetry */
FFT_HD constexpr float cos128_q(int k)
{
    constexpr float t[33] = {
        1.000000000e+00f, 9.987954562e-01f, 9.951847267e-01f, 9.891765100e-01f,
        9.807852804e-01f, 9.700312532e-01f, 9.569403357e-01f, 9.415440652e-01f,
        9.238795325e-01f, 9.039892931e-01f, 8.819212643e-01f, 8.577286100e-01f,
        8.314696123e-01f, 8.032075315e-01f, 7.730104534e-01f, 7.409511254e-01f,
        7.071067812e-01f, 6.715589548e-01f, 6.343932842e-01f, 5.956993045e-01f,
        5.555702330e-01f, 5.141027442e-01f, 4.713967368e-01f, 4.275550934e-01f,
        3.826834324e-01f, 3.368898534e-01f, 2.902846773e-01f, 2.429801799e-01f,
        1.950903220e-01f, 1.467304745e-01f, 9.801714033e-02f, 4.906767433e-02f,
        0.000000000e+00f};
    return t[k];
}
/* cos / sin of 2*pi*k/128 for any integer k >= 0 */
FFT_HD constexpr float cos128(int k)
{
    k &= 127;
    return k <= 32 ? cos128_q(k) : (k <= 64 ? -cos128_q(64 - k) : (k <= 96 ? -cos128_q(k - 64) : cos128_q(128 - k)));
}
FFT_HD constexpr float sin128(int k) { return cos128(k + 96); }   /* sin(x) = cos(x - pi/2) = cos(x + 3*pi/2) */

/* x * w_R^K,  w_R = exp(-2*pi*i/R),  K and R compile-time (R divides 128) */
template <int R, int K> FFT_HD cf32 mul_wconst(cf32 x)
{
    constexpr int k = ((K % R) + R) % R;
    if (k == 0) return x;
    if (4 * k == R) return mul_mi(x);
    if (2 * k == R) return {-x.x, -x.y};
    if (4 * k == 3 * R) return mul_pi(x);
    constexpr float c = cos128(k * (128 / R)), s = sin128(k * (128 / R));
    return mul_conj(x, cf32{c, s});
}

/* position of output t inside the array after the in-place recursive DIF below */
FFT_HD constexpr int dft_pos(int R, int t) { return R <= 4 ? t : (R / 4) * (t % 4) + dft_pos(R / 4, t / 4); }
/* twiddle values a post-twiddled radix-R butterfly reads: 3 per radix-4 level (+ R-1 at the base) */
FFT_HD constexpr int dft_tw_slots(int R) { return R <= 4 ? R - 1 : 3 + dft_tw_slots(R / 4); }

FFT_HD void dft4(cf32 &x0, cf32 &x1, cf32 &x2, cf32 &x3)
{
    cf32 a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3), a3 = csub(x1, x3);
    x0 = cadd(a0, a2);
    x2 = csub(a0, a2);
    x1 = cadd_mi(a1, a3);
    x3 = cadd_pi(a1, a3);
}

/*
 * In-place radix-R DIF butterfly on x[OFF + STR*q], q < R, optionally followed by the
 * Stockham post-twiddle  out[t] *= W^(e*t)  (TW = true), factored level by level:
 * with q = q1 + (R/4) q2 and t = r2 + 4 t2,
 *      out[t] = DFT_{R/4} over q1 { w_R^(q1 r2) * W^(e r2) * DFT_4 over q2 { x } } [t2] * W^(4 e t2)
 * so each radix-4 level needs only W^(e'), W^(2e'), W^(3e') (e' = 4^level * e) -- the three
 * twiddles the reference's radix-4/8 stages read per butterfly (arm_cfft_radix8_f32.c:149-287)
 * -- instead of R-1 distinct table entries: 9 loads for R = 64 instead of 63.
 * tw[s] = the s-th value of that list (Pass::fill order).  Output t ends up at dft_pos(R, t).
 */
template <int R, int OFF, int STR, bool TW> struct DftRec {
    static FFT_HD void run(cf32 *x, const cf32 *tw)
    {
        constexpr int Q = R / 4;
#pragma unroll
        for (int q1 = 0; q1 < Q; q1++)
            dft4(x[OFF + STR * q1], x[OFF + STR * (q1 + Q)], x[OFF + STR * (q1 + 2 * Q)], x[OFF + STR * (q1 + 3 * Q)]);
        step2<1>(x, tw);
        step2<2>(x, tw);
        step2<3>(x, tw);
        DftRec<Q, OFF, STR, TW>::run(x, tw + 3);
        DftRec<Q, OFF + STR * Q, STR, TW>::run(x, tw + 3);
        DftRec<Q, OFF + STR * 2 * Q, STR, TW>::run(x, tw + 3);
        DftRec<Q, OFF + STR * 3 * Q, STR, TW>::run(x, tw + 3);
    }
    template <int R2> static FFT_HD void step2(cf32 *x, const cf32 *tw) { qloop<R2, 0>(x, tw); }
    template <int R2, int Q1> static FFT_HD void qloop(cf32 *x, const cf32 *tw)
    {
        constexpr int Q = R / 4;
        if constexpr (Q1 < Q) {
            cf32 v = mul_wconst<R, Q1 * R2>(x[OFF + STR * (Q1 + R2 * Q)]);
            if constexpr (TW) v = mul_conj(v, tw[R2 - 1]);
            x[OFF + STR * (Q1 + R2 * Q)] = v;
            qloop<R2, Q1 + 1>(x, tw);
        }
    }
};
template <int OFF, int STR, bool TW> struct DftRec<4, OFF, STR, TW> {
    static FFT_HD void run(cf32 *x, const cf32 *tw)
    {
        dft4(x[OFF], x[OFF + STR], x[OFF + 2 * STR], x[OFF + 3 * STR]);
        if constexpr (TW) {
            x[OFF + STR] = mul_conj(x[OFF + STR], tw[0]);
            x[OFF + 2 * STR] = mul_conj(x[OFF + 2 * STR], tw[1]);
            x[OFF + 3 * STR] = mul_conj(x[OFF + 3 * STR], tw[2]);
        }
    }
};
template <int OFF, int STR, bool TW> struct DftRec<2, OFF, STR, TW> {
    static FFT_HD void run(cf32 *x, const cf32 *tw)
    {
        cf32 a = x[OFF], b = x[OFF + STR];
        x[OFF] = cadd(a, b);
        x[OFF + STR] = csub(a, b);
        if constexpr (TW) x[OFF + STR] = mul_conj(x[OFF + STR], tw[0]);
    }
};
template <int OFF, int STR, bool TW> struct DftRec<1, OFF, STR, TW> {
    static FFT_HD void run(cf32 *, const cf32 *) {}
};

/* natural-order radix-R butterfly: x[t] = DFT_R(x)[t] (* W^(e t) when TW) */
template <int R, int T0> FFT_HD void dft_unpermute(const cf32 *x, cf32 *y)
{
    if constexpr (T0 < R) {
        y[T0] = x[dft_pos(R, T0)];
        dft_unpermute<R, T0 + 1>(x, y);
    }
}
template <int R, int T0> FFT_HD void dft_copy(const cf32 *y, cf32 *x)
{
    if constexpr (T0 < R) {
        x[T0] = y[T0];
        dft_copy<R, T0 + 1>(y, x);
    }
}
template <int R, bool TW> FFT_HD void dft_f32(cf32 *x, const cf32 *tw)
{
    DftRec<R, 0, 1, TW>::run(x, tw);
    if constexpr (R > 4) {
        cf32 y[R];
        dft_unpermute<R, 0>(x, y);
        dft_copy<R, 0>(y, x);
    }
}

/* An Arith names four representations of a complex point: elem (HBM), work (registers),
 * xelem (shared-memory exchange) and telem (device twiddle table). */
struct ArithF32 {
    static constexpr bool kBiased = false;       /* see ArithQ15::bfly4 */
    static constexpr bool kPreShift = false;     /* see ArithQ15::load_shifted */
    static constexpr bool kDirectTw = false;
    static constexpr int kTableNum = 1, kTableDen = 1;
    typedef cf32 elem;
    typedef cf32 work;
    typedef cf32 twid;
    typedef cf32 xelem;
    typedef cf32 telem;
    static FFT_HD work load(elem e) { return e; }
    static FFT_HD elem store(work w) { return w; }
    static FFT_HD work xload(xelem e) { return e; }
    static FFT_HD xelem xstore(work w) { return w; }
    static FFT_HD telem tw_expand(elem e) { return e; }
    static FFT_HD work shl1(work w) { return w; }
};

/* ------------------------------------------------------------------ q31 */

FFT_HD int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
FFT_HD int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
FFT_HD int32_t wshl1(int32_t a) { return (int32_t)((uint32_t)a << 1); }
/* truncating high product hi32(a*b), two spellings for the two pipes that can produce it on sm_100:
 *   hi32_xu   __mulhi -> IMAD.HI, executed by the XU pipe (one warp instruction per 8 cycles per scheduler)
 *   hi32_fma  64-bit product -> IMAD.WIDE on the FMA pipe, upper register taken as is
 * A q31 butterfly needs 12 of them per 4 points; all on the XU pipe that pipe bounds the kernel (55 % busy at
 * 60 % of the HBM peak, profiles/r1_e_ncu_q31.txt).  Measured (profiles/r1_e_notes.md): all on the FMA pipe is
 * the fastest for every length but 32 (+4..12 points of HBM peak at N = 64, 256, 2048, 4096); half and half
 * performs like all-XU.  N = 32 keeps IMAD.HI: the wide products cost it 15 registers and one resident CTA. */
FFT_HD int32_t hi32_xu(int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    return __mulhi(a, b);
#else
    return (int32_t)(((int64_t)a * b) >> 32);
#endif
}
FFT_HD int32_t hi32_fma(int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    int32_t hi;
    asm("{\n\t.reg .b64 t;\n\t.reg .b32 lo;\n\tmul.wide.s32 t, %1, %2;\n\tmov.b64 {lo, %0}, t;\n\t}" : "=r"(hi) : "r"(a), "r"(b));
    return hi;
#else
    return (int32_t)(((int64_t)a * b) >> 32);
#endif
}
#if defined(FFT_HI32_XU)
FFT_HD int32_t hi32(int32_t a, int32_t b) { return hi32_xu(a, b); }
#else
FFT_HD int32_t hi32(int32_t a, int32_t b) { return hi32_fma(a, b); }
#endif
/* SMMULR / SMMLAR / SMMLSR (none.h:185-194).
 *
 * The accumulating forms only ever keep the upper word, and `a << 32` does not touch the lower one, so
 *     ((a << 32) + x*y + 2^31) >> 32 = a + ((x*y + 2^31) >> 32)
 *     ((a << 32) - x*y + 2^31) >> 32 = a - ((x*y + 2^31 - 1) >> 32)       (floor(-v) = -ceil(v))
 * with wrap-around adds (the reference's 64-bit sum wraps the same way).  Written like that every rounding
 * multiply-accumulate is ONE IMAD.HI with the constant as its 64-bit addend (XU pipe) plus a share of a 3-input IADD3;
 * written as the 64-bit expression ptxas emits IMAD.WIDE + a carry-generating IADD3 + IADD3.X (27 instructions per bin of
 * the real-FFT split stage against 13), all on the FMA / ALU pipes that bound those kernels.  FFT_RMAC_WIDE restores the
 * 64-bit form for A/B runs. */
FFT_HD int32_t rhi32(int32_t x, int32_t y) { return (int32_t)(((int64_t)x * y + 0x80000000LL) >> 32); }
FFT_HD int32_t rlo32(int32_t x, int32_t y) { return (int32_t)(((int64_t)x * y + 0x7fffffffLL) >> 32); }
#if defined(FFT_RMAC_WIDE)
FFT_HD int32_t rhi32_acc(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) + (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
FFT_HD int32_t rhi32_sub(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) - (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
#else
FFT_HD int32_t rhi32_acc(int32_t a, int32_t x, int32_t y) { return (int32_t)((uint32_t)a + (uint32_t)rhi32(x, y)); }
FFT_HD int32_t rhi32_sub(int32_t a, int32_t x, int32_t y) { return (int32_t)((uint32_t)a - (uint32_t)rlo32(x, y)); }
#endif

enum StageKind { ST_PRE2 = 0, ST_FIRST4 = 1, ST_MID4 = 2, ST_LAST4 = 3 };

template <bool INV> FFT_HD ci32 rot_q31(int32_t R, int32_t S, ci32 w)
{
    if (!INV) return {wadd(hi32(R, w.x), hi32(S, w.y)), wsub(hi32(S, w.x), hi32(R, w.y))};
    return {wsub(hi32(R, w.x), hi32(S, w.y)), wadd(hi32(S, w.x), hi32(R, w.y))};
}

#if defined(FFT_FIX_DIRECT_TW)
#define FFT_FIX_DIRECT_TW_VALUE true
#else
#define FFT_FIX_DIRECT_TW_VALUE false
#endif
struct ArithQ31 {
    static constexpr bool kBiased = false;       /* see ArithQ15::bfly4 */
    static constexpr bool kPreShift = false;     /* see ArithQ15::load_shifted */
    static constexpr bool kDirectTw = FFT_FIX_DIRECT_TW_VALUE;     /* fft_frame.cuh: PassFix::kDirect */
    static constexpr int kTableNum = 3, kTableDen = 4;        /* reference table: 3N/4 entries */
    typedef ci32 elem;      /* storage element */
    typedef ci32 work;      /* register element */
    typedef ci32 twid;
    typedef ci32 xelem;
    typedef ci32 telem;
    static FFT_HD work load(elem e) { return e; }
    static FFT_HD elem store(work w) { return w; }
    static FFT_HD work xload(xelem e) { return e; }
    static FFT_HD xelem xstore(work w) { return w; }
    static FFT_HD telem tw_expand(elem e) { return e; }
    static FFT_HD twid tload(telem e) { return e; }
    static FFT_HD work shl1(work w) { return {wshl1(w.x), wshl1(w.y)}; }

    /* radix-4 DIF stage butterfly; outputs in residue order (a', b'[W^1], c'[W^2], d'[W^3]).
     * w1/w2/w3 = table entries (cos,+sin) of W^1, W^2, W^3 for this butterfly. */
    template <int KIND, bool INV, bool TAIL = false, bool PRE = false>
    static FFT_HD void bfly4(work &A, work &B, work &C, work &D, twid w1, twid w2, twid w3, int32_t = 0)
    {
        work a = A, b = B, c = C, e = D;
        if (KIND == ST_FIRST4) {
            a.x >>= 4; a.y >>= 4; b.x >>= 4; b.y >>= 4; c.x >>= 4; c.y >>= 4; e.x >>= 4; e.y >>= 4;
        }
        int32_t r1 = wadd(a.x, c.x), r2 = wsub(a.x, c.x), s1 = wadd(a.y, c.y), s2 = wsub(a.y, c.y);
        int32_t t1 = wadd(b.x, e.x), t2 = wadd(b.y, e.y), u1 = wsub(b.y, e.y), u2 = wsub(b.x, e.x);
        if (KIND == ST_LAST4) {
            A = {wadd(r1, t1), wadd(s1, t2)};
            C = {wsub(r1, t1), wsub(s1, t2)};
            work p = {wadd(r2, u1), wsub(s2, u2)}, q = {wsub(r2, u1), wadd(s2, u2)};
            B = INV ? q : p;
            D = INV ? p : q;
            return;
        }
        work oa = {wadd(r1, t1), wadd(s1, t2)};
        work oc = rot_q31<INV>(wsub(r1, t1), wsub(s1, t2), w2);
        work ob = INV ? rot_q31<INV>(wsub(r2, u1), wadd(s2, u2), w1) : rot_q31<INV>(wadd(r2, u1), wsub(s2, u2), w1);
        work od = INV ? rot_q31<INV>(wadd(r2, u1), wsub(s2, u2), w3) : rot_q31<INV>(wsub(r2, u1), wadd(s2, u2), w3);
        if (KIND == ST_FIRST4) {
            oc = shl1(oc); ob = shl1(ob); od = shl1(od);
        } else {
            oa.x >>= 2; oa.y >>= 2;
            oc.x >>= 1; oc.y >>= 1; ob.x >>= 1; ob.y >>= 1; od.x >>= 1; od.y >>= 1;
        }
        A = oa; B = ob; C = oc; D = od;
    }
    /* radix-2 pre-pass of the N = 2*4^m lengths (arm_cfft_q31.c:777-797 / :838-856) */
    template <bool INV, bool PRE = false> static FFT_HD void bfly2(work &A, work &B, twid w)
    {
        work a = A, b = B;
        int32_t xt = wsub(a.x >> 2, b.x >> 2), yt = wsub(a.y >> 2, b.y >> 2);
        A = {wadd(a.x >> 2, b.x >> 2), wadd(b.y >> 2, a.y >> 2)};
        int32_t p0 = rhi32(xt, w.x), p1 = rhi32(yt, w.x);
        if (!INV) { p0 = rhi32_acc(p0, yt, w.y); p1 = rhi32_sub(p1, xt, w.y); }
        else      { p0 = rhi32_sub(p0, yt, w.y); p1 = rhi32_acc(p1, xt, w.y); }
        B = {wshl1(p0), wshl1(p1)};
    }

    /* ---- real FFT (arm_rfft_q31.c): S1 = X[k], S2 = X[L2-k]; the reference's exact sequence of rounding
     * multiply-accumulates, which reads only A[2k], A[2k+1] and B[2k] (B[2k+1] == -A[2k+1]) ---- */
    static FFT_HD work split_fwd(work s1, work s2, ci32x4 c)          /* arm_split_rfft_q31, :290-325 */
    {
        int32_t r = rhi32(s1.x, c.a0), i = rhi32(s1.x, c.a1);
        r = rhi32_sub(r, s1.y, c.a1);
        i = rhi32_acc(i, s1.y, c.a0);
        r = rhi32_sub(r, s2.y, c.a1);
        i = rhi32_sub(i, s2.y, c.b0);
        r = rhi32_acc(r, s2.x, c.b0);
        i = rhi32_sub(i, s2.x, c.a1);
        return {r, i};
    }
    static FFT_HD work split_inv(work s1, work s2, ci32x4 c)          /* arm_split_rifft_q31, :438-466 */
    {
        int32_t r = rhi32(s1.x, c.a0), i = rhi32(s1.x, wsub(0, c.a1));
        r = rhi32_acc(r, s1.y, c.a1);
        i = rhi32_acc(i, s1.y, c.a0);
        r = rhi32_acc(r, s2.y, c.a1);
        i = rhi32_sub(i, s2.y, c.b0);
        r = rhi32_acc(r, s2.x, c.b0);
        i = rhi32_acc(i, s2.x, c.a1);
        return {r, i};
    }
    static FFT_HD work split_dc(work x0) { return {wadd(x0.x, x0.y) >> 1, 0}; }       /* :339-340 */
    static FFT_HD work split_nyquist(work x0) { return {wsub(x0.x, x0.y) >> 1, 0}; }  /* :336-337 */
    static FFT_HD work mirror(work o) { return {o.x, wsub(0, o.y)}; }                 /* :327-329 */
    static FFT_HD work sat_shl1(work w)                                               /* arm_shift_q31(.., 1, ..) */
    {
        /* clamp, then double: two VIMNMX and an add per value (the 64-bit compare-and-select form costs five) */
        const int32_t x = w.x > 0x3fffffff ? 0x3fffffff : (w.x < -0x40000000 ? -0x40000000 : w.x);
        const int32_t y = w.y > 0x3fffffff ? 0x3fffffff : (w.y < -0x40000000 ? -0x40000000 : w.y);
        return {x * 2, y * 2};
    }
    static FFT_HD elem store_sat_shl1(work w) { return store(sat_shl1(w)); }
};

/* ------------------------------------------------------------------ q15
 *
 * int16 values are carried sign-extended in 32-bit registers, through the shared-memory
 * exchange as well (xelem = ci32: no pack/unpack between passes), and the twiddle table is
 * expanded to 32-bit pairs on upload.  The arithmetic below restates the reference's generic
 * branch (arm_cfft_radix4_q15.c:572-970, 1434-1813; arm_cfft_q15.c:782-827, 881-927) with every
 * __SSAT / (q15_t) truncation that can never fire REMOVED, which is what makes the kernel
 * affordable (the compiler does not discover these ranges):
 *   - a first stage works on inputs >> 2, |v| <= 2^13: sums of two stay below 2^14, sums of
 *     four below 2^15, so none of its __SSAT(.,16) saturates and no store wraps;
 *   - (a >> 1) +- (b >> 1) of int16 values always fits int16;
 *   - (co*x + si*y) >> 16 with int16 operands and |(co,si)| <= 2^15 is < 2^15 in magnitude and
 *     the 32-bit sum cannot overflow (|.| <= 2^15 * 2^15 * sqrt(2) < 2^31).
 * What remains are the eight saturating adds at the head of the middle and last stages, done
 * as VIADDMNMX + VIMNMX.  tests/test_emulator.py and the GPU parity tests compare against the
 * oracle on full-scale inputs (all 0x8000 / 0x7FFF / alternating) where those fire. */

/* Arithmetic right shifts on the FMA pipe.  The q15 kernels are bound by the ALU pipe (shifts, min/max, byte
 * permutes: one warp instruction per two cycles per scheduler) while the FMA pipe, just as wide, runs at a quarter
 * of that.  x >> k is the upper word of the 64-bit product x * 2^(32-k), an IMAD.WIDE -- provided the factor is not
 * a compile-time constant (the compiler turns it back into a shift), so it lives in constant memory. */
#if defined(__CUDACC__)
static __constant__ int32_t g_shr_mul[5] = {1 << 16, 1 << 30, (int32_t)0x80000000u, 1, -1};   /* >> 16, >> 2, -(x) >> 1; +1, -1 */
#endif
#if !defined(__CUDA_ARCH__)
/* host (the kernel emulator of the CPU tests): the same factors, so the emulator checks these very formulas */
static const int32_t g_shr_mul_host[5] = {1 << 16, 1 << 30, (int32_t)0x80000000u, 1, -1};
#define g_shr_mul g_shr_mul_host
#endif
/* a + b and a - b pinned to the FMA pipe (IMAD by a +-1 the compiler cannot see): ptxas balances adds between the
 * pipes by its own model and leaves a third of them on the ALU pipe, which is the one these kernels saturate */
#ifndef FFT_Q15_FMA_ADDS
#define FFT_Q15_FMA_ADDS 1           /* measured (profiles/r2_notes.md): +2..5 points of HBM peak at every length */
#endif
FFT_HD int32_t qadd(int32_t a, int32_t b) { return FFT_Q15_FMA_ADDS ? a * g_shr_mul[3] + b : a + b; }
FFT_HD int32_t qsub(int32_t a, int32_t b) { return FFT_Q15_FMA_ADDS ? b * g_shr_mul[4] + a : a - b; }
FFT_HD int32_t shr16_fma(int32_t v) { return hi32_fma(v, g_shr_mul[0]); }
FFT_HD int32_t shr2_fma(int32_t v) { return hi32_fma(v, g_shr_mul[1]); }

FFT_HD int32_t sat16(int32_t v)
{
#if defined(__CUDA_ARCH__)
    return max(-32768, min(32767, v));
#else
    return v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
#endif
}
FFT_HD int32_t sat_add16(int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    return max(__viaddmin_s32(a, b, 32767), -32768);
#else
    return sat16(a + b);
#endif
}
FFT_HD int32_t sat_sub16(int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    return max(__viaddmin_s32(a, -b, 32767), -32768);
#else
    return sat16(a - b);
#endif
}
/* A value the optimiser knows nothing about.  int16 data sign-extended into 32-bit registers keeps its 16-bit range
 * in LLVM's eyes, which then narrows the butterfly arithmetic to i16 and legalises every step with a sign-extending
 * PRMT and register moves (seen in the first stage: ~80 extra instructions per 16 points).  The mov is free. */
FFT_HD int32_t opaque32(int32_t v)
{
#if defined(__CUDA_ARCH__)
    int32_t r;
    asm("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
#else
    return v;
#endif
}
/* sat16(a + b) >> 1 and sat16(a - b) >> 1, the head of every middle / last stage butterfly (arm_cfft_radix4_q15.c:
 * 803-846), in TWO instructions each instead of three (add-and-min, max, shift).  Between stages the first operand of
 * each pair (points a and b of a butterfly; c and d are the second operands) is carried with a bias of +32768:
 *     relu(min(a' + c, 65535)) = sat16(a + c) + 32768   and   relu(min(a' - c, 65535)) = sat16(a - c) + 32768
 * are single VIADDMNMX.RELU instructions (clamp to [0, 65535] = both saturation bounds at once), and
 *     ((s + 32768) >> 1) - 16384 = s >> 1        (32768 is even: exact)
 * is one shift-and-add (LEA.HI.SX32).  The producer of a point adds the bias for free: its last instruction is a shift
 * followed by nothing ((acc >> 16) becomes (acc >> 16) + bias, again one LEA.HI.SX32).  Which points are a / b of
 * their next butterfly is a compile-time fact inside a pass and a per-thread constant across an exchange
 * (PassFix::compute, Engine::compute). */
FFT_HD int32_t relu_add_min16(int32_t a, int32_t b)          /* clamp(a + b, 0, 65535) */
{
#if defined(__CUDA_ARCH__)
    return __viaddmin_s32_relu(a, b, 65535);
#else
    const int32_t v = a + b;
    return v > 65535 ? 65535 : (v < 0 ? 0 : v);
#endif
}
/* ab = a + 32768 (biased), c plain */
FFT_HD int32_t half_sat_add16(int32_t ab, int32_t c) { return (relu_add_min16(ab, c) >> 1) - 16384; }
FFT_HD int32_t half_sat_sub16(int32_t ab, int32_t c) { return (relu_add_min16(ab, -c) >> 1) - 16384; }
FFT_HD int32_t q15w(int32_t v) { return (int32_t)(int16_t)(uint16_t)(uint32_t)v; }   /* wrap to int16, keep in a register */

/* Measured and rejected: twiddles pre-shifted by 16 and the sum taken as the upper word of two wide
 * products (no shift on the ALU pipe, which bounds this kernel).  ptxas ends such a pair with IMAD.HI, an
 * XU-pipe instruction on sm_100, and the kernels lost 5-10 % (profiles/r1_e_notes.md). */
/* (r, s) * conj(W) >> 16 (forward) / (r, s) * W >> 16 (inverse), + ob: the bias the result carries into its next
 * butterfly (0 or 32768; shift and add are one instruction).  ALUSHIFT is kept for A/B runs: the other spelling
 * takes the shift as the upper word of a product by 2^16 on the FMA pipe (IMAD.WIDE), which measured SLOWER --
 * IMAD.WIDE issues at half the rate of IMAD (tools/probes/probe_dpx.cu, profiles/r2_notes.md). */
template <bool INV, bool ALUSHIFT> FFT_HD ci32 rot_q15(int32_t x, int32_t y, ci32 w, int32_t ob)
{
    if (ALUSHIFT) {
        if (!INV) return {((w.x * x + w.y * y) >> 16) + ob, ((w.x * y - w.y * x) >> 16) + ob};
        return {((w.x * x - w.y * y) >> 16) + ob, ((w.y * x + w.x * y) >> 16) + ob};
    }
    if (!INV) return {shr16_fma(w.x * x + w.y * y) + ob, shr16_fma(w.x * y - w.y * x) + ob};
    return {shr16_fma(w.x * x - w.y * y) + ob, shr16_fma(w.y * x + w.x * y) + ob};
}
#ifndef FFT_Q15_SHIFT_MODE
#define FFT_Q15_SHIFT_MODE 0     /* 0: every rot shift on the ALU pipe (measured best), 1: FMA pipe except at the end of a pass, 2: FMA pipe always */
#endif
struct ArithQ15 {
    static constexpr bool kAluShift(bool tail) { return FFT_Q15_SHIFT_MODE == 0 || (FFT_Q15_SHIFT_MODE == 1 && tail); }
    static constexpr bool kDirectTw = FFT_FIX_DIRECT_TW_VALUE;
    static constexpr int kTableNum = 3, kTableDen = 4;
    typedef ci16 elem;
    typedef ci32 work;      /* int16 values carried sign-extended in 32-bit registers */
    typedef ci32 twid;
    typedef ci32 xelem;
    typedef ci32 telem;
    static FFT_HD work load(elem e) { return {opaque32((int32_t)e.x), opaque32((int32_t)e.y)}; }
    /* A frame's first stage shifts every input right (>> 2 radix-4, >> 1 radix-2 pre-pass): taken straight from the
     * packed 32-bit word {x, y} that shift is also the unpacking -- y >> SH = word >> (16 + SH), x >> SH =
     * (word << 16) >> (16 + SH) -- three instructions per point instead of two sign extensions and two shifts.
     * The stage is then told that its inputs are already shifted (PRE). */
    static constexpr bool kPreShift = true;
    template <int SH> static FFT_HD work load_shifted(uint32_t u)      /* u = the point as one little-endian word */
    {
        return {opaque32((int32_t)opaque32((int32_t)(u << 16)) >> (16 + SH)), opaque32((int32_t)u >> (16 + SH))};
    }
    static FFT_HD elem store(work w) { return {(int16_t)w.x, (int16_t)w.y}; }
    static FFT_HD work xload(xelem e) { return e; }
    static FFT_HD xelem xstore(work w) { return w; }
    static FFT_HD telem tw_expand(elem e) { return {(int32_t)e.x, (int32_t)e.y}; }
    static FFT_HD twid tload(telem e) { return e; }
    static FFT_HD work shl1(work w) { return {q15w((int32_t)((uint32_t)w.x << 1)), q15w((int32_t)((uint32_t)w.y << 1))}; }

    static constexpr bool kBiased = true;      /* points a, b of a middle / last stage butterfly carry +32768 (see half_sat_add16) */
    /* ob: the bias (0 or 32768) every output of this butterfly carries (the outputs of one butterfly always play the
     * same part in their next butterflies); 0 for the last stage */
    template <int KIND, bool INV, bool TAIL = false, bool PRE = false>
    static FFT_HD void bfly4(work &A, work &B, work &C, work &D, twid w1, twid w2, twid w3, int32_t ob = 0)
    {
        if (KIND == ST_FIRST4) {
            /* inputs >> 2 (plain, unbiased): nothing below can saturate or wrap (see the header of this section) */
            const int32_t T0 = PRE ? A.x : A.x >> 2, T1 = PRE ? A.y : A.y >> 2, C0 = PRE ? C.x : C.x >> 2, C1 = PRE ? C.y : C.y >> 2;
            const int32_t B0 = PRE ? B.x : B.x >> 2, B1 = PRE ? B.y : B.y >> 2, U0 = PRE ? D.x : D.x >> 2, U1 = PRE ? D.y : D.y >> 2;
            const int32_t R0 = qadd(T0, C0), R1 = qadd(T1, C1), S0 = qsub(T0, C0), S1 = qsub(T1, C1);
            const int32_t V0 = qadd(B0, U0), V1 = qadd(B1, U1), D0 = qsub(B0, U0), D1 = qsub(B1, U1);
            A = {(R0 >> 1) + (V0 >> 1) + ob, (R1 >> 1) + (V1 >> 1) + ob};
            C = rot_q15<INV, kAluShift(TAIL)>(qsub(R0, V0), qsub(R1, V1), w2, ob);
            if (!INV) {
                B = rot_q15<INV, kAluShift(TAIL)>(qadd(S0, D1), qsub(S1, D0), w1, ob);
                D = rot_q15<INV, kAluShift(TAIL)>(qsub(S0, D1), qadd(S1, D0), w3, ob);
            } else {
                B = rot_q15<INV, kAluShift(TAIL)>(qsub(S0, D1), qadd(S1, D0), w1, ob);
                D = rot_q15<INV, kAluShift(TAIL)>(qadd(S0, D1), qsub(S1, D0), w3, ob);
            }
            return;
        }
        /* middle and last stages: saturating pair sums (A, B biased; C, D plain), then everything on halved operands */
        const int32_t R0 = half_sat_add16(A.x, C.x), R1 = half_sat_add16(A.y, C.y);
        const int32_t S0 = half_sat_sub16(A.x, C.x), S1 = half_sat_sub16(A.y, C.y);
        const int32_t V0 = half_sat_add16(B.x, D.x), V1 = half_sat_add16(B.y, D.y);
        const int32_t D0 = half_sat_sub16(B.x, D.x), D1 = half_sat_sub16(B.y, D.y);
        if (KIND == ST_MID4) {
            A = {(qadd(R0, V0) >> 1) + ob, (qadd(R1, V1) >> 1) + ob};
            C = rot_q15<INV, kAluShift(TAIL)>(qsub(R0, V0), qsub(R1, V1), w2, ob);
            if (!INV) {
                B = rot_q15<INV, kAluShift(TAIL)>(qadd(S0, D1), qsub(S1, D0), w1, ob);
                D = rot_q15<INV, kAluShift(TAIL)>(qsub(S0, D1), qadd(S1, D0), w3, ob);
            } else {
                B = rot_q15<INV, kAluShift(TAIL)>(qsub(S0, D1), qadd(S1, D0), w1, ob);
                D = rot_q15<INV, kAluShift(TAIL)>(qadd(S0, D1), qsub(S1, D0), w3, ob);
            }
        } else {
            A = {qadd(R0, V0), qadd(R1, V1)};
            C = {qsub(R0, V0), qsub(R1, V1)};
            const work p = {qadd(S0, D1), qsub(S1, D0)}, q = {qsub(S0, D1), qadd(S1, D0)};
            B = INV ? q : p;
            D = INV ? p : q;
        }
    }
    /* arm_cfft_q15.c:782-800 / :881-899 */
    template <bool INV, bool PRE = false> static FFT_HD void bfly2(work &A, work &B, twid w)
    {
        const int32_t ax = PRE ? A.x : A.x >> 1, ay = PRE ? A.y : A.y >> 1, bx = PRE ? B.x : B.x >> 1, by = PRE ? B.y : B.y >> 1;
        const int32_t xt = ax - bx, yt = ay - by;
        A = {(ax + bx) >> 1, (ay + by) >> 1};
        if (!INV) B = {((xt * w.x) >> 16) + ((yt * w.y) >> 16), ((yt * w.x) >> 16) - ((xt * w.y) >> 16)};
        else      B = {((xt * w.x) >> 16) - ((yt * w.y) >> 16), ((yt * w.x) >> 16) + ((xt * w.y) >> 16)};
    }

    /* ---- real FFT (arm_rfft_q15.c generic branch): 32-bit wrap-around sums of q15 x q15 products, >> 16;
     * the results fit int16, the negated mirror is truncated by store() ---- */
    static FFT_HD int32_t sum4(int32_t p, int32_t q, int32_t u, int32_t v)
    {
        return (int32_t)((uint32_t)p + (uint32_t)q + (uint32_t)u + (uint32_t)v) >> 16;
    }
    static FFT_HD work split_fwd(work s1, work s2, ci32x4 c)          /* arm_split_rfft_q15, :364-383 */
    {
        return {sum4(s1.x * c.a0, -(s1.y * c.a1), s2.x * c.b0, s2.y * c.b1),
                sum4(s2.x * c.b1, -(s2.y * c.b0), s1.y * c.a0, s1.x * c.a1)};
    }
    static FFT_HD work split_inv(work s1, work s2, ci32x4 c)          /* arm_split_rifft_q15, :554-562 */
    {
        return {sum4(s2.x * c.b0, -(s2.y * c.b1), s1.x * c.a0, s1.y * c.a1),
                sum4(s1.y * c.a0, -(s1.x * c.a1), -(s2.x * c.b1), -(s2.y * c.b0))};
    }
    static FFT_HD work split_dc(work x0) { return {(x0.x + x0.y) >> 1, 0}; }          /* :403-404 */
    static FFT_HD work split_nyquist(work x0) { return {(x0.x - x0.y) >> 1, 0}; }     /* :400-401 */
    static FFT_HD work mirror(work o) { return {o.x, -o.y}; }                         /* :393-394 */
    static FFT_HD work sat_shl1(work w) { return {sat16(w.x * 2), sat16(w.y * 2)}; }  /* arm_shift_q15(.., 1, ..) */
    /* the same followed by the store's packing: one cvt.pack.sat (I2IP.S16.S32.SAT) saturates both halves and packs them,
     * where two clamps per value and a byte permute took five instructions per point */
    static FFT_HD elem store_sat_shl1(work w)
    {
#if defined(__CUDA_ARCH__)
        uint32_t d;
        asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(w.y * 2), "r"(w.x * 2));      /* first source -> upper half */
        return {(int16_t)(d & 0xffffu), (int16_t)(d >> 16)};
#else
        return store(sat_shl1(w));
#endif
    }
};

/* ------------------------------------------------------------------ f64
 *
 * arm_cfft_f64 (Source/TransformFunctions/arm_cfft_f64.c) is a radix-4 DIF transform with the stage
 * structure of the fixed-point path -- arm_radix4_butterfly_f64 (:58-183) for N = 4^m, one radix-2
 * pre-pass and two half-size radix-4 transforms for N = 2*4^m (arm_cfft_radix4by2_f64, :193-239), results
 * in binary bit-reversed order (its armBitRevIndexTableF64_N are the fixed-point swap lists) -- and no
 * scaling, so it runs on the PassFix passes with this Arith: the reference's butterfly, operation for
 * operation: same operand order, and products rounded on their own (__dmul_rn: no FMA contraction), so
 * that with the same twiddle values the results are bit-identical to the reference's generic-C build
 * (-ffp-contract=off) -- the f64 kernels have the arithmetic headroom (16 bytes per point).  The inverse is conjugate -> forward -> conjugate / N (:262-312), done by CfftBody at the load and
 * the store like f32, so the butterflies have no inverse variant. */
struct ArithF64 {
    static constexpr bool kBiased = false;       /* see ArithQ15::bfly4 */
    static constexpr bool kPreShift = false;     /* see ArithQ15::load_shifted */
    static constexpr bool kDirectTw = false;
    static constexpr int kTableNum = 1, kTableDen = 1;        /* reference table: N entries */
    typedef cf64 elem;
    typedef cf64 work;
    typedef cf64 twid;
    typedef cf64 xelem;
    typedef cf64 telem;
    static FFT_HD work load(elem e) { return e; }
    static FFT_HD elem store(work w) { return w; }
    static FFT_HD work xload(xelem e) { return e; }
    static FFT_HD xelem xstore(work w) { return w; }
    static FFT_HD telem tw_expand(elem e) { return e; }
    static FFT_HD twid tload(telem e) { return e; }
    static FFT_HD work shl1(work w) { return w; }
    static FFT_HD double mul(double a, double b)
    {
#if defined(__CUDA_ARCH__)
        return __dmul_rn(a, b);
#else
        return a * b;
#endif
    }
    /* (r, s) * conj(W), W = (cos, +sin): arm_cfft_f64.c:139-143 */
    static FFT_HD work rot(double r, double s, twid w) { return {mul(r, w.x) + mul(s, w.y), mul(s, w.x) - mul(r, w.y)}; }

    /* outputs in residue order (a', b'[W^1], c'[W^2], d'[W^3]); arm_cfft_f64.c:108-170.  The last stage's
     * twiddles are W^0 = (1, 0) (:94-99 with ia1 = 0): the products are skipped. */
    template <int KIND, bool INV, bool TAIL = false, bool PRE = false>
    static FFT_HD void bfly4(work &A, work &B, work &C, work &D, twid w1, twid w2, twid w3, int32_t = 0)
    {
        double r1 = A.x + C.x, r2 = A.x - C.x, s1 = A.y + C.y, s2 = A.y - C.y;
        double t1 = B.x + D.x;
        const double ax = r1 + t1;
        r1 = r1 - t1;
        double t2 = B.y + D.y;
        const double ay = s1 + t2;
        s1 = s1 - t2;
        t1 = B.y - D.y;
        t2 = B.x - D.x;
        const work oc = (KIND == ST_LAST4) ? work{r1, s1} : rot(r1, s1, w2);
        r1 = r2 + t1;
        r2 = r2 - t1;
        s1 = s2 - t2;
        s2 = s2 + t2;
        const work ob = (KIND == ST_LAST4) ? work{r1, s1} : rot(r1, s1, w1);
        const work od = (KIND == ST_LAST4) ? work{r2, s2} : rot(r2, s2, w3);
        A = {ax, ay}; B = ob; C = oc; D = od;
    }
    /* radix-2 pre-pass of the N = 2*4^m lengths (arm_cfft_f64.c:205-230) */
    template <bool INV, bool PRE = false> static FFT_HD void bfly2(work &A, work &B, twid w)
    {
        const double a0 = A.x + B.x, xt = A.x - B.x, yt = A.y - B.y, a1 = B.y + A.y;
        A = {a0, a1};
        B = rot(xt, yt, w);
    }
};

/* the same arithmetic reading the reference-layout twiddle table (PassFix::kDirect, fft_frame.cuh): used for
 * N >= 2048, where the pass-ordered copy (16-byte entries) no longer fits next to the exchange buffers in L1.
 * Measured on B200 (profiles/r1_f_notes.md): N = 2048 68.6 -> 84.4 %, N = 4096 56.5 -> 68.7 % of the HBM peak;
 * N <= 1024 is 0-2 points faster with the pass-ordered copy.  The same switch for q31 / q15 (FFT_FIX_DIRECT_TW,
 * 8-byte entries, tables a third the size) measured 1-4 points SLOWER at every length and stays off. */
struct ArithF64D : ArithF64 {
#if defined(FFT_F64_ORDERED_TW)
    static constexpr bool kDirectTw = false;
#else
    static constexpr bool kDirectTw = true;
#endif
};

}  // namespace b200fft
