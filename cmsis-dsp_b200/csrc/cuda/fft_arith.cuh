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

/* ------------------------------------------------------------------ f32 */

FFT_HD cf32 cadd(cf32 a, cf32 b) { return {a.x + b.x, a.y + b.y}; }
FFT_HD cf32 csub(cf32 a, cf32 b) { return {a.x - b.x, a.y - b.y}; }
/* a * (-i) and a * (+i) */
FFT_HD cf32 mul_mi(cf32 a) { return {a.y, -a.x}; }
FFT_HD cf32 mul_pi(cf32 a) { return {-a.y, a.x}; }
/* a * conj(w), w = (cos, +sin) as stored in the reference's twiddle tables */
FFT_HD cf32 mul_conj(cf32 a, cf32 w)
{
#if defined(__CUDA_ARCH__)
    return {__fmaf_rn(a.x, w.x, a.y * w.y), __fmaf_rn(a.y, w.x, -(a.x * w.y))};
#else
    return {a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y};
#endif
}
/* a * (c - i s) with compile-time style constants */
FFT_HD cf32 mul_cs(cf32 a, float c, float s) { return mul_conj(a, cf32{c, s}); }

template <int R> struct DftF32;

template <> struct DftF32<2> {
    static FFT_HD void run(cf32 *x)
    {
        cf32 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};

template <> struct DftF32<4> {
    static FFT_HD void run(cf32 *x) { run(x[0], x[1], x[2], x[3]); }
    static FFT_HD void run(cf32 &x0, cf32 &x1, cf32 &x2, cf32 &x3)
    {
        cf32 a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3), a3 = csub(x1, x3);
        x0 = cadd(a0, a2);
        x2 = csub(a0, a2);
        x1 = {a1.x + a3.y, a1.y - a3.x};
        x3 = {a1.x - a3.y, a1.y + a3.x};
    }
};

template <> struct DftF32<8> {
    static FFT_HD void run(cf32 *x)
    {
        const float h = 0.70710678118654752f;
        cf32 b0 = cadd(x[0], x[4]), c0 = csub(x[0], x[4]);
        cf32 b1 = cadd(x[1], x[5]), c1 = csub(x[1], x[5]);
        cf32 b2 = cadd(x[2], x[6]), c2 = csub(x[2], x[6]);
        cf32 b3 = cadd(x[3], x[7]), c3 = csub(x[3], x[7]);
        c1 = {(c1.x + c1.y) * h, (c1.y - c1.x) * h};       /* * w8^1 */
        c2 = mul_mi(c2);                                    /* * w8^2 */
        c3 = {(c3.y - c3.x) * h, -(c3.x + c3.y) * h};      /* * w8^3 */
        DftF32<4>::run(b0, b1, b2, b3);
        DftF32<4>::run(c0, c1, c2, c3);
        x[0] = b0; x[2] = b1; x[4] = b2; x[6] = b3;
        x[1] = c0; x[3] = c1; x[5] = c2; x[7] = c3;
    }
};

template <> struct DftF32<16> {
    static FFT_HD void run(cf32 *x)
    {
        const float h = 0.70710678118654752f, c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
        /* step 1: four DFT4 over q2 for each q1 (elements q1 + 4 q2) */
        DftF32<4>::run(x[0], x[4], x[8], x[12]);
        DftF32<4>::run(x[1], x[5], x[9], x[13]);
        DftF32<4>::run(x[2], x[6], x[10], x[14]);
        DftF32<4>::run(x[3], x[7], x[11], x[15]);
        /* now x[q1 + 4 r2] holds z[q1][r2]; multiply by w16^(q1*r2) */
        x[5]  = mul_cs(x[5], c1, s1);                        /* 1 */
        x[9]  = {(x[9].x + x[9].y) * h, (x[9].y - x[9].x) * h};      /* 2 */
        x[13] = mul_cs(x[13], s1, c1);                       /* 3 */
        x[6]  = {(x[6].x + x[6].y) * h, (x[6].y - x[6].x) * h};      /* 2 */
        x[10] = mul_mi(x[10]);                               /* 4 */
        x[14] = {(x[14].y - x[14].x) * h, -(x[14].x + x[14].y) * h}; /* 6 */
        x[7]  = mul_cs(x[7], s1, c1);                        /* 3 */
        x[11] = {(x[11].y - x[11].x) * h, -(x[11].x + x[11].y) * h}; /* 6 */
        x[15] = mul_cs(x[15], -c1, -s1);                     /* 9 */
        /* step 2: for each r2, DFT4 over q1 -> outputs t = r2 + 4 t2 */
        DftF32<4>::run(x[0], x[1], x[2], x[3]);
        DftF32<4>::run(x[4], x[5], x[6], x[7]);
        DftF32<4>::run(x[8], x[9], x[10], x[11]);
        DftF32<4>::run(x[12], x[13], x[14], x[15]);
        /* x[4 r2 + t2] = y[r2 + 4 t2]: transpose to natural output order */
        cf32 t;
        t = x[1];  x[1]  = x[4];  x[4]  = t;
        t = x[2];  x[2]  = x[8];  x[8]  = t;
        t = x[3];  x[3]  = x[12]; x[12] = t;
        t = x[6];  x[6]  = x[9];  x[9]  = t;
        t = x[7];  x[7]  = x[13]; x[13] = t;
        t = x[11]; x[11] = x[14]; x[14] = t;
    }
};

struct ArithF32 {
    typedef cf32 elem;
    typedef cf32 work;
    typedef cf32 twid;
    static FFT_HD work load(elem e) { return e; }
    static FFT_HD elem store(work w) { return w; }
    static FFT_HD work shl1(work w) { return w; }
};

/* ------------------------------------------------------------------ q31 */

FFT_HD int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
FFT_HD int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
FFT_HD int32_t wshl1(int32_t a) { return (int32_t)((uint32_t)a << 1); }
FFT_HD int32_t hi32(int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    return __mulhi(a, b);
#else
    return (int32_t)(((int64_t)a * b) >> 32);
#endif
}
/* SMMULR / SMMLAR / SMMLSR (none.h:185-194) */
FFT_HD int32_t rhi32(int32_t x, int32_t y) { return (int32_t)(((int64_t)x * y + 0x80000000LL) >> 32); }
FFT_HD int32_t rhi32_acc(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) + (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
FFT_HD int32_t rhi32_sub(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) - (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}

enum StageKind { ST_PRE2 = 0, ST_FIRST4 = 1, ST_MID4 = 2, ST_LAST4 = 3 };

template <bool INV> FFT_HD ci32 rot_q31(int32_t R, int32_t S, ci32 w)
{
    if (!INV) return {wadd(hi32(R, w.x), hi32(S, w.y)), wsub(hi32(S, w.x), hi32(R, w.y))};
    return {wsub(hi32(R, w.x), hi32(S, w.y)), wadd(hi32(S, w.x), hi32(R, w.y))};
}

struct ArithQ31 {
    typedef ci32 elem;      /* storage element */
    typedef ci32 work;      /* register element */
    typedef ci32 twid;
    static FFT_HD work load(elem e) { return e; }
    static FFT_HD elem store(work w) { return w; }
    static FFT_HD work shl1(work w) { return {wshl1(w.x), wshl1(w.y)}; }

    /* radix-4 DIF stage butterfly; outputs in residue order (a', b'[W^1], c'[W^2], d'[W^3]).
     * w1/w2/w3 = table entries (cos,+sin) of W^1, W^2, W^3 for this butterfly. */
    template <int KIND, bool INV>
    static FFT_HD void bfly4(work &A, work &B, work &C, work &D, twid w1, twid w2, twid w3)
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
    template <bool INV> static FFT_HD void bfly2(work &A, work &B, twid w)
    {
        work a = A, b = B;
        int32_t xt = wsub(a.x >> 2, b.x >> 2), yt = wsub(a.y >> 2, b.y >> 2);
        A = {wadd(a.x >> 2, b.x >> 2), wadd(b.y >> 2, a.y >> 2)};
        int32_t p0 = rhi32(xt, w.x), p1 = rhi32(yt, w.x);
        if (!INV) { p0 = rhi32_acc(p0, yt, w.y); p1 = rhi32_sub(p1, xt, w.y); }
        else      { p0 = rhi32_sub(p0, yt, w.y); p1 = rhi32_acc(p1, xt, w.y); }
        B = {wshl1(p0), wshl1(p1)};
    }
};

/* ------------------------------------------------------------------ q15 */

FFT_HD int32_t sat16(int32_t v)
{
#if defined(__CUDA_ARCH__)
    return max(-32768, min(32767, v));
#else
    return v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
#endif
}
FFT_HD int32_t q15w(int32_t v) { return (int32_t)(int16_t)(uint16_t)(uint32_t)v; }   /* wrap to int16, keep in a register */

template <bool INV> FFT_HD ci32 rot_q15(int32_t x, int32_t y, ci32 w)
{
    if (!INV) return {q15w((w.x * x + w.y * y) >> 16), q15w((-w.y * x + w.x * y) >> 16)};
    return {q15w((w.x * x - w.y * y) >> 16), q15w((w.y * x + w.x * y) >> 16)};
}

struct ArithQ15 {
    typedef ci16 elem;
    typedef ci32 work;      /* int16 values carried sign-extended in 32-bit registers */
    typedef ci32 twid;
    static FFT_HD work load(elem e) { return {(int32_t)e.x, (int32_t)e.y}; }
    static FFT_HD elem store(work w) { return {(int16_t)w.x, (int16_t)w.y}; }
    static FFT_HD work shl1(work w) { return {q15w((int32_t)((uint32_t)w.x << 1)), q15w((int32_t)((uint32_t)w.y << 1))}; }

    template <int KIND, bool INV>
    static FFT_HD void bfly4(work &A, work &B, work &C, work &D, twid w1, twid w2, twid w3)
    {
        const int sh = (KIND == ST_FIRST4) ? 2 : 0;
        int32_t T0 = A.x >> sh, T1 = A.y >> sh, S0 = C.x >> sh, S1 = C.y >> sh;
        int32_t B0 = B.x >> sh, B1 = B.y >> sh, U0 = D.x >> sh, U1 = D.y >> sh;
        int32_t R0 = sat16(T0 + S0), R1 = sat16(T1 + S1);
        S0 = sat16(T0 - S0); S1 = sat16(T1 - S1);
        T0 = sat16(B0 + U0); T1 = sat16(B1 + U1);
        int32_t D0 = sat16(B0 - U0), D1 = sat16(B1 - U1);
        if (KIND == ST_FIRST4) {
            A = {q15w((R0 >> 1) + (T0 >> 1)), q15w((R1 >> 1) + (T1 >> 1))};
            R0 = sat16(R0 - T0); R1 = sat16(R1 - T1);
            C = rot_q15<INV>(R0, R1, w2);
            int32_t P0, P1, Q0, Q1;
            if (!INV) { P0 = sat16(S0 + D1); P1 = sat16(S1 - D0); Q0 = sat16(S0 - D1); Q1 = sat16(S1 + D0); }
            else      { P0 = sat16(S0 - D1); P1 = sat16(S1 + D0); Q0 = sat16(S0 + D1); Q1 = sat16(S1 - D0); }
            B = rot_q15<INV>(P0, P1, w1);
            D = rot_q15<INV>(Q0, Q1, w3);
        } else if (KIND == ST_MID4) {
            A = {q15w(((R0 >> 1) + (T0 >> 1)) >> 1), q15w(((R1 >> 1) + (T1 >> 1)) >> 1)};
            R0 = (R0 >> 1) - (T0 >> 1); R1 = (R1 >> 1) - (T1 >> 1);
            C = rot_q15<INV>(R0, R1, w2);
            int32_t P0, P1, Q0, Q1;
            if (!INV) { P0 = (S0 >> 1) + (D1 >> 1); P1 = (S1 >> 1) - (D0 >> 1); Q0 = (S0 >> 1) - (D1 >> 1); Q1 = (S1 >> 1) + (D0 >> 1); }
            else      { P0 = (S0 >> 1) - (D1 >> 1); P1 = (S1 >> 1) + (D0 >> 1); Q0 = (S0 >> 1) + (D1 >> 1); Q1 = (S1 >> 1) - (D0 >> 1); }
            B = rot_q15<INV>(P0, P1, w1);
            D = rot_q15<INV>(Q0, Q1, w3);
        } else {
            A = {q15w((R0 >> 1) + (T0 >> 1)), q15w((R1 >> 1) + (T1 >> 1))};
            C = {q15w((R0 >> 1) - (T0 >> 1)), q15w((R1 >> 1) - (T1 >> 1))};
            work p = {q15w((S0 >> 1) + (D1 >> 1)), q15w((S1 >> 1) - (D0 >> 1))};
            work q = {q15w((S0 >> 1) - (D1 >> 1)), q15w((S1 >> 1) + (D0 >> 1))};
            B = INV ? q : p;
            D = INV ? p : q;
        }
    }
    /* arm_cfft_q15.c:782-800 / :881-899 */
    template <bool INV> static FFT_HD void bfly2(work &A, work &B, twid w)
    {
        work a = A, b = B;
        int32_t xt = q15w((a.x >> 1) - (b.x >> 1)), yt = q15w((a.y >> 1) - (b.y >> 1));
        A = {q15w(((a.x >> 1) + (b.x >> 1)) >> 1), q15w(((b.y >> 1) + (a.y >> 1)) >> 1)};
        if (!INV)
            B = {q15w(q15w((xt * w.x) >> 16) + q15w((yt * w.y) >> 16)), q15w(q15w((yt * w.x) >> 16) - q15w((xt * w.y) >> 16))};
        else
            B = {q15w(q15w((xt * w.x) >> 16) - q15w((yt * w.y) >> 16)), q15w(q15w((yt * w.x) >> 16) + q15w((xt * w.y) >> 16))};
    }
};

}  // namespace b200fft
