/*
 * fft_body.cuh -- kernel bodies (host/device) built from the Engine: the phases of one
 * frame, with the global-memory prologue/epilogue of each entry point.
 *
 *   CfftBody      arm_cfft_{f32,q31,q15}: load frame -> passes -> store frame (in place)
 *                 reference: arm_cfft_f32.c:1243-1298, arm_cfft_q31.c:704-755, arm_cfft_q15.c:671-722
 *   RfftFwdBody   arm_rfft_fast_f32 forward: N/2-point CFFT + split stage fused in the
 *                 epilogue (arm_rfft_fast_f32.c:316-402,694-697)
 *   RfftInvBody   arm_rfft_fast_f32 inverse: merge stage fused in the prologue + inverse
 *                 N/2-point CFFT (arm_rfft_fast_f32.c:405-462,686-690)
 *
 * A body is a list of PHASES; the device kernel runs them with __syncthreads() between,
 * the CPU emulator (tests/emu) runs each phase for every thread in turn.
 */
#pragma once
#include <math.h>
#include "fft_frame.cuh"

namespace b200fft {

#if defined(__CUDA_ARCH__)
template <class V> FFT_HD V ld_stream(const V *p);
template <> FFT_HD cf32 ld_stream<cf32>(const cf32 *p) { float2 v = __ldcs(reinterpret_cast<const float2 *>(p)); return {v.x, v.y}; }
template <> FFT_HD ci32 ld_stream<ci32>(const ci32 *p) { int2 v = __ldcs(reinterpret_cast<const int2 *>(p)); return {v.x, v.y}; }
template <> FFT_HD ci16 ld_stream<ci16>(const ci16 *p) { short2 v = __ldcs(reinterpret_cast<const short2 *>(p)); return {v.x, v.y}; }
template <> FFT_HD cf64 ld_stream<cf64>(const cf64 *p) { double2 v = __ldcs(reinterpret_cast<const double2 *>(p)); return {v.x, v.y}; }
FFT_HD void st_stream(cf32 *p, cf32 v) { __stcs(reinterpret_cast<float2 *>(p), make_float2(v.x, v.y)); }
FFT_HD void st_stream(cf64 *p, cf64 v) { __stcs(reinterpret_cast<double2 *>(p), make_double2(v.x, v.y)); }
FFT_HD void st_stream(ci32 *p, ci32 v) { __stcs(reinterpret_cast<int2 *>(p), make_int2(v.x, v.y)); }
FFT_HD void st_stream(ci16 *p, ci16 v) { __stcs(reinterpret_cast<short2 *>(p), make_short2(v.x, v.y)); }
/* predicated streaming store: no branch, so the surrounding straight-line code stays one block */
FFT_HD void st_stream_if(bool pred, cf32 *p, cf32 v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.cs.v2.f32 [%0], {%1, %2};\n\t}"
                 ::"l"(p), "f"(v.x), "f"(v.y), "r"((int)pred));
}
#else
FFT_HD void st_stream_if(bool pred, cf32 *p, cf32 v) { if (pred) *p = v; }
template <class V> FFT_HD V ld_stream(const V *p) { return *p; }
template <class V> FFT_HD void st_stream(V *p, V v) { *p = v; }
#endif

/* a q15 point as one 32-bit word {x, y} (ArithQ15::load_shifted) */
FFT_HD uint32_t ld_word(const ci16 *p)
{
#if defined(__CUDA_ARCH__)
    return __ldcs(reinterpret_cast<const unsigned int *>(p));
#else
    return (uint32_t)(uint16_t)p->x | ((uint32_t)(uint16_t)p->y << 16);
#endif
}
template <class V> FFT_HD uint32_t ld_word(const V *) { return 0; }     /* other element types never take this path */

/* frame output: STAGED = the result goes into the frame's image in shared memory (a bulk copy
 * takes it home); otherwise straight to global memory with a streaming hint */
template <bool STAGED, class V> FFT_HD void st_out(V *p, V v)
{
    if (STAGED) *p = v;
    else st_stream(p, v);
}
template <bool STAGED> FFT_HD void st_out_if(bool pred, cf32 *p, cf32 v)
{
    if (STAGED) {
        if (pred) *p = v;
    } else {
        st_stream_if(pred, p, v);
    }
}
/* frame input: STAGED = the frame was brought into shared memory by a bulk (TMA) copy and `p`
 * points there; otherwise `p` is global memory and is read once with a streaming hint */
template <bool STAGED, class V> FFT_HD V ld_in(const V *p)
{
    if (STAGED) return *p;
    return ld_stream(p);
}

/* plain binary bit reversal of k, 0 <= k < n (n a power of two) */
FFT_HD constexpr int bitrev_const(int k, int n)
{
    int r = 0;
    for (int m = n >> 1; m; m >>= 1, k >>= 1) r = (r << 1) | (k & 1);
    return r;
}

/* number of phases for a plan with NP passes: 1 -> 1, 2 -> 2, 3 -> 4 (the middle pass is split
 * into load+compute / store so the single exchange buffer can be reused) */
template <int NP> struct PhaseCount { static constexpr int value = (NP == 1) ? 1 : (NP == 2 ? 2 : 4); };

template <class T> struct IsF32 { static constexpr bool value = false; };
template <> struct IsF32<cf32> { static constexpr bool value = true; };
/* floating-point element: the inverse is conjugate -> forward -> conjugate / N, no fixed-point << 1 */
template <class T> struct IsFloat { static constexpr bool value = IsF32<T>::value; };
template <> struct IsFloat<cf64> { static constexpr bool value = true; };

/* ------------------------------------------------------------------ CFFT */

/* PERM: bitReverseFlag == 0, results are scattered through the plan's output permutation
 * (a compile-time flavour: a run-time test would put a predicated table load and its
 * scoreboard wait in front of every store of the common natural-order case) */
/* RIFFT: the body is the core of the fixed-point inverse real FFT (arm_rfft_q31.c:160-167,
 * arm_rfft_q15.c:162-169): the load is the merge stage arm_split_rifft_* applied to the bins
 * X[k], X[N-k] of a 2N-bin spectrum frame, the store ends with arm_shift_*(.., 1, ..) */
/* WIN (f32): every input sample is first multiplied by a real window value (arm_cmplx_mult_real_f32 /
 * arm_mult_f32 fused into the load: the windowed frame never exists in memory) */
template <class PL, bool INV, bool PERM = false, bool STAGED = false, bool RIFFT = false, bool WIN = false> struct CfftBody {
    typedef Engine<PL> Eng;
    typedef typename Eng::A A;
    typedef typename A::elem elem;
    typedef typename A::xelem xelem;
    typedef typename A::telem telem;
    typedef typename A::work work;
    typedef typename Eng::Regs Regs;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N;
    static constexpr int kPhases = PhaseCount<NP>::value;
    static constexpr bool kF32 = IsFloat<elem>::value;      /* f32 or f64 */
    /* q15: the first stage's input shift is taken while unpacking the loaded word (ArithQ15::load_shifted) */
    static constexpr bool kPre = A::kPreShift && !RIFFT;

    struct Args {
        const elem *in;          /* frame base (device or emulated) */
        elem *out;               /* may alias in */
        const telem *tw;         /* pass-ordered twiddle table of this plan (Plan::build_twiddles) */
        const uint16_t *perm;    /* PERM only: destination position of X[k] */
        float scale;             /* f32 inverse: 1/N */
        int shl1;                /* fixed point, N = 2*4^m: final << 1 of every word (== PL::kOddLog2: the kernels use the constant) */
        const ci32x4 *coef;      /* RIFFT only: split-stage coefficients of bins 0..N-1 */
        const float *win;        /* WIN only: N window values */
    };
    /* frame-local view of the batch arguments; with STAGED the kernel overrides `in` */
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)(RIFFT ? 2 * N : N);
        a.out += frame * (uint64_t)N;
        return a;
    }
    static constexpr bool kStaged = STAGED, kInv = INV;

    static FFT_HD void gload(Regs &r, const Args &a, int i)
    {
        typedef typename PL::P0 PS;
#pragma unroll
        for (int b = 0; b < E / PS::R; b++)
#pragma unroll
            for (int e = 0; e < PS::R; e++) {
                const int idx = Eng::template in_index<0>(i, b, e);
                work w;
                if constexpr (RIFFT) {
                    /* every bin is read twice (as X[k] and as X[N-k]): plain cached loads */
                    w = A::split_inv(A::load(a.in[idx]), A::load(a.in[N - idx]), a.coef[idx]);
                } else if constexpr (kPre) {
                    w = A::template load_shifted<PS::kInShift>(ld_word(a.in + idx));
                } else {
                    w = A::load(ld_in<STAGED>(a.in + idx));
                }
                if constexpr (WIN) w = win_mul(w, a.win[idx]);
                if (kF32 && INV) w.y = -w.y;                       /* conjugate input (cfft_f32.c:1252-1261) */
                r.v[b * PS::R + e] = w;
            }
    }
    static FFT_HD cf32 win_mul(cf32 w, float v) { return {w.x * v, w.y * v}; }
    template <class W> static FFT_HD W win_mul(W w, float) { return w; }
    static FFT_HD void gstore(const Regs &r, const Args &a, int i)
    {
        typedef typename PassOf<PL, NP - 1>::type PS;
#pragma unroll
        for (int b = 0; b < E / PS::R; b++)
#pragma unroll
            for (int e = 0; e < PS::R; e++) {
                const int k = Eng::template out_index<NP - 1>(i, b, e);
                work w = r.v[b * PS::R + e];
                if (kF32) {
                    if (INV) w = scale_conj(w, a.scale);             /* cfft_f32.c:1285-1297 */
                } else if (PL::kOddLog2) {
                    w = A::shl1(w);                                  /* cfft_q31.c:803-820, cfft_q15.c:810-827 */
                }
                int pos = k;
                if constexpr (PERM) pos = (int)a.perm[k];
                if constexpr (RIFFT) st_stream(a.out + pos, A::store_sat_shl1(w));      /* arm_shift_*(.., 1, ..) of arm_rfft_* (inverse) */
                else st_stream(a.out + pos, A::store(w));
            }
    }
    static FFT_HD cf32 scale_conj(cf32 w, float s) { return {w.x * s, -w.y * s}; }
    static FFT_HD cf64 scale_conj(cf64 w, float s) { return {w.x * (double)s, -w.y * (double)s}; }   /* cfft_f64.c:298-310 */
    static FFT_HD ci32 scale_conj(ci32 w, float) { return w; }

    /* phase 0 in two halves, for kernels whose input buffer doubles as the exchange buffer */
    static constexpr bool kHasPre = false, kHasPost = false;
    static constexpr bool kImageOut = false;     /* pipelined kernel: results leave through the registers */
    static FFT_HD void set_scratch(Args &, xelem *) {}
    static FFT_HD void pre(const Args &, xelem *, int) {}
    static FFT_HD void post(const Args &, xelem *, int) {}
    static FFT_HD void phase0_in(Regs &r, const Args &a, xelem *, int i)
    {
        gload(r, a, i);
        Eng::template compute<0, INV, kPre>(r, a.tw, i);
    }
    /* Table values of this thread that do not depend on the frame (Hoist): the persistent kernel
     * fetches them once and parks them in shared memory, column `pk` (stride PL::kThreads); the
     * *_pk variants of the phases read them from there at the point of use. */
    struct Hoist { typename Eng::template TwRegs<0> t0; };
    static constexpr int kH0 = (int)(sizeof(typename Eng::template TwRegs<0>) / sizeof(telem));
    static FFT_HD void hoist(Hoist &h, const Args &a, int i) { Eng::template load_tw<0>(h.t0, a.tw, i); }
    static FFT_HD void phase0_in_pk(Regs &r, const Args &a, xelem *, int i, const telem *pk)
    {
        typename Eng::template TwRegs<0> t;
#pragma unroll
        for (int s = 0; s < kH0; s++) t.w[s] = pk[s * PL::kThreads];
        gload(r, a, i);
        Eng::template compute_pre<0, INV>(r, t);
    }
    static FFT_HD void post_pk(const Args &, xelem *, int, const telem *) {}
    static FFT_HD void pre_pk(const Args &, xelem *, int, const telem *) {}
    static FFT_HD void last_out_pk(Regs &r, const Args &a, int i, const telem *) { last_out(r, a, i); }
    static FFT_HD void phase0_out(const Regs &r, xelem *sm, int i) { Eng::template smem_store<0>(r, sm, i); }
    /* last phase of a two-pass plan */
    static FFT_HD void last_in(Regs &r, const xelem *sm, int i) { Eng::template smem_load<1>(r, sm, i); }
    static FFT_HD void last_out(Regs &r, const Args &a, int i)
    {
        Eng::template compute<1, INV>(r, a.tw, i);
        gstore(r, a, i);
    }
    static FFT_HD void last(Regs &r, const Args &a, xelem *sm, int i)
    {
        last_in(r, sm, i);
        last_out(r, a, i);
    }

    /* HOLD: the results of the last pass stay in the registers (no store): a fused epilogue follows */
    template <int PH, bool HOLD = false> static FFT_HD void phase(Regs &r, const Args &a, xelem *sm, int i)
    {
        if constexpr (PH == 0) {
            gload(r, a, i);
            Eng::template compute<0, INV, kPre>(r, a.tw, i);
            if constexpr (NP == 1) {
                if constexpr (!HOLD) gstore(r, a, i);
            } else {
                Eng::template smem_store<0>(r, sm, i);
            }
        } else if constexpr (NP == 2) {
            Eng::template smem_load<1>(r, sm, i);
            Eng::template compute<1, INV>(r, a.tw, i);
            if constexpr (!HOLD) gstore(r, a, i);
        } else if constexpr (NP == 3) {
            if constexpr (PH == 1) {
                Eng::template smem_load<1>(r, sm, i);
                Eng::template compute<1, INV>(r, a.tw, i);
            } else if constexpr (PH == 2) {
                Eng::template smem_store<1>(r, sm, i);
            } else {
                Eng::template smem_load<2>(r, sm, i);
                Eng::template compute<2, INV>(r, a.tw, i);
                if constexpr (!HOLD) gstore(r, a, i);
            }
        }
    }
};

/* ------------------------------------------------------------------ CFFT with a fused spectrum epilogue (f32)
 *
 * arm_cfft_f32 followed by arm_cmplx_mag_f32 / arm_cmplx_mag_squared_f32 (ComplexMathFunctions/arm_cmplx_mag_f32.c:
 * 252-264, arm_cmplx_mag_squared_f32.c) and, for SPEC_PEAK, arm_max_f32 (StatisticsFunctions/arm_max_f32.c: the first
 * maximum wins) -- the pipeline of Examples/ARM/arm_fft_bin_example/arm_fft_bin_example_f32.c:141-149.  The spectrum
 * never reaches HBM: a frame costs 8N bytes in and 4N bytes (magnitudes) or 8 bytes (peak value and index) out. */
enum SpectrumMode { SPEC_MAG = 0, SPEC_MAG_SQUARED = 1, SPEC_PEAK = 2 };

#if defined(__CUDA_ARCH__)
FFT_HD void st_stream(float *p, float v) { __stcs(p, v); }
FFT_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
FFT_HD float add_rn(float a, float b) { return __fadd_rn(a, b); }
#else
FFT_HD void st_stream(float *p, float v) { *p = v; }
FFT_HD float mul_rn(float a, float b) { return a * b; }
FFT_HD float add_rn(float a, float b) { return a + b; }
#endif

template <class PL, bool INV, int MODE, bool STAGED = false> struct CfftMagBody : CfftBody<PL, INV, false, STAGED> {
    typedef CfftBody<PL, INV, false, STAGED> C;
    typedef typename C::Eng Eng;
    typedef typename C::Regs Regs;
    typedef cf32 elem;
    typedef cf32 xelem;
    typedef cf32 telem;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;
    static_assert(IsF32<typename PL::Arith::elem>::value, "spectrum epilogues are f32");
    /* SPEC_PEAK with a frame that spans several warps: the warps' partial peaks meet in the exchange buffer */
    static constexpr bool kCross = (MODE == SPEC_PEAK) && (T > 32);
    static constexpr int kCfftPhases = PhaseCount<NP>::value;
    static constexpr int kPhases = kCfftPhases + (kCross ? 2 : 0);

    struct Args : C::Args {
        float *mag;              /* SPEC_MAG / SPEC_MAG_SQUARED: N floats per frame */
        float *peakVal;          /* SPEC_PEAK: one value and one index per frame */
        uint32_t *peakIdx;
    };
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        if (MODE == SPEC_PEAK) {
            a.peakVal += frame;
            a.peakIdx += frame;
        } else {
            a.mag += frame * (uint64_t)N;
        }
        return a;
    }
    static FFT_HD void set_scratch(Args &, xelem *) {}

    static FFT_HD float magnitude(cf32 w, float scale)
    {
        if (INV) w = cf32{w.x * scale, -w.y * scale};                       /* cfft_f32.c:1285-1297 */
        const float s = add_rn(mul_rn(w.x, w.x), mul_rn(w.y, w.y));         /* (real * real) + (imag * imag), no contraction */
        /* SPEC_PEAK compares squared magnitudes (sqrt is monotonic) and takes the root of the winner only */
        if (MODE != SPEC_MAG) return s;
#if defined(__CUDA_ARCH__)
        return __fsqrt_rn(s);
#else
        return sqrtf(s);
#endif
    }
    static FFT_HD bool better(float v1, int k1, float v2, int k2) { return v1 > v2 || (v1 == v2 && k1 < k2); }

    /* the registers hold the results of the last pass */
    static FFT_HD void epilogue(Regs &r, const Args &a, int i)
    {
        typedef typename PassOf<PL, NP - 1>::type PS;
        float bv = 0.0f;
        int bk = 0x7fffffff;
#pragma unroll
        for (int b = 0; b < E / PS::R; b++)
#pragma unroll
            for (int e = 0; e < PS::R; e++) {
                const int k = Eng::template out_index<NP - 1>(i, b, e);
                const float m = magnitude(r.v[b * PS::R + e], a.scale);
                if (MODE == SPEC_PEAK) {
                    if ((b == 0 && e == 0) || better(m, k, bv, bk)) { bv = m; bk = k; }
                } else {
                    st_stream(a.mag + k, m);
                }
            }
        if (MODE == SPEC_PEAK) {
#if defined(__CUDA_ARCH__)
#pragma unroll
            for (int o = (T < 32 ? T : 32) / 2; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (better(ov, ok, bv, bk)) { bv = ov; bk = ok; }
            }
            if (kCross) {
                r.v[0] = cf32{bv, __int_as_float(bk)};                      /* the warp's partial, for the two extra phases */
            } else if (i == 0) {
                *a.peakVal = __fsqrt_rn(bv);
                *a.peakIdx = (uint32_t)bk;
            }
#else
            (void)bv; (void)bk;                                             /* the peak reduction uses warp shuffles: device only */
#endif
        }
    }
    static FFT_HD void last_out(Regs &r, const Args &a, int i)
    {
        Eng::template compute<1, INV>(r, a.tw, i);
        epilogue(r, a, i);
    }
    static FFT_HD void last_out_pk(Regs &r, const Args &a, int i, const telem *) { last_out(r, a, i); }

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, xelem *sm, int i)
    {
        if constexpr (PH < kCfftPhases) {
            C::template phase<PH, true>(r, a, sm, i);
            if constexpr (PH == kCfftPhases - 1) epilogue(r, a, i);
        } else if constexpr (PH == kCfftPhases) {
            if ((i & 31) == 0) sm[i >> 5] = r.v[0];
        } else {
#if defined(__CUDA_ARCH__)
            if (i == 0) {
                float bv = sm[0].x;
                int bk = __float_as_int(sm[0].y);
#pragma unroll
                for (int w = 1; w < T / 32; w++)
                    if (better(sm[w].x, __float_as_int(sm[w].y), bv, bk)) { bv = sm[w].x; bk = __float_as_int(sm[w].y); }
                *a.peakVal = __fsqrt_rn(bv);
                *a.peakIdx = (uint32_t)bk;
            }
#endif
        }
    }
};

/* ------------------------------------------------------------------ fixed-point real FFT, forward
 *
 * arm_rfft_q31 / arm_rfft_q15 with ifftFlagR = 0 (arm_rfft_q31.c:169-178, arm_rfft_q15.c:171-180): the
 * N-point CFFT of the frame (N = fftLenReal / 2) and arm_split_rfft_* fused: the last pass leaves X in
 * the frame's exchange buffer in natural order, then every thread splits E bins, reading X[k] and X[N-k]
 * from shared memory and writing bin k and its conjugate mirror 2N-k (the reference writes the mirror
 * explicitly, :327-329).  HBM: N elements in, 2N elements out, once.  The inverse direction is
 * CfftBody<.., RIFFT = true>. */
/* PERM: bitReverseFlagR = 0 -- the complex transform's result is laid out in the unordered (bit-reversed) order
 * before the split stage reads it as if it were natural order (arm_rfft_q31.c:173-176 with the flag passed on) */
template <class PL, bool PERM = false> struct RfftFixFwdBody {
    typedef CfftBody<PL, false> C;
    typedef typename C::Eng Eng;
    typedef typename C::A A;
    typedef typename A::elem elem;
    typedef typename A::xelem xelem;
    typedef typename A::telem telem;
    typedef typename A::work work;
    typedef typename Eng::Regs Regs;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;
    static constexpr int kCfftPhases = PhaseCount<NP>::value;
    static constexpr int kPhases = (NP == 1) ? 1 : kCfftPhases + 2;
    static_assert(NP > 1 || T == 1, "single-pass plans are one thread per frame");

    struct Args {
        const elem *in;          /* N complex = fftLenReal scalars */
        elem *out;               /* 2N complex */
        const telem *tw;         /* pass-ordered twiddles of the N-point CFFT plan */
        const ci32x4 *coef;      /* split-stage coefficients of bins 0..N-1 */
        int shl1;
        const uint16_t *perm;    /* PERM: position of X[k] in the unordered layout */
    };
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        a.out += frame * (uint64_t)(2 * N);
        return a;
    }
    static FFT_HD void set_scratch(Args &, xelem *) {}
    static FFT_HD typename C::Args cfft_args(const Args &a) { return typename C::Args{a.in, nullptr, a.tw, nullptr, 0.0f, a.shl1, nullptr}; }

    static FFT_HD void put_bin0(const Args &a, work x0)
    {
        st_stream(a.out, A::store(A::split_dc(x0)));
        st_stream(a.out + N, A::store(A::split_nyquist(x0)));
    }
    static FFT_HD void put_bin_c(const Args &a, int k, work s1, work s2, ci32x4 c)
    {
        const work o = A::split_fwd(s1, s2, c);
        st_stream(a.out + k, A::store(o));
        st_stream(a.out + (2 * N - k), A::store(A::mirror(o)));
    }
    static FFT_HD void put_bin(const Args &a, int k, work s1, work s2) { put_bin_c(a, k, s1, s2, a.coef[k]); }
#if defined(FFT_RFIX_NO_PREFETCH)
    static constexpr bool kPrefetch = false;
#else
    static constexpr bool kPrefetch = (NP > 1) && (E % 2 == 0) && sizeof(work) == 8;
#endif
    template <int K> static FFT_HD void bins_in_regs(const Args &a, const work *y)
    {
        if constexpr (K < N) {
            put_bin(a, K, y[K], y[N - K]);
            bins_in_regs<K + 1>(a, y);
        }
    }

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, xelem *sm, int i)
    {
        typedef typename PassOf<PL, NP - 1>::type PSL;
        if constexpr (NP == 1) {
            C::template phase<0, true>(r, cfft_args(a), sm, i);
            work y[N];                                                /* y[k] = X[k] */
#pragma unroll
            for (int e = 0; e < N; e++) y[PERM ? bitrev_const(PSL::out_index(e), N) : PSL::out_index(e)] = PL::kOddLog2 ? A::shl1(r.v[e]) : r.v[e];
            put_bin0(a, y[0]);
            bins_in_regs<1>(a, y);
        } else if constexpr (PH < kCfftPhases) {
            C::template phase<PH, true>(r, cfft_args(a), sm, i);
        } else if constexpr (PH == kCfftPhases) {
#pragma unroll
            for (int b = 0; b < E / PSL::R; b++)
#pragma unroll
                for (int e = 0; e < PSL::R; e++) {
                    const int k = Eng::template out_index<NP - 1>(i, b, e);
                    const work w = r.v[b * PSL::R + e];
                    /* natural order, NOT padded: lanes store / load runs of consecutive bins in both directions */
                    int pos = k;
                    if constexpr (PERM) pos = (int)a.perm[k];
                    FFT_TRACE_SMEM(&sm[pos], (int)sizeof(xelem), 1);
                    sm[pos] = A::xstore(PL::kOddLog2 ? A::shl1(w) : w);             /* cfft_q31.c:803-820, cfft_q15.c:810-827 */
                }
            /* The data registers are free now: the split coefficients of this thread's first E/2 bins are fetched into
             * them so that the loads fly across the barrier; the other half is fetched in one batch at the top of the
             * next phase.  Loaded at the point of use (one 16-byte load per bin, each waiting for the previous bin's
             * stores to issue) they formed a chain of E exposed L1/L2 latencies -- what the fused f64 kernel showed too
             * (profiles/r1_f_notes.md). */
            if constexpr (kPrefetch) {
#pragma unroll
                for (int m = 0; m < E / 2; m++) {
                    const ci32x4 c = a.coef[i + T * m];
                    r.v[2 * m] = work{c.a0, c.a1};
                    r.v[2 * m + 1] = work{c.b0, c.b1};
                }
            }
        } else {
            ci32x4 late[kPrefetch ? E / 2 : 1];
            if constexpr (kPrefetch) {
#pragma unroll
                for (int m = E / 2; m < E; m++) late[m - E / 2] = a.coef[i + T * m];
            }
#pragma unroll
            for (int m = 0; m < E; m++) {
                const int k = i + T * m;
                FFT_TRACE_SMEM(&sm[k], (int)sizeof(xelem), 0);
                FFT_TRACE_SMEM(&sm[k ? N - k : 0], (int)sizeof(xelem), 0);
                if (m == 0 && i == 0) {
                    put_bin0(a, A::xload(sm[0]));
                } else if constexpr (kPrefetch) {
                    const ci32x4 c = (m < E / 2) ? ci32x4{r.v[2 * m].x, r.v[2 * m].y, r.v[2 * m + 1].x, r.v[2 * m + 1].y} : late[m < E / 2 ? 0 : m - E / 2];
                    put_bin_c(a, k, A::xload(sm[k]), A::xload(sm[N - k]), c);
                } else {
                    put_bin(a, k, A::xload(sm[k]), A::xload(sm[N - k]));
                }
            }
        }
    }
};

/* ------------------------------------------------------------------ arm_rfft_fast_f64, fused
 *
 * Forward (arm_rfft_fast_f64.c:224-231): the L-point CFFT of the frame (L = fftLenRFFT / 2 = PL::N) and
 * stage_rfft_f64 (:30-118) in one kernel, the way RfftFixFwdBody does it for the fixed-point types: the last pass
 * leaves X in the exchange buffer in natural order, then every thread splits E bins from X[k] and X[L-k] and writes
 * the packed spectrum {DC, Nyquist, Re1, Im1, ...}.  Inverse (:215-222): merge_rfft_f64 (:121-181) is the load of the
 * inverse CFFT.  HBM: 16 L bytes in, 16 L bytes out, once.  Products are rounded on their own (ArithF64::mul) and sums
 * run left to right as in the reference, so the results are bit-identical to the two-kernel adapter and the oracle. */
FFT_HD cf64 rfft64_split(cf64 a, cf64 b, cf64 tw)          /* a = X[k], b = X[L-k], tw = (twR, twI) = coef[2k], coef[2k+1] */
{
    const double t1a = b.x - a.x, t1b = b.y + a.y;
    const double p0 = ArithF64::mul(tw.x, t1a), p1 = ArithF64::mul(tw.y, t1a), p2 = ArithF64::mul(tw.x, t1b), p3 = ArithF64::mul(tw.y, t1b);
    return {ArithF64::mul(0.5, ((a.x + b.x) + p0) + p3), ArithF64::mul(0.5, ((a.y - b.y) + p1) - p2)};
}
FFT_HD cf64 rfft64_merge(cf64 a, cf64 b, cf64 tw)
{
    const double t1a = a.x - b.x, t1b = a.y + b.y;
    const double r = ArithF64::mul(tw.x, t1a), s = ArithF64::mul(tw.y, t1b), t = ArithF64::mul(tw.y, t1a), u = ArithF64::mul(tw.x, t1b);
    return {ArithF64::mul(0.5, ((a.x + b.x) - r) - s), ArithF64::mul(0.5, ((a.y - b.y) + t) - u)};
}
FFT_HD cf64 rfft64_split0(cf64 a)                          /* :47-65 */
{
    const double t1a = a.x + a.x, t1b = a.y + a.y;
    return {ArithF64::mul(0.5, t1a + t1b), ArithF64::mul(0.5, t1a - t1b)};
}
FFT_HD cf64 rfft64_merge0(cf64 a) { return {ArithF64::mul(0.5, a.x + a.y), ArithF64::mul(0.5, a.x - a.y)}; }   /* :138-144 */

template <class PL> struct RfftF64FwdBody {
    typedef CfftBody<PL, false> C;
    typedef typename C::Eng Eng;
    typedef typename C::A A;
    typedef cf64 elem;
    typedef cf64 xelem;
    typedef cf64 telem;
    typedef cf64 work;
    typedef typename Eng::Regs Regs;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;
    static constexpr int kCfftPhases = PhaseCount<NP>::value;
    static constexpr int kPhases = kCfftPhases + 2;
    static_assert(NP > 1, "the f64 plans have at least one exchange");

    struct Args {
        const cf64 *in;          /* L complex = fftLenRFFT doubles */
        cf64 *out;               /* L complex: packed spectrum */
        const cf64 *tw;          /* twiddles of the L-point CFFT plan */
        const cf64 *twr;         /* twiddleCoefF64_rfft: (sin, cos) pairs, L entries */
    };
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        a.out += frame * (uint64_t)N;
        return a;
    }
    static FFT_HD void set_scratch(Args &, xelem *) {}
    static FFT_HD typename C::Args cfft_args(const Args &a) { return typename C::Args{a.in, nullptr, a.tw, nullptr, 0.0f, 0, nullptr}; }

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, xelem *sm, int i)
    {
        typedef typename PassOf<PL, NP - 1>::type PSL;
        if constexpr (PH < kCfftPhases) {
            C::template phase<PH, true>(r, cfft_args(a), sm, i);
        } else if constexpr (PH == kCfftPhases) {
#pragma unroll
            for (int b = 0; b < E / PSL::R; b++)
#pragma unroll
                for (int e = 0; e < PSL::R; e++) {
                    const int k = Eng::template out_index<NP - 1>(i, b, e);
                    FFT_TRACE_SMEM(&sm[k], (int)sizeof(xelem), 1);
                    sm[k] = r.v[b * PSL::R + e];                      /* natural order, not padded */
                }
            /* the registers are free now: fetch this thread's real-stage twiddles so that the loads fly across the
             * barrier (loaded at the point of use they serialised on one register pair: 16 L2 latencies per thread,
             * the largest stall of the kernel, profiles/r1_f_ncu_hotspots_rfft_f64_fwd_4096.txt) */
#pragma unroll
            for (int m = 0; m < E; m++) r.v[m] = a.twr[i + T * m];
        } else {
#pragma unroll
            for (int m = 0; m < E; m++) {
                const int k = i + T * m;
                FFT_TRACE_SMEM(&sm[k], (int)sizeof(xelem), 0);
                FFT_TRACE_SMEM(&sm[k ? N - k : 0], (int)sizeof(xelem), 0);
                if (m == 0 && i == 0) st_stream(a.out, rfft64_split0(sm[0]));
                else st_stream(a.out + k, rfft64_split(sm[k], sm[N - k], r.v[m]));
            }
        }
    }
};

template <class PL> struct RfftF64InvBody {
    typedef CfftBody<PL, true> C;
    typedef typename C::Eng Eng;
    typedef typename C::A A;
    typedef cf64 elem;
    typedef cf64 xelem;
    typedef cf64 telem;
    typedef cf64 work;
    typedef typename Eng::Regs Regs;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;
    static constexpr int kPhases = PhaseCount<NP>::value;
    static_assert(NP > 1, "the f64 plans have at least one exchange");

    struct Args {
        const cf64 *in;          /* packed spectrum, L complex */
        cf64 *out;               /* L complex = fftLenRFFT doubles */
        const cf64 *tw;
        const cf64 *twr;
        float scale;             /* 1 / L */
    };
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        a.out += frame * (uint64_t)N;
        return a;
    }
    static FFT_HD void set_scratch(Args &, xelem *) {}
    static FFT_HD typename C::Args cfft_args(const Args &a) { return typename C::Args{a.in, a.out, a.tw, nullptr, a.scale, 0, nullptr}; }

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, xelem *sm, int i)
    {
        if constexpr (PH == 0) {
            typedef typename PL::P0 PS;
#pragma unroll
            for (int b = 0; b < E / PS::R; b++)
#pragma unroll
                for (int e = 0; e < PS::R; e++) {
                    const int idx = Eng::template in_index<0>(i, b, e);
                    /* every bin is read twice (as X[k] and as X[L-k]): plain cached loads */
                    work w = idx ? rfft64_merge(a.in[idx], a.in[N - idx], a.twr[idx]) : rfft64_merge0(a.in[0]);
                    w.y = -w.y;                                           /* conjugate input (arm_cfft_f64.c:262-272) */
                    r.v[b * PS::R + e] = w;
                }
            Eng::template compute<0, true>(r, a.tw, i);
            Eng::template smem_store<0>(r, sm, i);
        } else {
            C::template phase<PH>(r, cfft_args(a), sm, i);
        }
    }
};

/* ------------------------------------------------------------------ RFFT (f32 only) */

/* split stage for one bin: A = X[k], B = X[Nh-k], tw = twiddleCoef_rfft[k] = (sin,cos)
 * (arm_rfft_fast_f32.c:372-395):
 *   out = 0.5 * ( A + conj(B) + d.x * (tw.x, tw.y) + d.y * (tw.y, -tw.x) ),  d = (B.x - A.x, B.y + A.y) */
FFT_HD cf32 rfft_split(cf32 A, cf32 B, cf32 tw)
{
    const cf32 s = cadd(A, cf32{B.x, -B.y});
    const cf32 d = cadd(B, cf32{-A.x, A.y});
    const cf32 t = caxpy(mul_mi(tw), d.y, caxpy(tw, d.x, s));
    return cscale(t, 0.5f);
}
/* merge stage for one bin (arm_rfft_fast_f32.c:436-455):
 *   out = 0.5 * ( A + conj(B) - d.x * (tw.x, -tw.y) - d.y * (tw.y, tw.x) ),  d = (A.x - B.x, A.y + B.y) */
FFT_HD cf32 rfft_merge(cf32 A, cf32 B, cf32 tw)
{
    const cf32 s = cadd(A, cf32{B.x, -B.y});
    const cf32 d = cadd(A, cf32{-B.x, B.y});
    const cf32 t = caxpy(cf32{tw.y, tw.x}, -d.y, caxpy(cf32{tw.x, -tw.y}, -d.x, s));
    return cscale(t, 0.5f);
}
/* twiddleCoef_rfft entry of bin k + e*NBF from the entry of bin k: the angle grows by
 * e*pi/R = 2*pi*e*(64/R)/128, a compile-time rotation of the (sin, cos) pair */
template <int R, int E_> FFT_HD cf32 rfft_tw_rot(cf32 tw)
{
    if (E_ == 0) return tw;
    constexpr float c = cos128(E_ * (64 / R)), sn = sin128(E_ * (64 / R));
    return caxpy(mul_mi(tw), sn, cscale(tw, c));       /* (s c + c' sn , c' c - s sn) with tw = (s, c') */
}
/* entry of bin Nh - k from the entry of bin k: (sin, cos)(pi - x) = (sin x, -cos x) */
FFT_HD cf32 rfft_tw_mirror(cf32 tw) { return {tw.x, -tw.y}; }

/* Requires: the pass next to the real side is a Mirror pass, so thread i holds its butterflies
 * in pairs (p, NBF - p): slot e of the first pairs with slot R-1-e of the second, i.e. every
 * (k, Nh-k) pair is thread-local and shares one rfft twiddle (up to the mirror sign).  One table
 * entry is LOADED per butterfly pair; the entries of the other bins are compile-time rotations.
 *
 * Pair 0 of thread 0 is the exception: it holds butterflies 0 and NBF/2, i.e. the 2R bins
 * h*NBF/2, which pair among themselves (h <-> 2R-h; h = 0 is the packed DC/Nyquist bin, h = R
 * pairs with itself).  Treating them inside thread 0 would make every warp that contains a
 * thread 0 run a second, divergent copy of the whole epilogue, so these bins take a detour
 * through a 2R-element scratch area behind the frame's exchange buffer and are split / merged
 * by all T threads of the frame, a few bins each (phase 2 forward, phase 0 inverse). */
struct RfftFwdArgs {
    const cf32 *in;      /* real frame viewed as N complex */
    cf32 *out;           /* packed spectrum: N complex = 2N floats */
    const cf32 *tw;      /* pass-ordered CFFT twiddles of this plan */
    const cf32 *twr;     /* twiddleCoef_rfft_(2N): (sin,cos)(2*pi*k/(2N)), k < N */
    cf32 *scratch;       /* 2R elements of shared memory per frame (set by the kernel) */
    const cf32 *win;     /* WIN only: the 2N window values, viewed as N pairs */
};
struct RfftInvArgs {
    const cf32 *in;      /* packed spectrum, N complex */
    cf32 *out;           /* real frame viewed as N complex */
    const cf32 *tw;
    const cf32 *twr;
    float scale;         /* 1/N */
    cf32 *scratch;       /* 2R elements of shared memory per frame (set by the kernel) */
};

/* WIN: the real samples are multiplied by a window first (arm_mult_f32 as in arm_mfcc_f32.c:112, fused into the load) */
template <class PL, bool STAGED = false, bool WIN = false> struct RfftFwdBody {
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef cf32 elem;
    typedef cf32 xelem;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;   /* N = complex length = real length / 2 */
    static_assert(NP == 2, "rfft plans are two-pass plans");
    static constexpr int kPhases = 3;
    typedef typename PassOf<PL, NP - 1>::type PL_LAST;
    static constexpr int R = PL_LAST::R, NBF = N / R, NB = E / R;
    static_assert(PL_LAST::kMirror && NB % 2 == 0, "forward rfft needs a trailing Mirror pass with an even number of butterflies per thread");
    static_assert(PL::kSpecial == 2 * R, "plan must reserve the scratch area of the special bins");

    typedef RfftFwdArgs Args;
    static FFT_HD void set_scratch(Args &a, cf32 *p) { a.scratch = p; }
    /* frame-local view of the batch arguments; with STAGED the kernel overrides `in` */
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        a.out += frame * (uint64_t)N;
        return a;
    }
    static constexpr bool kStaged = STAGED;

    static FFT_HD void gload(Regs &r, const Args &a, int i)
    {
        typedef typename PL::P0 PS;
#pragma unroll
        for (int b = 0; b < E / PS::R; b++)
#pragma unroll
            for (int e = 0; e < PS::R; e++) {
                const int idx = Eng::template in_index<0>(i, b, e);
                cf32 v = ld_in<STAGED>(a.in + idx);
                if constexpr (WIN) {
                    const cf32 wv = a.win[idx];
                    v = cf32{v.x * wv.x, v.y * wv.y};
                }
                r.v[b * PS::R + e] = v;
            }
    }
    /* where a finished spectrum bin goes: by default into the packed spectrum at a.out (a fused
     * consumer such as the MFCC kernel passes its own sink and never materialises the spectrum) */
    struct SpectrumSink {
        cf32 *out;
        FFT_HD void put(int k, cf32 v, bool pred) const { st_out_if<STAGED>(pred, out + k, v); }
        FFT_HD void put_dc_nyquist(cf32 v) const { st_out<STAGED>(out, v); }      /* bin 0 = (DC, Nyquist) */
    };
    template <int E0, class SINK> static FFT_HD void pair_loop(const cf32 *A, const cf32 *B, int p, cf32 tw0, bool regular, const SINK &sink)
    {
        /* bins k0 = p + e*NBF (in A) and Nh - k0 = (NBF - p) + (R-1-e)*NBF (in B) */
        if constexpr (E0 < R) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const int k0 = p + E0 * NBF;
            const cf32 o0 = rfft_split(A[E0], B[R - 1 - E0], tw), o1 = rfft_split(B[R - 1 - E0], A[E0], rfft_tw_mirror(tw));
            sink.put(k0, o0, regular);
            sink.put(N - k0, o1, regular);
            pair_loop<E0 + 1>(A, B, p, tw0, regular, sink);
        }
    }
    static FFT_HD void split_store(const Regs &r, const Args &a, cf32 *scratch, int i)
    {
        cf32 ptw[NB / 2];
#pragma unroll
        for (int m = 0; m < NB / 2; m++) ptw[m] = a.twr[i + T * m];
        split_store(r, a, scratch, i, ptw);
    }
    static FFT_HD void split_store(const Regs &r, const Args &a, cf32 *scratch, int i, const cf32 *ptw)
    {
        split_store(r, scratch, i, ptw, SpectrumSink{a.out});
    }
    template <class SINK> static FFT_HD void split_store(const Regs &r, cf32 *scratch, int i, const cf32 *ptw, const SINK &sink)
    {
#pragma unroll
        for (int m = 0; m < NB / 2; m++) {
            const int p = i + T * m;
            const cf32 *A = &r.v[(2 * m) * R], *B = &r.v[(2 * m + 1) * R];
            pair_loop<0>(A, B, p, ptw[m], m != 0 || i != 0, sink);
            if (m == 0 && i == 0) {          /* X[h*NBF/2] -> scratch[h] */
#pragma unroll
                for (int e = 0; e < R; e++) {
                    scratch[2 * e] = A[e];
                    scratch[2 * e + 1] = B[e];
                }
            }
        }
    }
    /* rfft twiddles of the special bins this thread handles: h = i + T*q -> twr[h*NBF/2] (h = 0 -> twr[N/2]) */
    static constexpr int kNS = (R + T - 1) / T;
    static FFT_HD void load_special_tw(cf32 *stw, const Args &a, int i)
    {
#pragma unroll
        for (int q = 0; q < kNS; q++) {
            const int h = i + T * q;
            stw[q] = a.twr[h == 0 ? N / 2 : (h < R ? h * (NBF / 2) : 0)];
        }
    }
    static FFT_HD void special_bins(const Args &a, const cf32 *scratch, int i, const cf32 *stw)
    {
        special_bins(scratch, i, stw, SpectrumSink{a.out});
    }
    template <class SINK> static FFT_HD void special_bins(const cf32 *scratch, int i, const cf32 *stw, const SINK &sink)
    {
#pragma unroll
        for (int q = 0; q < kNS; q++) {
            const int h = i + T * q;
            if (h >= R) break;
            if (h == 0) {
                const cf32 X0 = scratch[0], XR = scratch[R];
                sink.put_dc_nyquist(cf32{X0.x + X0.y, X0.x - X0.y});             /* rfft_fast_f32.c:337-352 */
                sink.put(N / 2, rfft_split(XR, XR, stw[q]), true);
            } else {
                const int k = h * (NBF / 2);
                const cf32 A = scratch[h], B = scratch[2 * R - h], tw = stw[q];
                sink.put(k, rfft_split(A, B, tw), true);
                sink.put(N - k, rfft_split(B, A, rfft_tw_mirror(tw)), true);
            }
        }
    }
    static constexpr bool kHasPre = false, kHasPost = true;
    /* pipelined kernel: the packed spectrum is assembled in shared memory (regular bins and the 2R
     * special bins alike) and leaves with one bulk store -- full-line writes instead of 8-byte
     * pieces with holes at the special bins */
    static constexpr bool kImageOut = STAGED;
    static FFT_HD void pre(const Args &, cf32 *, int) {}
    static FFT_HD void phase0_in(Regs &r, const Args &a, cf32 *, int i)
    {
        gload(r, a, i);
        Eng::template compute<0, false>(r, a.tw, i);
    }
    static FFT_HD void phase0_out(const Regs &r, cf32 *sm, int i) { Eng::template smem_store<0>(r, sm, i); }
    static FFT_HD void last_in(Regs &r, const cf32 *sm, int i) { Eng::template smem_load<1>(r, sm, i); }
    static FFT_HD void last_out(Regs &r, const Args &a, int i)
    {
        Eng::template compute<1, false>(r, a.tw, i);
        split_store(r, a, a.scratch, i);
    }
    static FFT_HD void last(Regs &r, const Args &a, cf32 *sm, int i)
    {
        last_in(r, sm, i);
        last_out(r, a, i);
    }
    static FFT_HD void post(const Args &a, cf32 *, int i)
    {
        cf32 stw[kNS];
        load_special_tw(stw, a, i);
        special_bins(a, a.scratch, i, stw);
    }
    struct Hoist { typename Eng::template TwRegs<0> t0; cf32 stw[kNS]; cf32 ptw[NB / 2]; };
    static constexpr int kH0 = (int)(sizeof(typename Eng::template TwRegs<0>) / sizeof(cf32));
    static FFT_HD void hoist(Hoist &h, const Args &a, int i)
    {
        Eng::template load_tw<0>(h.t0, a.tw, i);
        load_special_tw(h.stw, a, i);
#pragma unroll
        for (int m = 0; m < NB / 2; m++) h.ptw[m] = a.twr[i + T * m];
    }
    static FFT_HD void pre_pk(const Args &, cf32 *, int, const cf32 *) {}
    static FFT_HD void phase0_in_pk(Regs &r, const Args &a, cf32 *, int i, const cf32 *pk)
    {
        typename Eng::template TwRegs<0> t;
#pragma unroll
        for (int s = 0; s < kH0; s++) t.w[s] = pk[s * PL::kThreads];
        gload(r, a, i);
        Eng::template compute_pre<0, false>(r, t);
    }
    static FFT_HD void last_out_pk(Regs &r, const Args &a, int i, const cf32 *pk)
    {
        cf32 ptw[NB / 2];
#pragma unroll
        for (int m = 0; m < NB / 2; m++) ptw[m] = pk[(kH0 + kNS + m) * PL::kThreads];
        Eng::template compute<1, false>(r, a.tw, i);
        split_store(r, a, a.scratch, i, ptw);
    }
    static FFT_HD void post_pk(const Args &a, cf32 *, int i, const cf32 *pk)
    {
        cf32 stw[kNS];
#pragma unroll
        for (int q = 0; q < kNS; q++) stw[q] = pk[(kH0 + q) * PL::kThreads];
        special_bins(a, a.scratch, i, stw);
    }

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, cf32 *sm, int i)
    {
        if constexpr (PH == 0) {
            phase0_in(r, a, sm, i);
            phase0_out(r, sm, i);
        } else if constexpr (PH == 1) {
            last(r, a, sm, i);
        } else {
            post(a, sm, i);
        }
    }
};

template <class PL, bool STAGED = false> struct RfftInvBody {
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef cf32 elem;
    typedef cf32 xelem;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;
    static_assert(NP == 2, "rfft plans are two-pass plans");
    static constexpr int kPhases = 3;
    static constexpr int R = PL::P0::R, NBF = N / R, NB = E / R;
    static_assert(PL::P0::kMirror && NB % 2 == 0, "inverse rfft needs a leading Mirror pass with an even number of butterflies per thread");
    static_assert(PL::kSpecial == 2 * R, "plan must reserve the scratch area of the special bins");

    typedef RfftInvArgs Args;
    static FFT_HD void set_scratch(Args &a, cf32 *p) { a.scratch = p; }
    /* frame-local view of the batch arguments; with STAGED the kernel overrides `in` */
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        a.out += frame * (uint64_t)N;
        return a;
    }
    static constexpr bool kStaged = STAGED;

    /* merge (rfft_fast_f32.c:405-462) then conjugate for the inverse CFFT (cfft_f32.c:1252-1261) */
    static FFT_HD cf32 mconj(cf32 z) { return {z.x, -z.y}; }

    /* the 2R bins h*NBF/2 of butterflies 0 and NBF/2, merged by all threads of the frame */
    static constexpr int kNS = (R + T - 1) / T;
    static FFT_HD void load_special_tw(cf32 *stw, const Args &a, int i)
    {
#pragma unroll
        for (int q = 0; q < kNS; q++) {
            const int h = i + T * q;
            stw[q] = a.twr[h == 0 ? N / 2 : (h < R ? h * (NBF / 2) : 0)];
        }
    }
    static FFT_HD void special_merge(const Args &a, cf32 *scratch, int i, const cf32 *stw)
    {
#pragma unroll
        for (int q = 0; q < kNS; q++) {
            const int h = i + T * q;
            if (h >= R) break;
            if (h == 0) {
                const cf32 g0 = ld_in<STAGED>(a.in), gR = ld_in<STAGED>(a.in + N / 2);
                scratch[0] = mconj(cf32{0.5f * (g0.x + g0.y), 0.5f * (g0.x - g0.y)});   /* :425-431 */
                scratch[R] = mconj(rfft_merge(gR, gR, stw[q]));
            } else {
                const int k = h * (NBF / 2);
                const cf32 ga = ld_in<STAGED>(a.in + k), gb = ld_in<STAGED>(a.in + (N - k)), tw = stw[q];
                scratch[h] = mconj(rfft_merge(ga, gb, tw));
                scratch[2 * R - h] = mconj(rfft_merge(gb, ga, rfft_tw_mirror(tw)));
            }
        }
    }
    template <int E0> static FFT_HD void pair_loop(cf32 *A, cf32 *B, cf32 tw0)
    {
        if constexpr (E0 < R) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const cf32 ga = A[E0], gb = B[R - 1 - E0];
            A[E0] = mconj(rfft_merge(ga, gb, tw));
            B[R - 1 - E0] = mconj(rfft_merge(gb, ga, rfft_tw_mirror(tw)));
            pair_loop<E0 + 1>(A, B, tw0);
        }
    }
    static FFT_HD void merge_load(Regs &r, const Args &a, const cf32 *scratch, int i)
    {
        cf32 ptw[NB / 2];
#pragma unroll
        for (int m = 0; m < NB / 2; m++) ptw[m] = a.twr[i + T * m];
        merge_load(r, a, scratch, i, ptw);
    }
    static FFT_HD void merge_load(Regs &r, const Args &a, const cf32 *scratch, int i, const cf32 *ptw)
    {
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int e = 0; e < R; e++) r.v[b * R + e] = ld_in<STAGED>(a.in + Eng::template in_index<0>(i, b, e));
#pragma unroll
        for (int m = 0; m < NB / 2; m++) {
            const int p = i + T * m;
            cf32 *A = &r.v[(2 * m) * R], *B = &r.v[(2 * m + 1) * R];
            (void)p;
            pair_loop<0>(A, B, ptw[m]);
            if (m == 0 && i == 0) {          /* butterflies 0 and NBF/2 come merged from the scratch area */
#pragma unroll
                for (int e = 0; e < R; e++) {
                    A[e] = scratch[2 * e];
                    B[e] = scratch[2 * e + 1];
                }
            }
        }
    }
    static FFT_HD void gstore(const Regs &r, const Args &a, int i)
    {
        typedef typename PassOf<PL, NP - 1>::type PS;
#pragma unroll
        for (int b = 0; b < E / PS::R; b++)
#pragma unroll
            for (int e = 0; e < PS::R; e++) {
                const int k = Eng::template out_index<NP - 1>(i, b, e);
                st_stream(a.out + k, cscale(mconj(r.v[b * PS::R + e]), a.scale));
            }
    }
    static constexpr bool kHasPre = true, kHasPost = false;
    static constexpr bool kImageOut = false;
    static FFT_HD void pre(const Args &a, cf32 *, int i)
    {
        cf32 stw[kNS];
        load_special_tw(stw, a, i);
        special_merge(a, a.scratch, i, stw);
    }
    static FFT_HD void post(const Args &, cf32 *, int) {}
    static FFT_HD void phase0_in(Regs &r, const Args &a, cf32 *, int i)
    {
        merge_load(r, a, a.scratch, i);
        Eng::template compute<0, false>(r, a.tw, i);
    }
    struct Hoist { typename Eng::template TwRegs<0> t0; cf32 stw[kNS]; cf32 ptw[NB / 2]; };
    static constexpr int kH0 = (int)(sizeof(typename Eng::template TwRegs<0>) / sizeof(cf32));
    static FFT_HD void hoist(Hoist &h, const Args &a, int i)
    {
        Eng::template load_tw<0>(h.t0, a.tw, i);
        load_special_tw(h.stw, a, i);
#pragma unroll
        for (int m = 0; m < NB / 2; m++) h.ptw[m] = a.twr[i + T * m];
    }
    static FFT_HD void pre_pk(const Args &a, cf32 *, int i, const cf32 *pk)
    {
        cf32 stw[kNS];
#pragma unroll
        for (int q = 0; q < kNS; q++) stw[q] = pk[(kH0 + q) * PL::kThreads];
        special_merge(a, a.scratch, i, stw);
    }
    static FFT_HD void post_pk(const Args &, cf32 *, int, const cf32 *) {}
    static FFT_HD void last_out_pk(Regs &r, const Args &a, int i, const cf32 *) { last_out(r, a, i); }
    static FFT_HD void phase0_in_pk(Regs &r, const Args &a, cf32 *, int i, const cf32 *pk)
    {
        cf32 ptw[NB / 2];
#pragma unroll
        for (int m = 0; m < NB / 2; m++) ptw[m] = pk[(kH0 + kNS + m) * PL::kThreads];
        merge_load(r, a, a.scratch, i, ptw);
        typename Eng::template TwRegs<0> t;
#pragma unroll
        for (int s = 0; s < kH0; s++) t.w[s] = pk[s * PL::kThreads];
        Eng::template compute_pre<0, false>(r, t);
    }
    static FFT_HD void phase0_out(const Regs &r, cf32 *sm, int i) { Eng::template smem_store<0>(r, sm, i); }
    static FFT_HD void last_in(Regs &r, const cf32 *sm, int i) { Eng::template smem_load<1>(r, sm, i); }
    static FFT_HD void last_out(Regs &r, const Args &a, int i)
    {
        Eng::template compute<1, false>(r, a.tw, i);
        gstore(r, a, i);
    }
    static FFT_HD void last(Regs &r, const Args &a, cf32 *sm, int i)
    {
        last_in(r, sm, i);
        last_out(r, a, i);
    }

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, cf32 *sm, int i)
    {
        if constexpr (PH == 0) {
            pre(a, sm, i);
        } else if constexpr (PH == 1) {
            phase0_in(r, a, sm, i);
            phase0_out(r, sm, i);
        } else {
            last(r, a, sm, i);
        }
    }
};

/* ------------------------------------------------------------------ one thread per frame (N <= 64)
 *
 * Single-pass plans (T = 1, E = N): a thread transforms a whole frame in registers, no exchange.
 * The frame is read and written with 16-byte vector accesses, which is what lets the kernel
 * that feeds these bodies (frame_kernel_tiny: per-frame bulk copies into a padded shared-memory
 * slot, and back) run conflict-free; in the emulator `in`/`out` are plain host frames. */

template <class ELEM> struct alignas(16) Vec16 { ELEM e[16 / sizeof(ELEM)]; };

template <class A, int N> struct TinyIO {
    typedef typename A::elem elem;
    typedef typename A::work work;
    static constexpr int V = 16 / (int)sizeof(elem);
    static FFT_HD void load(work *w, const elem *in)
    {
#pragma unroll
        for (int m = 0; m < N / V; m++) {
            const Vec16<elem> q = *reinterpret_cast<const Vec16<elem> *>(in + V * m);
#pragma unroll
            for (int u = 0; u < V; u++) w[V * m + u] = A::load(q.e[u]);
        }
    }
    /* w[k] = value of position k */
    static FFT_HD void store(const work *w, elem *out)
    {
#pragma unroll
        for (int m = 0; m < N / V; m++) {
            Vec16<elem> q;
#pragma unroll
            for (int u = 0; u < V; u++) q.e[u] = A::store(w[V * m + u]);
            *reinterpret_cast<Vec16<elem> *>(out + V * m) = q;
        }
    }
};

template <class PL, bool INV, bool PERM = false> struct TinyCfftBody {
    typedef CfftBody<PL, INV, PERM, true> Base;
    typedef typename Base::Eng Eng;
    typedef typename Base::A A;
    typedef typename A::elem elem;
    typedef typename A::xelem xelem;
    typedef typename A::work work;
    typedef typename Base::Regs Regs;
    typedef typename Base::Args Args;
    static constexpr int N = PL::N, kPhases = 1;
    static constexpr bool kF32 = Base::kF32;
    static_assert(PL::NP == 1 && PL::T == 1, "one thread per frame, one pass");
    static FFT_HD Args for_frame(Args a, uint64_t frame) { return Base::for_frame(a, frame); }
    static FFT_HD void set_scratch(Args &, xelem *) {}

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, xelem *, int)
    {
        typedef typename PL::P0 PS;
        TinyIO<A, N>::load(r.v, a.in);
        if (kF32 && INV) {
#pragma unroll
            for (int k = 0; k < N; k++) r.v[k].y = -r.v[k].y;            /* cfft_f32.c:1252-1261 */
        }
        Eng::template compute<0, INV>(r, a.tw, 0);
        work y[N];                                                        /* y[k] = X[k] */
#pragma unroll
        for (int e = 0; e < N; e++) {
            work w = r.v[e];
            if (kF32) {
                if (INV) w = Base::scale_conj(w, a.scale);                /* cfft_f32.c:1285-1297 */
            } else if (PL::kOddLog2) {
                w = A::shl1(w);
            }
            y[PS::out_index(e)] = w;
        }
        if constexpr (PERM) {
#pragma unroll
            for (int k = 0; k < N; k++) a.out[a.perm[k]] = A::store(y[k]);
        } else {
            TinyIO<A, N>::store(y, a.out);
        }
    }
};

template <class PL> struct TinyRfftFwdBody {
    typedef RfftFwdArgs Args;
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef cf32 elem;
    typedef cf32 xelem;
    static constexpr int N = PL::N, kPhases = 1;       /* complex length */
    static_assert(PL::NP == 1 && PL::T == 1, "one thread per frame, one pass");
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        a.out += frame * (uint64_t)N;
        return a;
    }
    static FFT_HD void set_scratch(Args &, cf32 *) {}
    template <int K> static FFT_HD void bins(const cf32 *X, cf32 *y, const cf32 *__restrict__ twr)
    {
        if constexpr (K < N / 2) {
            const cf32 tw = twr[K];
            y[K] = rfft_split(X[K], X[N - K], tw);
            y[N - K] = rfft_split(X[N - K], X[K], rfft_tw_mirror(tw));
            bins<K + 1>(X, y, twr);
        }
    }
    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, cf32 *, int)
    {
        TinyIO<ArithF32, N>::load(r.v, a.in);
        Eng::template compute<0, false>(r, a.tw, 0);
        cf32 y[N];
        y[0] = cf32{r.v[0].x + r.v[0].y, r.v[0].x - r.v[0].y};            /* rfft_fast_f32.c:337-352 */
        y[N / 2] = rfft_split(r.v[N / 2], r.v[N / 2], a.twr[N / 2]);
        bins<1>(r.v, y, a.twr);
        TinyIO<ArithF32, N>::store(y, a.out);
    }
};

template <class PL> struct TinyRfftInvBody {
    typedef RfftInvArgs Args;
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef cf32 elem;
    typedef cf32 xelem;
    static constexpr int N = PL::N, kPhases = 1;
    static_assert(PL::NP == 1 && PL::T == 1, "one thread per frame, one pass");
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)N;
        a.out += frame * (uint64_t)N;
        return a;
    }
    static FFT_HD void set_scratch(Args &, cf32 *) {}
    static FFT_HD cf32 mconj(cf32 z) { return {z.x, -z.y}; }
    template <int K> static FFT_HD void bins(const cf32 *G, cf32 *x, const cf32 *__restrict__ twr)
    {
        if constexpr (K < N / 2) {
            const cf32 tw = twr[K];
            x[K] = mconj(rfft_merge(G[K], G[N - K], tw));
            x[N - K] = mconj(rfft_merge(G[N - K], G[K], rfft_tw_mirror(tw)));
            bins<K + 1>(G, x, twr);
        }
    }
    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, cf32 *, int)
    {
        cf32 g[N];
        TinyIO<ArithF32, N>::load(g, a.in);
        r.v[0] = mconj(cf32{0.5f * (g[0].x + g[0].y), 0.5f * (g[0].x - g[0].y)});   /* :425-431 */
        r.v[N / 2] = mconj(rfft_merge(g[N / 2], g[N / 2], a.twr[N / 2]));
        bins<1>(g, r.v, a.twr);
        Eng::template compute<0, false>(r, a.tw, 0);
        cf32 y[N];
#pragma unroll
        for (int k = 0; k < N; k++) y[k] = cscale(mconj(r.v[k]), a.scale);
        TinyIO<ArithF32, N>::store(y, a.out);
    }
};

/* ------------------------------------------------------------------ thread per frame: the fused entry points
 *
 * Same data path as TinyCfftBody (the kernel brings the lane's frame into its padded slot with one bulk copy and
 * sends the slot home with another), for bodies whose input and output differ in size: kTinyInBytes / kTinyOutBytes,
 * tiny_home (where the slot goes) and tiny_bind (point the arguments at the slot).  Slots are read and written with
 * 16-byte vectors only (conflict-free across the lanes of a warp). */

template <class A, int COUNT> struct VecOut {
    typedef typename A::elem elem;
    typedef typename A::work work;
    static constexpr int V = 16 / (int)sizeof(elem);
    /* store z[G*V .. G*V+V-1] as one vector */
    template <int G> static FFT_HD void group(const work *z, elem *out)
    {
        Vec16<elem> q;
#pragma unroll
        for (int u = 0; u < V; u++) q.e[u] = A::store(z[G * V + u]);
        *reinterpret_cast<Vec16<elem> *>(out + G * V) = q;
    }
};

/* arm_rfft_q31 / arm_rfft_q15, complex length N <= 64 (real length <= 128) */
template <class PL, bool INV, bool PERM = false> struct TinyRfftFixBody {
    typedef typename PL::Arith A;
    typedef typename A::elem elem;
    typedef typename A::xelem xelem;
    typedef typename A::telem telem;
    typedef typename A::work work;
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef typename PL::P0 PS;
    static constexpr int N = PL::N, kPhases = 1, V = 16 / (int)sizeof(elem);
    static_assert(PL::NP == 1 && PL::T == 1, "one thread per frame, one pass");
    struct Args {
        const elem *in;
        elem *out;
        const telem *tw;
        const ci32x4 *coef;
        int shl1;
    };
    static FFT_HD Args for_frame(Args a, uint64_t frame)
    {
        a.in += frame * (uint64_t)(INV ? 2 * N : N);
        a.out += frame * (uint64_t)(INV ? N : 2 * N);
        return a;
    }
    static FFT_HD void set_scratch(Args &, xelem *) {}
    /* inverse: bins 0..N of the 2N-bin spectrum frame, rounded up to whole vectors (the frame is longer than that) */
    static constexpr int kInElems = INV ? ((N + 1 + V - 1) / V) * V : N;
    static constexpr int kTinyInBytes = kInElems * (int)sizeof(elem), kTinyOutBytes = (INV ? N : 2 * N) * (int)sizeof(elem);
    static FFT_HD void *tiny_home(const Args &a) { return a.out; }
    static FFT_HD void tiny_bind(Args &a, void *slot) { a.in = reinterpret_cast<const elem *>(slot); a.out = reinterpret_cast<elem *>(slot); }

    /* forward: bins k = K and 2N-K; a vector leaves as soon as its last element exists */
    template <int K> static FFT_HD void fwd_bins(const Args &a, const work *y, work *z)
    {
        if constexpr (K < N) {
            const work o = A::split_fwd(y[K], y[N - K], a.coef[K]);
            z[K] = o;
            z[2 * N - K] = A::mirror(o);
            if constexpr (K % V == V - 1) VecOut<A, 2 * N>::template group<K / V>(z, a.out);
            if constexpr ((2 * N - K) % V == 0) VecOut<A, 2 * N>::template group<(2 * N - K) / V>(z, a.out);
            fwd_bins<K + 1>(a, y, z);
        }
    }
    template <int G> static FFT_HD void out_groups(const work *y, elem *out)
    {
        if constexpr (G < N / V) {
            VecOut<A, N>::template group<G>(y, out);
            out_groups<G + 1>(y, out);
        }
    }
    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, xelem *, int)
    {
        if constexpr (!INV) {
            TinyIO<A, N>::load(r.v, a.in);
            Eng::template compute<0, false>(r, a.tw, 0);
            work y[N];                                                 /* y[k] = X[k] */
#pragma unroll
            for (int e = 0; e < N; e++) y[PERM ? bitrev_const(PS::out_index(e), N) : PS::out_index(e)] = PL::kOddLog2 ? A::shl1(r.v[e]) : r.v[e];
            work z[2 * N];
            z[0] = A::split_dc(y[0]);
            z[N] = A::split_nyquist(y[0]);
            fwd_bins<1>(a, y, z);
            VecOut<A, 2 * N>::template group<N / V>(z, a.out);        /* the vector that starts with bin N (Nyquist) */
        } else {
            work g[kInElems];
            TinyIO<A, kInElems>::load(g, a.in);
#pragma unroll
            for (int b = 0; b < N / PS::R; b++)
#pragma unroll
                for (int e = 0; e < PS::R; e++) {
                    const int idx = Eng::template in_index<0>(0, b, e);
                    r.v[b * PS::R + e] = A::split_inv(g[idx], g[N - idx], a.coef[idx]);
                }
            Eng::template compute<0, true>(r, a.tw, 0);
            work y[N];
#pragma unroll
            for (int e = 0; e < N; e++) {
                work w = r.v[e];
                if (PL::kOddLog2) w = A::shl1(w);
                y[PERM ? bitrev_const(PS::out_index(e), N) : PS::out_index(e)] = A::sat_shl1(w);
            }
            out_groups<0>(y, a.out);
        }
    }
};

/* arm_cfft_f32 + magnitude / peak, N <= 64 */
template <class PL, bool INV, int MODE> struct TinyCfftMagBody {
    typedef CfftMagBody<PL, INV, MODE, true> M;
    typedef typename M::Args Args;
    typedef typename M::Eng Eng;
    typedef typename M::Regs Regs;
    typedef cf32 elem;
    typedef cf32 xelem;
    typedef typename PL::P0 PS;
    static constexpr int N = PL::N, kPhases = 1;
    static_assert(PL::NP == 1 && PL::T == 1, "one thread per frame, one pass");
    static FFT_HD Args for_frame(Args a, uint64_t frame) { return M::for_frame(a, frame); }
    static FFT_HD void set_scratch(Args &, xelem *) {}
    static constexpr int kTinyInBytes = N * (int)sizeof(cf32), kTinyOutBytes = (MODE == SPEC_PEAK) ? 0 : N * (int)sizeof(float);
    static FFT_HD void *tiny_home(const Args &a) { return a.mag; }
    static FFT_HD void tiny_bind(Args &a, void *slot)
    {
        a.in = reinterpret_cast<const cf32 *>(slot);
        if (MODE != SPEC_PEAK) a.mag = reinterpret_cast<float *>(slot);
    }
    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, xelem *, int)
    {
        TinyIO<ArithF32, N>::load(r.v, a.in);
        if (INV) {
#pragma unroll
            for (int k = 0; k < N; k++) r.v[k].y = -r.v[k].y;                /* cfft_f32.c:1252-1261 */
        }
        Eng::template compute<0, INV>(r, a.tw, 0);
        float m[N];                                                           /* m[k] = |X[k]| (squared for SPEC_MAG_SQUARED / SPEC_PEAK) */
#pragma unroll
        for (int e = 0; e < N; e++) m[PS::out_index(e)] = M::magnitude(r.v[e], a.scale);
        if (MODE == SPEC_PEAK) {
            float bv = m[0];
            int bk = 0;
#pragma unroll
            for (int k = 1; k < N; k++)
                if (m[k] > bv) { bv = m[k]; bk = k; }                        /* the first maximum wins */
            *a.peakVal = sqrtf(bv);
            *a.peakIdx = (uint32_t)bk;
        } else {
#pragma unroll
            for (int g = 0; g < N / 4; g++)
                *reinterpret_cast<Vec16<float> *>(a.mag + 4 * g) = Vec16<float>{{m[4 * g], m[4 * g + 1], m[4 * g + 2], m[4 * g + 3]}};
        }
    }
};

}  // namespace b200fft
