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
#include "fft_frame.cuh"

namespace b200fft {

#if defined(__CUDA_ARCH__)
template <class V> FFT_HD V ld_stream(const V *p);
template <> FFT_HD cf32 ld_stream<cf32>(const cf32 *p) { float2 v = __ldcs(reinterpret_cast<const float2 *>(p)); return {v.x, v.y}; }
template <> FFT_HD ci32 ld_stream<ci32>(const ci32 *p) { int2 v = __ldcs(reinterpret_cast<const int2 *>(p)); return {v.x, v.y}; }
template <> FFT_HD ci16 ld_stream<ci16>(const ci16 *p) { short2 v = __ldcs(reinterpret_cast<const short2 *>(p)); return {v.x, v.y}; }
FFT_HD void st_stream(cf32 *p, cf32 v) { __stcs(reinterpret_cast<float2 *>(p), make_float2(v.x, v.y)); }
FFT_HD void st_stream(ci32 *p, ci32 v) { __stcs(reinterpret_cast<int2 *>(p), make_int2(v.x, v.y)); }
FFT_HD void st_stream(ci16 *p, ci16 v) { __stcs(reinterpret_cast<short2 *>(p), make_short2(v.x, v.y)); }
#else
template <class V> FFT_HD V ld_stream(const V *p) { return *p; }
template <class V> FFT_HD void st_stream(V *p, V v) { *p = v; }
#endif

/* frame input: STAGED = the frame was brought into shared memory by a bulk (TMA) copy and `p`
 * points there; otherwise `p` is global memory and is read once with a streaming hint */
template <bool STAGED, class V> FFT_HD V ld_in(const V *p)
{
    if (STAGED) return *p;
    return ld_stream(p);
}

/* number of phases for a plan with NP passes: 1 -> 1, 2 -> 2, 3 -> 4 (the middle pass is split
 * into load+compute / store so the single exchange buffer can be reused) */
template <int NP> struct PhaseCount { static constexpr int value = (NP == 1) ? 1 : (NP == 2 ? 2 : 4); };

template <class T> struct IsF32 { static constexpr bool value = false; };
template <> struct IsF32<cf32> { static constexpr bool value = true; };

/* ------------------------------------------------------------------ CFFT */

template <class PL, bool INV, bool STAGED = false> struct CfftBody {
    typedef Engine<PL> Eng;
    typedef typename Eng::A A;
    typedef typename A::elem elem;
    typedef typename A::work work;
    typedef typename Eng::Regs Regs;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N;
    static constexpr int kPhases = PhaseCount<NP>::value;
    static constexpr bool kF32 = IsF32<elem>::value;

    struct Args {
        const elem *in;          /* frame base (device or emulated) */
        elem *out;               /* may alias in */
        const elem *tw;          /* pass-ordered twiddle table of this plan (Plan::build_twiddles) */
        const uint16_t *perm;    /* null => natural order; else destination position of X[k] */
        float scale;             /* f32 inverse: 1/N */
        int shl1;                /* fixed point, N = 2*4^m: final << 1 of every word */
    };
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
                work w = A::load(ld_in<STAGED>(a.in + Eng::template in_index<0>(i, b, e)));
                if (kF32 && INV) w.y = -w.y;                       /* conjugate input (cfft_f32.c:1252-1261) */
                r.v[b * PS::R + e] = w;
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
                work w = r.v[b * PS::R + e];
                if (kF32) {
                    if (INV) w = scale_conj(w, a.scale);             /* cfft_f32.c:1285-1297 */
                } else if (a.shl1) {
                    w = A::shl1(w);                                  /* cfft_q31.c:803-820, cfft_q15.c:810-827 */
                }
                const int pos = a.perm ? (int)a.perm[k] : k;
                st_stream(a.out + pos, A::store(w));
            }
    }
    static FFT_HD cf32 scale_conj(cf32 w, float s) { return {w.x * s, -w.y * s}; }
    static FFT_HD ci32 scale_conj(ci32 w, float) { return w; }

    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, elem *sm, int i)
    {
        if constexpr (PH == 0) {
            gload(r, a, i);
            Eng::template compute<0, INV>(r, a.tw, i);
            if constexpr (NP == 1) gstore(r, a, i);
            else Eng::template smem_store<0>(r, sm, i);
        } else if constexpr (NP == 2) {
            Eng::template smem_load<1>(r, sm, i);
            Eng::template compute<1, INV>(r, a.tw, i);
            gstore(r, a, i);
        } else if constexpr (NP == 3) {
            if constexpr (PH == 1) {
                Eng::template smem_load<1>(r, sm, i);
                Eng::template compute<1, INV>(r, a.tw, i);
            } else if constexpr (PH == 2) {
                Eng::template smem_store<1>(r, sm, i);
            } else {
                Eng::template smem_load<2>(r, sm, i);
                Eng::template compute<2, INV>(r, a.tw, i);
                gstore(r, a, i);
            }
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
 * (k, Nh-k) pair is thread-local and shares one rfft twiddle (up to the mirror sign).  Pair 0 of
 * thread 0 is (0, NBF/2): butterfly 0 pairs e <-> R-e (e = 0 is the packed DC/Nyquist bin,
 * e = R/2 pairs with itself), butterfly NBF/2 pairs e <-> R-1-e.  One table entry is LOADED per
 * butterfly pair; the entries of the other bins are compile-time rotations of it. */
template <class PL, bool STAGED = false> struct RfftFwdBody {
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef cf32 elem;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;   /* N = complex length = real length / 2 */
    static constexpr int kPhases = PhaseCount<NP>::value;
    typedef typename PassOf<PL, NP - 1>::type PL_LAST;
    static constexpr int R = PL_LAST::R, NBF = N / R, NB = E / R;
    static_assert(PL_LAST::kMirror && NB % 2 == 0, "forward rfft needs a trailing Mirror pass with an even number of butterflies per thread");

    struct Args {
        const cf32 *in;      /* real frame viewed as N complex */
        cf32 *out;           /* packed spectrum: N complex = 2N floats */
        const cf32 *tw;      /* pass-ordered CFFT twiddles of this plan */
        const cf32 *twr;     /* twiddleCoef_rfft_(2N): (sin,cos)(2*pi*k/(2N)), k < N */
    };
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
            for (int e = 0; e < PS::R; e++)
                r.v[b * PS::R + e] = ld_in<STAGED>(a.in + Eng::template in_index<0>(i, b, e));
    }
    template <int E0, int E1> static FFT_HD void pair_loop(const cf32 *A, const cf32 *B, const Args &a, int p, cf32 tw0)
    {
        /* bins k0 = p + e*NBF (in A) and Nh - k0 = (NBF - p) + (R-1-e)*NBF (in B), e = E0..E1-1 */
        if constexpr (E0 < E1) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const int k0 = p + E0 * NBF;
            st_stream(a.out + k0, rfft_split(A[E0], B[R - 1 - E0], tw));
            st_stream(a.out + (N - k0), rfft_split(B[R - 1 - E0], A[E0], rfft_tw_mirror(tw)));
            pair_loop<E0 + 1, E1>(A, B, a, p, tw0);
        }
    }
    /* butterfly 0 of thread 0: bins e*NBF pair with (R-e)*NBF */
    template <int E0> static FFT_HD void self_loop0(const cf32 *A, const Args &a, cf32 tw0)
    {
        if constexpr (E0 <= R / 2) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const int k0 = E0 * NBF;
            st_stream(a.out + k0, rfft_split(A[E0], A[R - E0], tw));
            if (E0 != R / 2) st_stream(a.out + (N - k0), rfft_split(A[R - E0], A[E0], rfft_tw_mirror(tw)));
            self_loop0<E0 + 1>(A, a, tw0);
        }
    }
    /* butterfly NBF/2 of thread 0: bins NBF/2 + e*NBF pair with NBF/2 + (R-1-e)*NBF */
    template <int E0> static FFT_HD void self_loop1(const cf32 *B, const Args &a, cf32 tw0)
    {
        if constexpr (E0 < R / 2) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const int k0 = NBF / 2 + E0 * NBF;
            st_stream(a.out + k0, rfft_split(B[E0], B[R - 1 - E0], tw));
            st_stream(a.out + (N - k0), rfft_split(B[R - 1 - E0], B[E0], rfft_tw_mirror(tw)));
            self_loop1<E0 + 1>(B, a, tw0);
        }
    }
    static FFT_HD void split_store(const Regs &r, const Args &a, int i)
    {
#pragma unroll
        for (int m = 0; m < NB / 2; m++) {
            const int p = i + T * m;
            const cf32 *A = &r.v[(2 * m) * R], *B = &r.v[(2 * m + 1) * R];
            if (m != 0 || i != 0) {
                pair_loop<0, R>(A, B, a, p, a.twr[p]);
            } else {
                const cf32 X0 = A[0];
                st_stream(a.out + 0, cf32{X0.x + X0.y, X0.x - X0.y});       /* rfft_fast_f32.c:337-352 */
                self_loop0<1>(A, a, a.twr[0]);
                self_loop1<0>(B, a, a.twr[NBF / 2]);
            }
        }
    }
    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, cf32 *sm, int i)
    {
        if constexpr (PH == 0) {
            gload(r, a, i);
            Eng::template compute<0, false>(r, a.tw, i);
            if constexpr (NP == 1) split_store(r, a, i);
            else Eng::template smem_store<0>(r, sm, i);
        } else if constexpr (NP == 2) {
            Eng::template smem_load<1>(r, sm, i);
            Eng::template compute<1, false>(r, a.tw, i);
            split_store(r, a, i);
        } else if constexpr (NP == 3) {
            if constexpr (PH == 1) {
                Eng::template smem_load<1>(r, sm, i);
                Eng::template compute<1, false>(r, a.tw, i);
            } else if constexpr (PH == 2) {
                Eng::template smem_store<1>(r, sm, i);
            } else {
                Eng::template smem_load<2>(r, sm, i);
                Eng::template compute<2, false>(r, a.tw, i);
                split_store(r, a, i);
            }
        }
    }
};

template <class PL, bool STAGED = false> struct RfftInvBody {
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef cf32 elem;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N, T = PL::T;
    static constexpr int kPhases = PhaseCount<NP>::value;
    static constexpr int R = PL::P0::R, NBF = N / R, NB = E / R;
    static_assert(PL::P0::kMirror && NB % 2 == 0, "inverse rfft needs a leading Mirror pass with an even number of butterflies per thread");

    struct Args {
        const cf32 *in;      /* packed spectrum, N complex */
        cf32 *out;           /* real frame viewed as N complex */
        const cf32 *tw;
        const cf32 *twr;
        float scale;         /* 1/N */
    };
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

    template <int E0, int E1> static FFT_HD void pair_loop(cf32 *A, cf32 *B, cf32 tw0)
    {
        if constexpr (E0 < E1) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const cf32 ga = A[E0], gb = B[R - 1 - E0];
            A[E0] = mconj(rfft_merge(ga, gb, tw));
            B[R - 1 - E0] = mconj(rfft_merge(gb, ga, rfft_tw_mirror(tw)));
            pair_loop<E0 + 1, E1>(A, B, tw0);
        }
    }
    template <int E0> static FFT_HD void self_loop0(cf32 *A, cf32 tw0)
    {
        if constexpr (E0 <= R / 2) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const cf32 ga = A[E0], gb = A[R - E0];
            A[E0] = mconj(rfft_merge(ga, gb, tw));
            if (E0 != R / 2) A[R - E0] = mconj(rfft_merge(gb, ga, rfft_tw_mirror(tw)));
            self_loop0<E0 + 1>(A, tw0);
        }
    }
    template <int E0> static FFT_HD void self_loop1(cf32 *B, cf32 tw0)
    {
        if constexpr (E0 < R / 2) {
            const cf32 tw = rfft_tw_rot<R, E0>(tw0);
            const cf32 ga = B[E0], gb = B[R - 1 - E0];
            B[E0] = mconj(rfft_merge(ga, gb, tw));
            B[R - 1 - E0] = mconj(rfft_merge(gb, ga, rfft_tw_mirror(tw)));
            self_loop1<E0 + 1>(B, tw0);
        }
    }
    static FFT_HD void merge_load(Regs &r, const Args &a, int i)
    {
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int e = 0; e < R; e++) r.v[b * R + e] = ld_in<STAGED>(a.in + Eng::template in_index<0>(i, b, e));
#pragma unroll
        for (int m = 0; m < NB / 2; m++) {
            const int p = i + T * m;
            cf32 *A = &r.v[(2 * m) * R], *B = &r.v[(2 * m + 1) * R];
            if (m != 0 || i != 0) {
                pair_loop<0, R>(A, B, a.twr[p]);
            } else {
                const cf32 g0 = A[0];
                A[0] = mconj(cf32{0.5f * (g0.x + g0.y), 0.5f * (g0.x - g0.y)});   /* :425-431 */
                self_loop0<1>(A, a.twr[0]);
                self_loop1<0>(B, a.twr[NBF / 2]);
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
                cf32 w = r.v[b * PS::R + e];
                st_stream(a.out + k, cf32{w.x * a.scale, -w.y * a.scale});
            }
    }
    template <int PH> static FFT_HD void phase(Regs &r, const Args &a, cf32 *sm, int i)
    {
        if constexpr (PH == 0) {
            merge_load(r, a, i);
            Eng::template compute<0, false>(r, a.tw, i);
            if constexpr (NP == 1) gstore(r, a, i);
            else Eng::template smem_store<0>(r, sm, i);
        } else if constexpr (NP == 2) {
            Eng::template smem_load<1>(r, sm, i);
            Eng::template compute<1, false>(r, a.tw, i);
            gstore(r, a, i);
        } else if constexpr (NP == 3) {
            if constexpr (PH == 1) {
                Eng::template smem_load<1>(r, sm, i);
                Eng::template compute<1, false>(r, a.tw, i);
            } else if constexpr (PH == 2) {
                Eng::template smem_store<1>(r, sm, i);
            } else {
                Eng::template smem_load<2>(r, sm, i);
                Eng::template compute<2, false>(r, a.tw, i);
                gstore(r, a, i);
            }
        }
    }
};

}  // namespace b200fft
