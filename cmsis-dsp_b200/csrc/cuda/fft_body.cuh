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
 * (arm_rfft_fast_f32.c:372-395) */
FFT_HD cf32 rfft_split(cf32 A, cf32 B, cf32 tw)
{
    float t1a = B.x - A.x, t1b = B.y + A.y;
    float p0 = tw.x * t1a, p1 = tw.y * t1a, p2 = tw.x * t1b, p3 = tw.y * t1b;
    return {0.5f * (A.x + B.x + p0 + p3), 0.5f * (A.y - B.y + p1 - p2)};
}
/* merge stage for one bin (arm_rfft_fast_f32.c:436-455) */
FFT_HD cf32 rfft_merge(cf32 A, cf32 B, cf32 tw)
{
    float t1a = A.x - B.x, t1b = A.y + B.y;
    float r = tw.x * t1a, s = tw.y * t1b, t = tw.y * t1a, u = tw.x * t1b;
    return {0.5f * (A.x + B.x - r - s), 0.5f * (A.y - B.y + t - u)};
}

/* Requires: the pass next to the real side is a Mirror8 pass with 2 butterflies per thread
 * (E == 16), so thread i holds bins {j + t*NBF} and {NBF - j + t*NBF}: every (k, Nh-k) pair is
 * thread-local.  Slot m of butterfly 0 pairs with slot 15-m (thread 0: butterfly 0 pairs
 * t <-> 8-t with t = 0 the packed DC/Nyquist bin, butterfly 1 pairs t <-> 7-t). */
template <class PL, bool STAGED = false> struct RfftFwdBody {
    typedef Engine<PL> Eng;
    typedef typename Eng::Regs Regs;
    typedef cf32 elem;
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N;   /* N = complex length = real length / 2 */
    static constexpr int kPhases = PhaseCount<NP>::value;
    typedef typename PassOf<PL, NP - 1>::type PL_LAST;
    static_assert(PL_LAST::kMirror && E == 16, "forward rfft needs a trailing Mirror8 pass, 16 points per thread");
    static constexpr int NBF = N / 8;

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
    static FFT_HD void split_store(const Regs &r, const Args &a, int i)
    {
        if (i != 0) {
            const int j0 = i, j1 = NBF - i;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const int k0 = j0 + t * NBF, k1 = j1 + t * NBF;
                st_stream(a.out + k0, rfft_split(r.v[t], r.v[15 - t], a.twr[k0]));
                st_stream(a.out + k1, rfft_split(r.v[8 + t], r.v[7 - t], a.twr[k1]));
            }
        } else {
            const cf32 X0 = r.v[0];
            st_stream(a.out + 0, cf32{X0.x + X0.y, X0.x - X0.y});       /* rfft_fast_f32.c:337-352 */
#pragma unroll
            for (int t = 1; t < 8; t++) {
                const int k0 = t * NBF;
                st_stream(a.out + k0, rfft_split(r.v[t], r.v[8 - t], a.twr[k0]));
            }
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const int k1 = NBF / 2 + t * NBF;
                st_stream(a.out + k1, rfft_split(r.v[8 + t], r.v[15 - t], a.twr[k1]));
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
    static constexpr int NP = PL::NP, E = PL::E, N = PL::N;
    static constexpr int kPhases = PhaseCount<NP>::value;
    static_assert(PL::P0::kMirror && E == 16, "inverse rfft needs a leading Mirror8 pass, 16 points per thread");
    static constexpr int NBF = N / 8;

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

    static FFT_HD void merge_load(Regs &r, const Args &a, int i)
    {
        cf32 g[16];
        if (i != 0) {
            const int j0 = i, j1 = NBF - i;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                g[t] = ld_in<STAGED>(a.in + j0 + t * NBF);
                g[8 + t] = ld_in<STAGED>(a.in + j1 + t * NBF);
            }
#pragma unroll
            for (int t = 0; t < 8; t++) {
                r.v[t] = mconj(rfft_merge(g[t], g[15 - t], a.twr[j0 + t * NBF]));
                r.v[8 + t] = mconj(rfft_merge(g[8 + t], g[7 - t], a.twr[j1 + t * NBF]));
            }
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) {
                g[t] = ld_in<STAGED>(a.in + t * NBF);
                g[8 + t] = ld_in<STAGED>(a.in + NBF / 2 + t * NBF);
            }
            r.v[0] = mconj(cf32{0.5f * (g[0].x + g[0].y), 0.5f * (g[0].x - g[0].y)});   /* :425-431 */
#pragma unroll
            for (int t = 1; t < 8; t++) r.v[t] = mconj(rfft_merge(g[t], g[8 - t], a.twr[t * NBF]));
#pragma unroll
            for (int t = 0; t < 8; t++) r.v[8 + t] = mconj(rfft_merge(g[8 + t], g[15 - t], a.twr[NBF / 2 + t * NBF]));
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
