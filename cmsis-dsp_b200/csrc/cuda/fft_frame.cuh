/*
 * fft_frame.cuh -- one-frame-per-thread-group Stockham FFT engine (host/device).
 *
 * A frame of N complex points is processed by T threads, E = N/T points per thread,
 * in NP <= 3 register passes separated by shared-memory exchanges:
 *
 *   pass p (radix R, s = product of the radices before it), butterfly j in [0, N/R):
 *       reads   x[j + t*N/R],               t = 0..R-1         (always lane-contiguous)
 *       p = j / s, q = j % s
 *       writes  y[q + s*(R*p + t')] = out_t' ,  out = DIF butterfly (twiddle AFTER the
 *               butterfly: out_t' = DFT_R(x)[t'] * W_N^(s*p*t'))
 *
 * which is the decimation-in-frequency Stockham autosort recursion: same butterflies,
 * same twiddles and same operand order as the reference's in-place DIF passes
 * (arm_radix8_butterfly_f32 / arm_radix4_butterfly_q31 / _q15), but results come out in
 * natural order, so arm_bitreversal2.c's swap pass disappears (bitReverseFlag = 0 is
 * served by scattering through the plan's permutation instead).
 *
 * Twiddles are read from a PASS-ORDERED copy of the reference tables (Plan::build_twiddles):
 * entry [slot][j] for butterfly j, so a warp's twiddle load is one contiguous 256-byte
 * request instead of a gather with stride t*8 bytes.
 *
 * The first pass reads HBM directly into registers (lane-contiguous => coalesced) and
 * the last pass writes HBM directly from registers; only the NP-1 exchanges touch
 * shared memory (padded layout, see Plan::pad).
 */
#pragma once
#include "fft_arith.cuh"

#ifndef FFT_TRACE_SMEM
#define FFT_TRACE_SMEM(ptr, bytes, is_store) ((void)0)
#endif

namespace b200fft {

/* ---------------------------------------------------------------- pass descriptors */

/* f32 pass: radix-R butterfly with the post-twiddle factored per radix-4 level (DftRec),
 * twiddles from the N-entry (cos,+sin) table of the reference */
template <int R_> struct PassF32 {
    static constexpr int R = R_;
    static constexpr bool kMirror = false;
    static constexpr bool kDirect = false;
    typedef cf32 telem;
    static FFT_HD constexpr int out_index(int e) { return e; }
    /* twiddle slots per butterfly in the pass-ordered table (none on the last pass: W^0) */
    static constexpr int slots(bool lastPass) { return lastPass ? 0 : dft_tw_slots(R); }
    /* butterfly j, e = S*(j/S): level l reads W^(4^l e m), m = 1..3; the base level W^(4^L e t) */
    template <int N, int S> static void fill(const cf32 *base, cf32 *out, bool lastPass)
    {
        constexpr int NBF = N / R;
        if (lastPass) return;
        for (int j = 0; j < NBF; j++) {
            int s = 0, r = R;
            long ec = (long)S * (j / S);
            for (; r > 4; r /= 4, ec = (ec * 4) % N)
                for (int m = 1; m <= 3; m++) out[(s++) * NBF + j] = base[(ec * m) % N];
            for (int t = 1; t < r; t++) out[(s++) * NBF + j] = base[(ec * t) % N];
        }
    }
    /* the butterfly's twiddle values into registers / the butterfly on preloaded values */
    template <int N, bool LASTPASS> static FFT_HD void load_tw(cf32 *w, const cf32 *__restrict__ twp, int j)
    {
        constexpr int NBF = N / R;
#pragma unroll
        for (int s = 0; s < slots(LASTPASS); s++) w[s] = twp[s * NBF + j];
    }
    template <bool INV, int N, bool LASTPASS> static FFT_HD void compute_w(cf32 *x, const cf32 *w)
    {
        dft_f32<R, !LASTPASS>(x, w);
    }
    template <bool INV, int N, bool LASTPASS>
    static FFT_HD void compute(cf32 *x, const cf32 *__restrict__ twp, int j)
    {
        cf32 w[dft_tw_slots(R)];
        load_tw<N, LASTPASS>(w, twp, j);
        compute_w<INV, N, LASTPASS>(x, w);
    }
};
/* pass whose butterflies come in mirror pairs (j, N/R - j) per thread: used next to the real
 * side of arm_rfft_fast_f32 so split/merge stay thread-local */
template <int R_> struct PassF32Mirror : PassF32<R_> { static constexpr bool kMirror = true; };
typedef PassF32Mirror<8> PassF32Mirror8;

/* fixed-point pass: one, two or three of the reference's DIF stages, executed back to back on
 * the R = ra*rb*rc points held by the thread.  Point t = u + rc*(w + rb*v) (u < rc, w < rb,
 * v < ra) sits at frame position j + t*N/R: stage a butterflies run over v, stage b over w,
 * stage c over u -- spans N/ra, N/(ra rb), N/R of the reference's in-place passes. */
template <class ARITH, int KA, int KB = -1, int KC = -1> struct PassFix {
    static constexpr int ra = (KA == ST_PRE2) ? 2 : 4;
    static constexpr int rb = (KB < 0) ? 1 : 4;
    static constexpr int rc = (KC < 0) ? 1 : 4;
    static constexpr int R = ra * rb * rc;
    static constexpr int kInShift = (KA == ST_PRE2) ? 1 : ((KA == ST_FIRST4) ? 2 : 0);   /* q15: input shift of stage a */
    static constexpr bool kMirror = false;
    typedef typename ARITH::work work;
    typedef typename ARITH::twid twid;
    typedef typename ARITH::telem telem;    /* device twiddle table element (q15: expanded to 32-bit pairs) */
    typedef typename ARITH::elem selem;     /* element of the reference-layout source table */

    /* element e = u + rc*(w + rb*v) holds residue v + ra*(w + rb*u) after the pass */
    static FFT_HD constexpr int out_index(int e) { return (e / (rb * rc)) + ra * (((e / rc) % rb) + rb * (e % rc)); }

    static constexpr int tw_of(int K) { return K < 0 ? 0 : (K == ST_PRE2 ? 1 : (K == ST_LAST4 ? 0 : 3)); }
    static constexpr int ta = tw_of(KA), tb = tw_of(KB), tc = tw_of(KC);   /* twiddles per butterfly of each stage */
    static constexpr int kOffB = rb * rc * ta, kOffC = kOffB + rc * tb;
    /* kDirect (Arith::kDirectTw): the pass reads the REFERENCE-layout table at the reference's own indices
     * (compute_direct) instead of a pass-ordered copy.  The pass-ordered copy makes every warp load contiguous
     * but is ~3 N entries per plan; with 16-byte entries (f64, N = 4096: 184 KiB against a 64 KiB table) it
     * falls out of L1 and every frame re-reads it from L2 -- more bytes than the frame itself. */
    static constexpr bool kDirect = ARITH::kDirectTw;
    static constexpr int slots(bool) { return kDirect ? 0 : kOffC + tc; }

    /* pass-ordered table of butterfly j, sp = S*(j/S):
     *   stage a, sub-butterfly o < rb*rc : slots o*ta + m-1        = W^(m * (sp + (N/R) o))
     *   stage b, sub-butterfly u < rc    : slots kOffB + u*tb + m-1 = W^(m * ra * (sp + (N/R) u))
     *   stage c                          : slots kOffC + m-1        = W^(m * ra * rb * sp)
     * i.e. the entries ia, 2ia, 3ia the reference reads from twiddleCoef_N_q31/_q15 with its
     * twidCoefModifier (arm_cfft_radix4_q31.c:229-266,307-317; arm_cfft_q31.c:777-801) */
    template <int N, int S> static void fill(const selem *base, telem *out, bool)
    {
        constexpr int NBF = N / R;
        for (int j = 0; j < (kDirect ? 0 : NBF); j++) {
            const int sp = S * (j / S);
            for (int o = 0; o < rb * rc; o++)
                for (int m = 1; m <= ta; m++) out[(o * ta + m - 1) * NBF + j] = ARITH::tw_expand(base[m * (sp + (N / R) * o)]);
            for (int u = 0; u < rc; u++)
                for (int m = 1; m <= tb; m++) out[(kOffB + u * tb + m - 1) * NBF + j] = ARITH::tw_expand(base[m * ra * (sp + (N / R) * u)]);
            for (int m = 1; m <= tc; m++) out[(kOffC + m - 1) * NBF + j] = ARITH::tw_expand(base[m * ra * rb * sp]);
        }
    }

    /* TAIL: the stage is the last one of its pass (its results go to shared or global memory next) */
    template <int K, bool INV, int NBF, bool TAIL, bool PRE = false>
    static FFT_HD void stage4(work &a, work &b, work &c, work &d, const telem *__restrict__ twp, int slot, int j, int32_t ob)
    {
        if (K == ST_LAST4) {
            twid z = {0, 0};
            ARITH::template bfly4<K, INV, TAIL>(a, b, c, d, z, z, z, 0);
        } else {
            ARITH::template bfly4<K, INV, TAIL, PRE>(a, b, c, d, ARITH::tload(twp[slot * NBF + j]), ARITH::tload(twp[(slot + 1) * NBF + j]),
                                          ARITH::tload(twp[(slot + 2) * NBF + j]), ob);
        }
    }
    /* Bias carried by the outputs of a butterfly (Arith::kBiased: ArithQ15): +32768 when they are points a or b of
     * their next butterfly.  Inside a pass that is a compile-time fact of the register index: stage b reads
     * (x[n], x[n + rc], x[n + 2 rc], x[n + 3 rc]) as (a, b, c, d), stage c reads x[4 g .. 4 g + 3]. */
    static constexpr int32_t kBias = ARITH::kBiased ? 32768 : 0;
    static FFT_HD constexpr int32_t bias_into_b(int o) { return ((o / rc) % 4 < 2) ? kBias : 0; }   /* outputs of stage a, butterfly o */
    static FFT_HD constexpr int32_t bias_into_c(int u) { return (u % 4 < 2) ? kBias : 0; }          /* outputs of stage b, column u */

    /* radix-4 stage on the reference-layout table: W^1, W^2, W^3 at ia, 2 ia, 3 ia (arm_cfft_radix4_q31.c:229-266) */
    template <int K, bool INV, bool TAIL, bool PRE = false>
    static FFT_HD void stage4d(work &a, work &b, work &c, work &d, const telem *__restrict__ tw, int ia, int32_t ob)
    {
        if (K == ST_LAST4) {
            twid z = {0, 0};
            ARITH::template bfly4<K, INV, TAIL>(a, b, c, d, z, z, z, 0);
        } else {
            ARITH::template bfly4<K, INV, TAIL, PRE>(a, b, c, d, ARITH::tload(tw[ia]), ARITH::tload(tw[2 * ia]), ARITH::tload(tw[3 * ia]), ob);
        }
    }
    /* same butterflies as compute(), twiddle indices as in fill(): S = product of the radices of the earlier passes */
    template <bool INV, int N, int S, bool PRE = false>
    static FFT_HD void compute_direct(work *x, const telem *__restrict__ tw, int j, int32_t tailBias = 0)
    {
        constexpr int NBF = N / R, Q = rb * rc;
        const int sp = S * (j / S);
#pragma unroll
        for (int o = 0; o < Q; o++) {                 /* stage a */
            const int ia = sp + NBF * o;
            if (KA == ST_PRE2)
                ARITH::template bfly2<INV, PRE>(x[o], x[o + Q], ARITH::tload(tw[ia]));
            else
                stage4d<KA, INV, (KB < 0), PRE>(x[o], x[o + Q], x[o + 2 * Q], x[o + 3 * Q], tw, ia, KB < 0 ? tailBias : bias_into_b(o));
        }
        if constexpr (KB >= 0) {                      /* stage b */
#pragma unroll
            for (int v = 0; v < ra; v++)
#pragma unroll
                for (int u = 0; u < rc; u++) {
                    const int base = u + Q * v;
                    stage4d<KB, INV, (KC < 0)>(x[base], x[base + rc], x[base + 2 * rc], x[base + 3 * rc], tw, ra * (sp + NBF * u),
                                               KC < 0 ? tailBias : bias_into_c(u));
                }
        }
        if constexpr (KC >= 0) {                      /* stage c */
#pragma unroll
            for (int g = 0; g < ra * rb; g++)
                stage4d<KC, INV, true>(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3], tw, ra * rb * sp, tailBias);
        }
    }

    /* PRE: the inputs of stage a carry that stage's input shift already (Arith::load_shifted) */
    /* tailBias: bias of the results of the pass's last stage (Engine::compute: 0 in the frame's last pass) */
    template <bool INV, int N, bool LASTPASS, bool PRE = false>
    static FFT_HD void compute(work *x, const telem *__restrict__ twp, int j, int32_t tailBias = 0)
    {
        constexpr int NBF = N / R, Q = rb * rc;
#pragma unroll
        for (int o = 0; o < Q; o++) {                 /* stage a */
            if (KA == ST_PRE2)
                ARITH::template bfly2<INV, PRE>(x[o], x[o + Q], ARITH::tload(twp[o * NBF + j]));
            else
                stage4<KA, INV, NBF, (KB < 0), PRE>(x[o], x[o + Q], x[o + 2 * Q], x[o + 3 * Q], twp, o * ta, j, KB < 0 ? tailBias : bias_into_b(o));
        }
        if constexpr (KB >= 0) {                      /* stage b */
#pragma unroll
            for (int v = 0; v < ra; v++)
#pragma unroll
                for (int u = 0; u < rc; u++) {
                    const int base = u + Q * v;
                    stage4<KB, INV, NBF, (KC < 0)>(x[base], x[base + rc], x[base + 2 * rc], x[base + 3 * rc], twp, kOffB + u * tb, j,
                                                   KC < 0 ? tailBias : bias_into_c(u));
                }
        }
        if constexpr (KC >= 0) {                      /* stage c */
#pragma unroll
            for (int g = 0; g < ra * rb; g++)
                stage4<KC, INV, NBF, true>(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3], twp, kOffC, j, tailBias);
        }
    }
};

/* ---------------------------------------------------------------- plan */

struct NoPass {
    static constexpr int R = 1;
    static constexpr bool kMirror = false;
    static constexpr bool kDirect = false;
    static constexpr int slots(bool) { return 0; }
};

/* entries of the reference-layout twiddle table of an Arith with direct passes */
template <class A, int N> constexpr int direct_entries() { return A::kDirectTw ? A::kTableNum * N / A::kTableDen : 0; }

template <class ARITH_, int N_, int T_, int F_, int PADA_, int PADB_, class P0_, class P1_ = NoPass, class P2_ = NoPass>
struct Plan {
    typedef ARITH_ Arith;
    typedef P0_ P0; typedef P1_ P1; typedef P2_ P2;
    static constexpr int N = N_;          /* complex points per frame */
    static constexpr int T = T_;          /* threads per frame */
    static constexpr int F = F_;          /* frames per CTA */
    static constexpr int E = N / T;       /* points per thread */
    static constexpr int NP = (P1::R == 1) ? 1 : ((P2::R == 1) ? 2 : 3);
    /* N = 2*4^m: the fixed-point transforms start with a radix-2 stage and end with << 1 (arm_cfft_q31.c:803-820) */
    static constexpr bool kOddLog2 = (N_ & 0xAAAAAAAA) != 0;
    static constexpr int S0 = 1, S1 = P0::R, S2 = P0::R * P1::R;
    static_assert(P0::R * P1::R * P2::R == N, "radices must multiply to N");
    static_assert(E % P0::R == 0 && E % P1::R == 0 && E % P2::R == 0, "E must be a multiple of every radix");
    /* padded exchange layout: PADB extra elements after every 2^PADA elements */
    static FFT_HD int pad(int i) { return PADB_ ? (i + ((i >> PADA_) * PADB_)) : i; }
    /* pad(base + off) for a compile-time `off`: when off is a multiple of the padding period the padding of the two
     * terms adds up, so a thread's accesses are ONE padded base plus immediate offsets (the compiler does not
     * discover this by itself: it rebuilt every padded index from scratch, ~6 instructions per point on q15) */
    template <int OFF> static FFT_HD int pad_at(int base, int paddedBase)
    {
        if constexpr (PADB_ == 0) return base + OFF;
        else if constexpr (OFF % (1 << PADA_) == 0) return paddedBase + OFF + (OFF >> PADA_) * PADB_;
        else return pad(base + OFF);
    }
    static constexpr int kPadded = PADB_ ? (N + ((N - 1) >> PADA_) * PADB_ + PADB_) : N;
    /* scratch behind the exchange area for the 2R self-paired bins of an rfft Mirror pass (fft_body.cuh) */
    static constexpr int kSpecial = P0_::kMirror ? 2 * P0_::R : (P1_::kMirror ? 2 * P1_::R : (P2_::kMirror ? 2 * P2_::R : 0));
    /* frames that share a shared-memory wavefront (T lanes each) start T elements apart modulo the
     * wavefront width, so their lanes land in different banks */
    static constexpr int kLanesPerWave = 128 / (int)sizeof(typename ARITH_::xelem);
    static constexpr int kFrameElems = (PADB_ && T_ < kLanesPerWave)
        ? kPadded + ((T_ - kPadded % kLanesPerWave) + kLanesPerWave) % kLanesPerWave : kPadded;
    /* CTA layout: F exchange areas of kFrameElems, then F scratch areas of kSpecial */
    static constexpr int kSmemElems = F * (kFrameElems + kSpecial);
    static constexpr int kSmemBytes = (NP > 1 || kSpecial) ? kSmemElems * (int)sizeof(typename ARITH_::xelem) : 0;
    static constexpr int kThreads = T * F;
    /* the same plan with another number of frames per CTA */
    template <int F2> using with_frames = Plan<ARITH_, N_, T_, F2, PADA_, PADB_, P0_, P1_, P2_>;
    /* pass-ordered twiddle table: pass p owns slots_p * (N / R_p) entries starting at kTwOff<p> */
    static constexpr int kTw0 = P0::slots(NP == 1) * (N / P0::R);
    static constexpr int kTw1 = (NP > 1) ? P1::slots(NP == 2) * (N / P1::R) : 0;
    static constexpr int kTw2 = (NP > 2) ? P2::slots(true) * (N / P2::R) : 0;
    /* direct passes (PassFix::kDirect) share one copy of the reference-layout table */
    static constexpr bool kDirectTw = P0::kDirect;
    static_assert((P1::R == 1 || P1::kDirect == kDirectTw) && (P2::R == 1 || P2::kDirect == kDirectTw), "all passes alike");
    static constexpr int kTwEntries = kDirectTw ? direct_entries<ARITH_, N>() : kTw0 + kTw1 + kTw2;
    /* build the table from the reference-layout twiddles (N entries f32 / f64, 3N/4 entries q31/q15) */
    static void build_twiddles(const typename ARITH_::elem *base, typename ARITH_::telem *out)
    {
        if constexpr (kDirectTw) {
            for (int i = 0; i < kTwEntries; i++) out[i] = ARITH_::tw_expand(base[i]);
            return;
        }
        P0::template fill<N, S0>(base, out, NP == 1);
        if constexpr (NP > 1) P1::template fill<N, S1>(base, out + kTw0, NP == 2);
        if constexpr (NP > 2) P2::template fill<N, S2>(base, out + kTw0 + kTw1, true);
    }
};

template <class PL, int P> struct PassOf;
template <class PL> struct PassOf<PL, 0> { typedef typename PL::P0 type; static constexpr int S = PL::S0, TWOFF = 0; };
template <class PL> struct PassOf<PL, 1> { typedef typename PL::P1 type; static constexpr int S = PL::S1, TWOFF = PL::kTw0; };
template <class PL> struct PassOf<PL, 2> { typedef typename PL::P2 type; static constexpr int S = PL::S2, TWOFF = PL::kTw0 + PL::kTw1; };

/* butterfly index handled by thread i as its b-th butterfly of a pass with NBF butterflies.
 * Mirror passes hand out PAIRS (p, NBF - p), p = i + T*(b/2) in [0, NBF/2); pair 0 is (0, NBF/2). */
template <bool MIRROR, int T, int NBF> FFT_HD int bfly_index(int i, int b)
{
    if (!MIRROR) return i + T * b;
    const int p = i + T * (b >> 1);
    return (b & 1) ? (p == 0 ? NBF / 2 : NBF - p) : p;
}

/* ---------------------------------------------------------------- engine */

template <class PL> struct Engine {
    typedef typename PL::Arith A;
    typedef typename A::elem elem;
    typedef typename A::xelem xelem;
    typedef typename A::work work;
    static constexpr int N = PL::N, T = PL::T, E = PL::E, NP = PL::NP;

    struct Regs { work v[E]; };

    /* ---- shared-memory exchange ---- */
    template <int P, int B0, int E0> static FFT_HD void store_elems(const Regs &r, xelem *sm, int base, int pbase)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int R = PS::R, S = PassOf<PL, P>::S;
        if constexpr (E0 < R) {
            const int idx = PL::template pad_at<S * PS::out_index(E0)>(base, pbase);
            FFT_TRACE_SMEM(&sm[idx], (int)sizeof(xelem), 1);
            sm[idx] = A::xstore(r.v[B0 * R + E0]);
            store_elems<P, B0, E0 + 1>(r, sm, base, pbase);
        }
    }
    template <int P, int B0> static FFT_HD void store_bflies(const Regs &r, xelem *sm, int i)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int R = PS::R, S = PassOf<PL, P>::S, NB = E / R, NBF = N / R;
        if constexpr (B0 < NB) {
            const int j = bfly_index<PS::kMirror, T, NBF>(i, B0);
            const int base = (j % S) + S * R * (j / S);
            store_elems<P, B0, 0>(r, sm, base, PL::pad(base));
            store_bflies<P, B0 + 1>(r, sm, i);
        }
    }
    template <int P> static FFT_HD void smem_store(const Regs &r, xelem *sm, int i) { store_bflies<P, 0>(r, sm, i); }

    template <int P, int B0, int E0> static FFT_HD void load_elems(Regs &r, const xelem *sm, int j, int pj)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int R = PS::R, NBF = N / R;
        if constexpr (E0 < R) {
            const int idx = PL::template pad_at<E0 * NBF>(j, pj);
            FFT_TRACE_SMEM(&sm[idx], (int)sizeof(xelem), 0);
            r.v[B0 * R + E0] = A::xload(sm[idx]);
            load_elems<P, B0, E0 + 1>(r, sm, j, pj);
        }
    }
    template <int P, int B0> static FFT_HD void load_bflies(Regs &r, const xelem *sm, int i)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int R = PS::R, NB = E / R, NBF = N / R;
        if constexpr (B0 < NB) {
            const int j = bfly_index<PS::kMirror, T, NBF>(i, B0);
            load_elems<P, B0, 0>(r, sm, j, PL::pad(j));
            load_bflies<P, B0 + 1>(r, sm, i);
        }
    }
    template <int P> static FFT_HD void smem_load(Regs &r, const xelem *sm, int i) { load_bflies<P, 0>(r, sm, i); }

    /* ---- butterflies of pass P on the registers ---- */
    template <int P, bool INV, bool PRE = false, class TW> static FFT_HD void compute(Regs &r, const TW *__restrict__ tw, int i)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int R = PS::R, S = PassOf<PL, P>::S, NB = E / R, NBF = N / R;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int j = bfly_index<PS::kMirror, T, NBF>(i, b);
            if constexpr (A::kBiased) {
                /* Biased arithmetic (ArithQ15::bfly4): the results of butterfly j land at positions
                 * (j % S) + S (R (j / S) + t), t < R, and the next pass reads position p as point p / (N/4) of a
                 * radix-4 butterfly (a, b, c, d): a or b -- the biased ones -- iff p < N/2 iff j < NBF/2, one
                 * value per butterfly.  The frame's last pass produces plain results. */
                static_assert(P == NP - 1 || (NBF % (2 * S) == 0), "the halves of the next pass must be whole butterfly blocks");
                const int32_t tailBias = (P == NP - 1) ? 0 : (j < NBF / 2 ? 32768 : 0);
                if constexpr (PS::kDirect) PS::template compute_direct<INV, N, S, PRE>(&r.v[b * R], tw, j, tailBias);
                else PS::template compute<INV, N, (P == NP - 1), PRE>(&r.v[b * R], tw + PassOf<PL, P>::TWOFF, j, tailBias);
            } else if constexpr (PRE) {
                if constexpr (PS::kDirect) PS::template compute_direct<INV, N, S, true>(&r.v[b * R], tw, j);
                else PS::template compute<INV, N, (P == NP - 1), true>(&r.v[b * R], tw + PassOf<PL, P>::TWOFF, j);
            } else {
                if constexpr (PS::kDirect) PS::template compute_direct<INV, N, S>(&r.v[b * R], tw, j);
                else PS::template compute<INV, N, (P == NP - 1)>(&r.v[b * R], tw + PassOf<PL, P>::TWOFF, j);
            }
        }
    }

    /* twiddles of pass P held in registers (f32 passes): lets a kernel issue the table loads
     * before it starts waiting for the frame data */
    template <int P> struct TwRegs {
        typedef typename PassOf<PL, P>::type PS;
        static constexpr int kSlots = PS::slots(P == NP - 1), kNB = E / PS::R;
        typename PS::telem w[kNB * kSlots > 0 ? kNB * kSlots : 1];
    };
    template <int P, class TW> static FFT_HD void load_tw(TwRegs<P> &t, const TW *__restrict__ tw, int i)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int NBF = N / PS::R;
#pragma unroll
        for (int b = 0; b < TwRegs<P>::kNB; b++)
            PS::template load_tw<N, (P == NP - 1)>(&t.w[b * TwRegs<P>::kSlots], tw + PassOf<PL, P>::TWOFF,
                                                   bfly_index<PS::kMirror, T, NBF>(i, b));
    }
    template <int P, bool INV> static FFT_HD void compute_pre(Regs &r, const TwRegs<P> &t)
    {
        typedef typename PassOf<PL, P>::type PS;
#pragma unroll
        for (int b = 0; b < TwRegs<P>::kNB; b++)
            PS::template compute_w<INV, N, (P == NP - 1)>(&r.v[b * PS::R], &t.w[b * TwRegs<P>::kSlots]);
    }

    /* index (within the frame) of register slot (b, e) on the input side of pass P / output side */
    template <int P> static FFT_HD int in_index(int i, int b, int e)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int NBF = N / PS::R;
        return bfly_index<PS::kMirror, T, NBF>(i, b) + e * NBF;
    }
    template <int P> static FFT_HD int out_index(int i, int b, int e)
    {
        typedef typename PassOf<PL, P>::type PS;
        constexpr int R = PS::R, S = PassOf<PL, P>::S, NBF = N / R;
        const int j = bfly_index<PS::kMirror, T, NBF>(i, b);
        return (j % S) + S * (R * (j / S) + PS::out_index(e));
    }
};

}  // namespace b200fft
