/*
 * tests/emu/emu.cpp -- CPU emulator of the CUDA kernel bodies (TEST INFRASTRUCTURE).
 *
 * Compiles the very same plan / pass / body templates the kernels are built from
 * (cmsis-dsp_b200/csrc/cuda/fft_*.cuh) with g++, and executes a CTA as "for each phase,
 * for each thread", which is exactly what __syncthreads() between phases guarantees on
 * the device.  It lets the non-GPU test-suite check index math and the bit-exact
 * fixed-point arithmetic against the oracle, and it records every shared-memory access
 * so the bank-conflict degree of each exchange can be asserted without a GPU.
 *
 * It is not a fallback: nothing in the product links it.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <type_traits>
#include <vector>

struct TraceRec { int phase, seq, tid, off, bytes, st; };
/* The file is compiled in EMU_PARTS parts (-DEMU_PART=k selects the entry points of part k; no EMU_PART: everything)
 * so that the build runs in parallel (tests/emu/build.py): one translation unit takes minutes.  The trace state is
 * shared between the parts (C++17 inline variables). */
inline bool g_trace_on = false;
inline std::vector<TraceRec> g_trace;
inline const char *g_cta_base = nullptr;
inline int g_cur_phase = 0, g_cur_tid = 0, g_cur_seq = 0;
#if defined(EMU_PART)
#define EMU_HAS(k) (EMU_PART == (k))
#else
#define EMU_HAS(k) 1
#endif

#define FFT_TRACE_SMEM(ptr, bytes, is_store)                                                                  \
    do {                                                                                                      \
        if (g_trace_on)                                                                                       \
            g_trace.push_back({g_cur_phase, g_cur_seq++, g_cur_tid, (int)((const char *)(ptr) - g_cta_base), \
                               (bytes), (is_store)});                                                         \
    } while (0)

#include "fft_plans.cuh"

using namespace b200fft;

template <class BODY, int PH> struct PhaseRunner {
    template <class ARGS, class ELEM>
    static void run(std::vector<typename BODY::Regs> &regs, const ARGS *args, ELEM *smem, int T, int F, int frameElems,
                    const bool *valid)
    {
        g_cur_phase = PH;
        for (int tid = 0; tid < T * F; tid++) {
            int fl = tid / T, i = tid % T;
            g_cur_tid = tid;
            g_cur_seq = 0;
            if (!valid[fl]) continue;
            BODY::template phase<PH>(regs[tid], args[fl], smem + (size_t)fl * frameElems, i);
        }
        if (PH + 1 < BODY::kPhases) PhaseRunner<BODY, (PH + 1 < BODY::kPhases ? PH + 1 : PH)>::run(regs, args, smem, T, F, frameElems, valid);
    }
};

/* run one CTA worth of frames [f0, f0+F) */
template <class PL, class BODY, class MAKEARGS>
static void run_batch(uint64_t nFrames, MAKEARGS make)
{
    typedef typename BODY::xelem elem;
    constexpr int T = PL::T, F = PL::F;
    std::vector<elem> smem((size_t)PL::kSmemElems + 16);
    std::vector<typename BODY::Regs> regs((size_t)T * F);
    std::vector<typename BODY::Args> args(F);
    bool valid[F];
    for (uint64_t f0 = 0; f0 < nFrames; f0 += F) {
        for (int fl = 0; fl < F; fl++) {
            valid[fl] = (f0 + fl) < nFrames;
            if (valid[fl]) {
                args[fl] = make(f0 + fl);
                BODY::set_scratch(args[fl], smem.data() + (size_t)F * PL::kFrameElems + (size_t)fl * PL::kSpecial);
            }
        }
        g_cta_base = (const char *)smem.data();
        PhaseRunner<BODY, 0>::run(regs, args.data(), smem.data(), T, F, PL::kFrameElems, valid);
        if (g_trace_on && f0 == 0) g_trace_on = false;   /* trace only the first CTA */
    }
}

template <class AR, class PL, bool INV, bool PERM>
static void cfft_run_p(typename AR::elem *data, uint64_t nFrames, const void *tw, const uint16_t *perm, int shl1)
{
    /* the body the product uses for this plan: one thread per frame for the single-pass plans */
    typedef typename std::conditional<(PL::NP == 1 && PL::T == 1), TinyCfftBody<PL, INV, PERM>, CfftBody<PL, INV, PERM>>::type BODY;
    typedef typename AR::elem elem;
    std::vector<typename AR::telem> ordered((size_t)PL::kTwEntries + 1);
    PL::build_twiddles((const elem *)tw, ordered.data());     /* same re-ordering the shim uploads */
    run_batch<PL, BODY>(nFrames, [&](uint64_t f) {
        typename BODY::Args a;
        a.in = data + f * PL::N;
        a.out = data + f * PL::N;
        a.tw = ordered.data();
        a.perm = perm;
        a.scale = 1.0f / (float)PL::N;
        a.shl1 = shl1;
        return a;
    });
}
template <class AR, class PL, bool INV>
static void cfft_run(typename AR::elem *data, uint64_t nFrames, const void *tw, const uint16_t *perm, int shl1)
{
    if (perm) cfft_run_p<AR, PL, INV, true>(data, nFrames, tw, perm, shl1);
    else cfft_run_p<AR, PL, INV, false>(data, nFrames, tw, perm, shl1);
}

template <int N> static void cfft_f32_n(float *d, uint64_t n, int inv, const void *tw, const uint16_t *perm)
{
    typedef typename PlanCfftF32<N>::type PL;
    if (inv) cfft_run<ArithF32, PL, true>((cf32 *)d, n, tw, perm, 0);
    else cfft_run<ArithF32, PL, false>((cf32 *)d, n, tw, perm, 0);
}
template <int N> static void cfft_f64_n(double *d, uint64_t n, int inv, const void *tw, const uint16_t *perm)
{
    typedef typename PlanCfftF64<N>::type PL;
    if (inv) cfft_run<typename PL::Arith, PL, true>((cf64 *)d, n, tw, perm, 0);
    else cfft_run<typename PL::Arith, PL, false>((cf64 *)d, n, tw, perm, 0);
}
template <class AR, int N> static void cfft_fix_n(void *d, uint64_t n, int inv, const void *tw, const uint16_t *perm)
{
    typedef typename PlanCfftFix<AR, N>::type PL;
    const int lg = __builtin_ctz(N), shl1 = lg & 1;
    if (inv) cfft_run<AR, PL, true>((typename AR::elem *)d, n, tw, perm, shl1);
    else cfft_run<AR, PL, false>((typename AR::elem *)d, n, tw, perm, shl1);
}


/* the body the product uses by default for (type, complex length n): thread per frame for n <= 64 */
template <class AR, int n>
static void rfft_fix_run(const void *in, void *out, uint64_t nFrames, int ifft, const void *tw, const ci32x4 *coef, int shl1)
{
    typedef typename PlanCfftFix<AR, n>::type PL;
    typedef typename AR::elem elem;
    std::vector<typename AR::telem> ordered((size_t)PL::kTwEntries + 1);
    PL::build_twiddles((const elem *)tw, ordered.data());
    if constexpr (PL::NP == 1 && PL::T == 1) {
        if (!ifft) {
            typedef TinyRfftFixBody<PL, false> BODY;
            run_batch<PL, BODY>(nFrames, [&](uint64_t f) {
                return typename BODY::Args{(const elem *)in + f * n, (elem *)out + f * 2 * n, ordered.data(), coef, shl1};
            });
        } else {
            typedef TinyRfftFixBody<PL, true> BODY;
            run_batch<PL, BODY>(nFrames, [&](uint64_t f) {
                return typename BODY::Args{(const elem *)in + f * 2 * n, (elem *)out + f * n, ordered.data(), coef, shl1};
            });
        }
    } else if (!ifft) {
        typedef RfftFixFwdBody<PL> BODY;
        run_batch<PL, BODY>(nFrames, [&](uint64_t f) {
            return typename BODY::Args{(const elem *)in + f * n, (elem *)out + f * 2 * n, ordered.data(), coef, shl1};
        });
    } else {
        typedef CfftBody<PL, true, false, false, true> BODY;
        run_batch<PL, BODY>(nFrames, [&](uint64_t f) {
            return typename BODY::Args{(const elem *)in + f * 2 * n, (elem *)out + f * n, ordered.data(), nullptr, 0.0f, shl1, coef};
        });
    }
}

#define FOR_ALL_N(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)
#define FOR_RFFT_NC(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

extern "C" {

#if EMU_HAS(0)
int emu_cfft(int type, uint32_t N, void *data, uint64_t nFrames, int ifft, int bitrev, const void *tw, const uint16_t *perm)
{
    const uint16_t *pp = bitrev ? nullptr : perm;
    int inv = (ifft == 1);
    switch (N) {
#define CASE(n)                                                                      \
    case n:                                                                          \
        if (type == 0) cfft_f32_n<n>((float *)data, nFrames, inv, tw, pp);          \
        else if (type == 1) cfft_fix_n<ArithQ31, n>(data, nFrames, inv, tw, pp);    \
        else if (type == 3) cfft_f64_n<n>((double *)data, nFrames, inv, tw, pp);    \
        else cfft_fix_n<ArithQ15, n>(data, nFrames, inv, tw, pp);                   \
        return 0;
        FOR_ALL_N(CASE)
#undef CASE
    default: return -1;
    }
}
#endif

#if EMU_HAS(1)
int emu_rfft(uint32_t Nreal, const float *in, float *out, uint64_t nFrames, int ifft, const void *tw, const void *twr)
{
    switch (Nreal / 2) {
#define CASE(nc)                                                                                     \
    case nc: {                                                                                       \
        if (!ifft) {                                                                                 \
            typedef std::conditional<(nc <= 64), PlanCfftF32<(nc <= 64 ? nc : 16)>::type, PlanRfftFwd<nc>::type>::type PL; \
            typedef std::conditional<(nc <= 64), TinyRfftFwdBody<PlanCfftF32<(nc <= 64 ? nc : 16)>::type>, RfftFwdBody<PlanRfftFwd<nc>::type>>::type BODY; \
            std::vector<cf32> ordered((size_t)PL::kTwEntries + 1);                                   \
            PL::build_twiddles((const cf32 *)tw, ordered.data());                                    \
            run_batch<PL, BODY>(nFrames, [&](uint64_t f) {                                           \
                BODY::Args a;                                                                        \
                a.in = (const cf32 *)in + f * nc; a.out = (cf32 *)out + f * nc;                      \
                a.tw = ordered.data(); a.twr = (const cf32 *)twr;                                    \
                return a;                                                                            \
            });                                                                                      \
        } else {                                                                                     \
            typedef std::conditional<(nc <= 64), PlanCfftF32<(nc <= 64 ? nc : 16)>::type, PlanRfftInv<nc>::type>::type PL; \
            typedef std::conditional<(nc <= 64), TinyRfftInvBody<PlanCfftF32<(nc <= 64 ? nc : 16)>::type>, RfftInvBody<PlanRfftInv<nc>::type>>::type BODY; \
            std::vector<cf32> ordered((size_t)PL::kTwEntries + 1);                                   \
            PL::build_twiddles((const cf32 *)tw, ordered.data());                                    \
            run_batch<PL, BODY>(nFrames, [&](uint64_t f) {                                           \
                BODY::Args a;                                                                        \
                a.in = (const cf32 *)in + f * nc; a.out = (cf32 *)out + f * nc;                      \
                a.tw = ordered.data(); a.twr = (const cf32 *)twr; a.scale = 1.0f / (float)nc;        \
                return a;                                                                            \
            });                                                                                      \
        }                                                                                            \
        return 0;                                                                                    \
    }
        FOR_RFFT_NC(CASE)
#undef CASE
    default: return -1;
    }
}
#endif

/* arm_rfft_q31 (type 1) / arm_rfft_q15 (type 2): Nreal = real length 32..8192; forward frames Nreal
 * scalars in -> 2*Nreal scalars out, inverse frames 2*Nreal scalars in -> Nreal out.  tw = the
 * reference-layout twiddles of the Nreal/2-point CFFT, coefA/B = realCoefA/B tables (8192 entries). */
#if EMU_HAS(2)
int emu_rfft_fix(int type, uint32_t Nreal, const void *in, void *out, uint64_t nFrames, int ifft, const void *tw,
                 const void *coefA, const void *coefB)
{
    const uint32_t L2 = Nreal / 2, mod = 8192u / Nreal;
    std::vector<ci32x4> coef(L2);
    for (uint32_t k = 0; k < L2; k++) {
        if (type == 1) {
            const int32_t *A = (const int32_t *)coefA, *B = (const int32_t *)coefB;
            coef[k] = ci32x4{A[2 * k * mod], A[2 * k * mod + 1], B[2 * k * mod], B[2 * k * mod + 1]};
        } else {
            const int16_t *A = (const int16_t *)coefA, *B = (const int16_t *)coefB;
            coef[k] = ci32x4{A[2 * k * mod], A[2 * k * mod + 1], B[2 * k * mod], B[2 * k * mod + 1]};
        }
    }
    const int shl1 = __builtin_ctz(L2) & 1;
    switch (L2) {
#define CASE(n)                                                     \
    case n:                                                         \
        if (type == 1) rfft_fix_run<ArithQ31, n>(in, out, nFrames, ifft, tw, coef.data(), shl1); \
        else rfft_fix_run<ArithQ15, n>(in, out, nFrames, ifft, tw, coef.data(), shl1);           \
        return 0;
        FOR_ALL_N(CASE)
#undef CASE
    default: return -1;
    }
}
#endif

/* arm_cfft_f32 + arm_cmplx_mag[_squared]_f32 fused (CfftMagBody, modes 0 / 1; the peak mode reduces with warp
 * shuffles and is exercised on the device only) */
#if EMU_HAS(3)
int emu_cfft_mag(uint32_t N, const float *in, float *mag, uint64_t nFrames, int ifft, int squared, const void *tw)
{
    switch (N) {
#define RUNM(n, INV, MODE)                                                                         \
    {                                                                                              \
        typedef PlanCfftF32<n>::type PL;                                                           \
        typedef std::conditional<(PL::NP == 1 && PL::T == 1), TinyCfftMagBody<PL, INV, MODE>, CfftMagBody<PL, INV, MODE>>::type BODY; \
        std::vector<cf32> ordered((size_t)PL::kTwEntries + 1);                                     \
        PL::build_twiddles((const cf32 *)tw, ordered.data());                                      \
        run_batch<PL, BODY>(nFrames, [&](uint64_t f) {                                             \
            BODY::Args a{};                                                                        \
            a.in = (const cf32 *)in + f * n; a.tw = ordered.data(); a.scale = 1.0f / (float)n;     \
            a.mag = mag + f * n;                                                                   \
            return a;                                                                              \
        });                                                                                        \
    }
#define CASE(n)                                                                                    \
    case n:                                                                                        \
        if (ifft) { if (squared) RUNM(n, true, SPEC_MAG_SQUARED) else RUNM(n, true, SPEC_MAG) }    \
        else      { if (squared) RUNM(n, false, SPEC_MAG_SQUARED) else RUNM(n, false, SPEC_MAG) }  \
        return 0;
        FOR_ALL_N(CASE)
#undef CASE
#undef RUNM
    default: return -1;
    }
}
#endif

#if EMU_HAS(4)
int emu_rfft64(uint32_t Nreal, const double *in, double *out, uint64_t nFrames, int ifft, const void *tw, const void *twr)
{
    switch (Nreal / 2) {
#define CASE(nc)                                                                                     \
    case nc: {                                                                                       \
        typedef PlanCfftF64<nc>::type PL;                                                            \
        std::vector<cf64> ordered((size_t)PL::kTwEntries + 1);                                       \
        PL::build_twiddles((const cf64 *)tw, ordered.data());                                        \
        if (!ifft) {                                                                                 \
            typedef RfftF64FwdBody<PL> BODY;                                                         \
            run_batch<PL, BODY>(nFrames, [&](uint64_t f) {                                           \
                return BODY::Args{(const cf64 *)in + f * nc, (cf64 *)out + f * nc, ordered.data(), (const cf64 *)twr}; \
            });                                                                                      \
        } else {                                                                                     \
            typedef RfftF64InvBody<PL> BODY;                                                         \
            run_batch<PL, BODY>(nFrames, [&](uint64_t f) {                                           \
                return BODY::Args{(const cf64 *)in + f * nc, (cf64 *)out + f * nc, ordered.data(), (const cf64 *)twr, 1.0f / (float)nc}; \
            });                                                                                      \
        }                                                                                            \
        return 0;                                                                                    \
    }
        FOR_RFFT_NC(CASE)
#undef CASE
    default: return -1;
    }
}
#endif

#if EMU_HAS(4)
void emu_trace_begin(void) { g_trace.clear(); g_trace_on = true; }

/* Per (phase, is_store): number of warp-level requests and the shared-memory wavefronts they
 * need (32 banks x 4 B; a request of 8 B/lane is served per half-warp, 16 B/lane per quarter
 * warp).  out rows: phase, is_store, bytes, requests, ideal_wavefronts, wavefronts. Returns rows. */
int emu_trace_stats(int64_t *out, int maxRows)
{
    struct Key { int phase, st, bytes; bool operator<(const Key &o) const { return phase != o.phase ? phase < o.phase : (st != o.st ? st < o.st : bytes < o.bytes); } };
    std::map<Key, std::map<std::pair<int, int>, std::vector<TraceRec>>> groups;   /* key -> (warp, seq) -> lanes */
    for (const TraceRec &r : g_trace) groups[{r.phase, r.st, r.bytes}][{r.tid / 32, r.seq}].push_back(r);
    int row = 0;
    for (auto &kv : groups) {
        int64_t req = 0, ideal = 0, wf = 0;
        const int lanesPerPhase = kv.first.bytes <= 4 ? 32 : (kv.first.bytes == 8 ? 16 : 8);
        for (auto &g : kv.second) {
            req++;
            for (int ph = 0; ph < 32 / lanesPerPhase; ph++) {
                std::map<int, std::vector<int>> bankWords;
                bool any = false;
                for (const TraceRec &r : g.second) {
                    int lane = r.tid % 32;
                    if (lane / lanesPerPhase != ph) continue;
                    any = true;
                    for (int b = 0; b < r.bytes; b += 4) {
                        int word = (r.off + b) / 4;
                        auto &v = bankWords[word % 32];
                        bool seen = false;
                        for (int w : v) seen |= (w == word);
                        if (!seen) v.push_back(word);
                    }
                }
                if (!any) continue;
                int mx = 1;
                for (auto &bw : bankWords) mx = std::max(mx, (int)bw.second.size());
                ideal += 1;
                wf += mx;
            }
        }
        if (row < maxRows) {
            int64_t *o = out + 6 * row;
            o[0] = kv.first.phase; o[1] = kv.first.st; o[2] = kv.first.bytes; o[3] = req; o[4] = ideal; o[5] = wf;
        }
        row++;
    }
    return row;
}
#endif

}  /* extern "C" */
