/*
 * kernel_unit.cu -- ONE (op, length) pair of the batched FFT kernels, compiled once per pair
 * (-DKU_OP=<KernelOp> -DKU_N=<complex length>) so that the 44 pairs build in parallel.
 *
 * One launch processes a whole batch of independent frames: a CTA holds PL::F frames,
 * PL::T threads each (E = N/T points per thread, 16..64), and runs the phases of the plan's
 * body (fft_body.cuh) with a barrier between them.  HBM is touched exactly once per point on
 * the way in and once on the way out; the only other traffic is the shared-memory exchange(s).
 */
#include <cuda_runtime.h>
#include <atomic>
#include <stdlib.h>

#include <type_traits>

#if defined(KU_N) && KU_N == 32
#define FFT_HI32_XU 1        /* q31 high products as IMAD.HI for this length (fft_arith.cuh: hi32) */
#endif
/* Launch bounds of the direct kernel: (threads, 1) lets ptxas spend registers (12-60 more) instead of chasing resident
 * CTAs; (threads, 0) = no second bound = its occupancy heuristic.  Chosen per (op, length) from an A/B sweep of all
 * units (profiles/r1_e_notes.md): free registers win for the f32 units, cfft_q31 up to N = 2048 (+12 points at 512),
 * rfft_q31 forward up to complex 256 and rfft_q31 inverse; they lose 2-13 points for the 256-thread CTAs of N = 4096
 * and for the q15 real FFT, and are neutral for cfft_q15.  Fused rfft_fast_f64 (profiles/r1_f_notes.md): free registers win
 * for the forward units at every length (+1..22 points: real N = 256 62.9 -> 84.9 %, 1024 60.7 -> 73.9 %) and lose 1-15
 * points for the inverse ones except the shortest. */
#ifndef KU_MINB
#if KU_OP == 10 && (KU_N == 1024 || KU_N == 2048)
/* cfft_f64: 128 registers / 4 CTAs at N = 1024 (92.1 vs 91.1 % of the HBM peak with free registers), 96 / 5 CTAs at
 * N = 2048 (84.9 vs 84.3 %); N = 4096 stays on ptxas' own choice (128 registers, 2 CTAs of 256 threads: a third CTA
 * needs 80 registers, spills, 68.5 -> 48.8 %) */
#define KU_MINB 4        /* round 2: N = 2048 with 128 / 4 as well: 84.0 -> 90.6 % (six CTAs of 80 registers: 55 %; profiles/r2_af_minb.txt) */
#elif KU_OP == 5 && KU_N >= 128 && KU_N <= 2048
/* rfft_q31 forward with the split coefficients prefetched across the barrier (RfftFixFwdBody::kPrefetch): 96 registers / 5
 * CTAs; ptxas' own choice (72 registers) spills the prefetched values (profiles/r2_notes.md: +1..7 points over no prefetch) */
#define KU_MINB 5
#elif KU_OP == 1 && (KU_N == 256 || KU_N == 1024)
/* cfft_q31: ptxas' free choice is 82 registers, two above the step to a sixth resident CTA of 128 threads; 80 / 6 CTAs:
 * N = 256 103.4 -> 106.3 %, N = 1024 85.0 -> 90.9 % of the HBM peak.  The same step for N = 512 (78 -> 72 registers, 7 CTAs)
 * loses 3.6 points and is neutral at 2048 (profiles/r2_x_thresholds.txt) */
#define KU_MINB 6
#elif KU_OP == 2 && KU_N == 256
/* cfft_q15 N = 256: 62 -> 56 registers, 9 CTAs instead of 8: 72.2 -> 75.2 % */
#define KU_MINB 9
#elif (KU_OP == 1 || KU_OP == 2) && KU_N == 4096
/* 256-thread CTAs: a bound of three lets ptxas take 78 (q31: three CTAs instead of four) / 64 (q15: still four) registers and
 * schedule for them: cfft_q31 66.5 -> 68.2 %, cfft_q15 44.8 -> 47.1 %; a bound of two: 60.1 / 47.1 % (profiles/r2_ad_4096_minb.txt) */
#define KU_MINB 3
#elif KU_OP == 2 && KU_N == 512
/* cfft_q15 N = 512: 64 registers / 8 CTAs instead of 54 / 9: 59.4 -> 60.5 % (N = 128 +0.3, N = 2048 -0.1: left alone; profiles/r2_ae_minb.txt) */
#define KU_MINB 8
#elif KU_OP == 12 && KU_N >= 256 && KU_N <= 2048
/* rfft_fast_f64 inverse (KU_N = complex length), launch-bound sweep of round 2 (profiles/r2_ag_f64_minb.txt; % of the HBM peak
 * at 3 / 4 / 5 CTAs aimed at, ptxas' own choice first): 256: 86.3 | 88.5 / 90.8 / 84.9; 512: 68.8 | 80.6 / 74.5 / 77.4;
 * 1024: 65.9 | 67.5 / 66.3 / 73.8; 2048: 62.1 | 63.8 / 64.0 / 64.9 */
#define KU_MINB (KU_N == 256 ? 4 : (KU_N == 512 ? 3 : 5))
#elif KU_OP == 11 && KU_N == 256
#define KU_MINB 4        /* rfft_fast_f64 forward, real N = 512: 93.1 -> 94.2 % */
#elif KU_OP == 5 && KU_N == 4096
/* rfft_q31 forward, real N = 8192 (256-thread CTAs): 64 registers / 4 CTAs instead of 77 / 3: 67.2 -> 67.9 % */
#define KU_MINB 4
#elif KU_OP == 7 && KU_N == 4096
/* rfft_q15 forward, real N = 8192 (256-thread CTAs): 43.2 -> 44.4 % */
#define KU_MINB 3
#elif KU_OP == 2 && KU_N == 1024
/* cfft_q15 N = 1024: 55 -> 48 registers (no spills), 10 CTAs instead of 9: 55.5 -> 56.6 %, twice in a row; the same bound
 * loses 1-6 points at N = 128 / 512 / 2048 (profiles/r2_ac_q15_minb10.txt) */
#define KU_MINB 10
#elif KU_OP == 6 && KU_N >= 128
/* rfft_q31 inverse after the rounding multiply-accumulates moved to IMAD.HI (fft_arith.cuh: rhi32_acc): 96 registers / 5 CTAs
 * (profiles/r2_q_rmac.txt: real N = 256 ... 8192 66 / 69 / 62 / 58 / 54 / 46 % against 60 / 64 / 56 / 52 / 47 / 29 % with free
 * registers (142, 3 CTAs); ptxas' own 56-72 registers / 7-9 CTAs are within half a point of 5) */
#define KU_MINB (KU_N >= 256 && KU_N <= 1024 ? 7 : (KU_N == 4096 ? 3 : 5))      /* 72 registers / 7 CTAs at real N = 512 ... 2048: 69.1 / 63.0 / 60.3 % (profiles/r2_aa_rifft_minb.txt) */
#elif KU_OP == 0 || KU_OP == 3 || KU_OP == 4 || KU_OP == 9 || KU_OP == 6 || (KU_OP == 10 && KU_N <= 512) || KU_OP == 11 || (KU_OP == 12 && KU_N == 16) || (KU_OP == 1 && KU_N <= 2048) || (KU_OP == 5 && KU_N <= 64)
#define KU_MINB 1
#else
#define KU_MINB 0
#endif
#endif

/* frames per CTA of the fixed-point N = 1024 plan (fft_plans.cuh; the twiddle tables do not depend on it): one frame = one CTA of
 * two warps for the q15 complex FFT and the forward q15 real FFT -- the CTA-wide barrier between the passes then waits for one
 * frame only: cfft_q15 56.6 -> 58.7 %, rfft_q15 forward 55.9 -> 57.5 %; q31 and the inverse real FFTs lose 0.3-2.4 points and
 * keep two frames (profiles/r2_ak_fix1024_f1.txt) */
#if !defined(FFT_FIX1024_F) && (KU_OP == 2 || KU_OP == 7) && KU_N == 1024
#define FFT_FIX1024_F 1
#endif

#include "../../../include/cmsisdsp_cuda.h"
#include "fft_plans.cuh"
#include "kernel_entry.h"
#include "bulk_copy.cuh"

/* Resident CTAs per SM the pipelined kernel's register allocation aims at (the second launch bound; it changes ptxas'
 * register budget AND its schedule).  Chosen per (op, length) from an A/B/C sweep of 2 / 3 / 4 over every pipelined unit
 * (profiles/r1_e_notes.md).  Default two: the rfft epilogues want > 200 registers and the plain cfft_f32 is HBM-bound
 * either way.  Three: forward rfft of complex length 2048 -- the bench kernel: same 3 CTAs/SM and 168 instead of 167
 * registers, but 0.380 -> 0.345 ms (86.5 -> 95.2 % of the HBM peak) -- and 1024 (79 -> 84 %), magnitude / peak epilogues at
 * N >= 2048 (+5..8 points).  Four: forward rfft of complex length 256 (89.8 -> 94.0 %), epilogues at N = 1024 (peak
 * 71.5 -> 75.9 %).  Four at N >= 2048 spills and loses 30-40 points. */
#ifndef KU_PIPE_MINB
#if (KU_OP == 3 && KU_N == 256) || (KU_OP == 9 && KU_N == 1024)
#define KU_PIPE_MINB 4
#elif (KU_OP == 3 && KU_N >= 1024) || (KU_OP == 9 && KU_N >= 2048)
#define KU_PIPE_MINB 3
#else
#define KU_PIPE_MINB 2
#endif
#endif

/* Upper bound on the resident CTAs per SM the persistent pipelined kernel is launched with (its grid is occupancy x SMs).
 * Forward rfft of complex length 512: the round-2 addressing rewrite freed registers (134 -> 116), a fourth CTA fits, and
 * the kernel went from 95 to 86 % of the HBM peak -- four CTAs were already measured slower for this unit in round 1
 * (profiles/r1_e_notes.md: 94.0 % at 3 CTAs, 87.9 % at 4).  More CTAs in flight only lengthen every frame's turn at the
 * memory system here; three saturate it. */
/* Round-2 sweep of the cap over every pipelined unit (profiles/r2_j_pipe_occupancy_caps.txt; 2 / 3 / 4 CTAs): the in-place
 * complex FFT of N = 512 and 1024 -- the north_star's first target kernel -- runs at 98 % of the measured HBM peak with TWO
 * resident CTAs (8 warps per SM) against 92 % with the four that fit; inverse rfft of complex length 256 and cfft + magnitude
 * of N = 512 gain 3-6 points at three.  Everything else is best at what fits. */
#ifndef KU_PIPE_MAXOCC
#if KU_OP == 0 && KU_N == 2048
#define KU_PIPE_MAXOCC 1             /* 95.9 % with one CTA (4 frames in flight per SM), 90.8 % with the two that fit */
#elif KU_OP == 0 && (KU_N == 512 || KU_N == 1024)
#define KU_PIPE_MAXOCC 2
#elif (KU_OP == 3 && KU_N == 512) || (KU_OP == 4 && KU_N == 256) || (KU_OP == 9 && KU_N == 512)
#define KU_PIPE_MAXOCC 3
#else
#define KU_PIPE_MAXOCC 64
#endif
#endif

#if !defined(KU_OP) || !defined(KU_N)
#error "compile with -DKU_OP=<0..12> -DKU_N=<length>"
#endif

using namespace b200fft;

/* barrier between two phases of a frame: when a frame's T threads sit inside one warp the
 * exchange is warp-private and __syncwarp() is enough (no CTA-wide stall) */
template <class PL> __device__ __forceinline__ void frame_sync()
{
    if constexpr (PL::T <= 32) __syncwarp();
    else __syncthreads();
}

/* the phases of a body, a frame barrier between two of them */
template <class BODY, class PL, int PH>
__device__ __forceinline__ void run_phases(typename BODY::Regs &r, const typename BODY::Args &a, typename BODY::xelem *sm, int i, bool valid)
{
    if (valid) BODY::template phase<PH>(r, a, sm, i);
    if constexpr (PH + 1 < BODY::kPhases) {
        frame_sync<PL>();
        run_phases<BODY, PL, PH + 1>(r, a, sm, i, valid);
    }
}

/* L2 prefetch distance of the direct kernel in units of "CTAs resident on the whole device" (0: off).  A CTA of the direct
 * kernel starts with its frame's global loads and nothing else to do: ncu shows 12-15 % of the fixed-point kernels' warp
 * time waiting for them (first use of the loaded words).  The frame that the CTA taking over this CTA's slot will want is
 * blockIdx + (resident CTAs) away; every thread asks L2 for one 128-byte line of it (prefetch.global.L2, no register, no
 * scoreboard), so that CTA's loads are L2 hits. */
/* Measured per unit (profiles/r2_ar_prefetch.txt, r2_as_prefetch_b.txt; % of the HBM peak without -> with a distance of one
 * device-load of CTAs): cfft_f64 N = 512 / 1024 / 2048 101.7 / 92.9 / 91.3 -> 104.3 / 99.2 / 94.8, cfft_q31 N = 512 / 1024
 * 96.3 / 90.9 -> 97.6 / 92.4, rfft_q31 forward real N = 512 ... 4096 98.3 / 90.6 / 83.9 / 78.3 -> 102.3 / 92.0 / 85.0 / 78.9,
 * rfft_fast_f64 forward real N = 256 / 2048 +1.9 / +1.5 (within the run-to-run spread), inverse real N = 128 / 256 / 512 / 1024
 * 71.8 / 85.4 / 92.6 / 81.4 -> 81.5 / 98.4 / 97.5 / 83.6 (r2_ax_f64_prefetch.txt).  Neutral or slower (0 ... -2.5 points) for q15 (issue-bound: the 1-2
 * extra instructions per thread cost more than the wait they remove), the inverse real FFTs, the short f32 lengths, and with
 * twice the distance; cfft_f64 N = 4096 loses a resident CTA to two more registers.  Off everywhere else. */
#ifndef KU_PREFETCH
#if (KU_OP == 10 && (KU_N == 512 || KU_N == 1024 || KU_N == 2048)) || (KU_OP == 1 && (KU_N == 512 || KU_N == 1024)) || \
    (KU_OP == 5 && KU_N >= 256 && KU_N <= 2048) || (KU_OP == 11 && (KU_N == 128 || KU_N == 1024)) || (KU_OP == 12 && KU_N >= 64 && KU_N <= 512)
#define KU_PREFETCH 1
#else
#define KU_PREFETCH 0
#endif
#endif

template <class BODY, class PL>
__global__ void __launch_bounds__(PL::kThreads, KU_MINB) frame_kernel(typename BODY::Args base, uint64_t nFrames, uint32_t residentCtas)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename BODY::xelem xelem;                  /* shared-memory exchange element */
    const int tid = threadIdx.x;
    const int fl = tid / PL::T, i = tid % PL::T;
    const uint64_t frame = (uint64_t)blockIdx.x * PL::F + fl;
    const bool valid = frame < nFrames;
    if constexpr (KU_PREFETCH > 0) {
        const uint64_t ahead = frame + (uint64_t)KU_PREFETCH * residentCtas * PL::F;
        if (ahead < nFrames) {
            const char *p = reinterpret_cast<const char *>(BODY::for_frame(base, ahead).in);
            const size_t bytes = (size_t)(reinterpret_cast<const char *>(BODY::for_frame(base, 1).in) - reinterpret_cast<const char *>(base.in));
            for (size_t off = (size_t)i * 128; off < bytes; off += (size_t)PL::T * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
        }
    }
    xelem *sm = reinterpret_cast<xelem *>(smem_raw) + fl * PL::kFrameElems;
    typename BODY::Args a = BODY::for_frame(base, valid ? frame : 0);
    BODY::set_scratch(a, reinterpret_cast<xelem *>(smem_raw) + PL::F * PL::kFrameElems + fl * PL::kSpecial);
    typename BODY::Regs r;

    run_phases<BODY, PL, 0>(r, a, sm, i, valid);
}

/* ------------------------------------------------------------------ persistent TMA-fed kernel
 *
 * For the two-pass f32 plans (32/64 points per thread) a warp owns whole frames, so few warps
 * fit on an SM and a warp that waits for its own global loads leaves HBM idle.  Here the loads
 * are taken off the warps: a CTA (ONE warp, or two for N = 4096) loops over frame groups, and
 * the group's ONE shared-memory buffer is, in turn,
 *   1. the destination of a bulk async copy (cp.async.bulk = the 1-D TMA path, completion on an
 *      mbarrier) of the group's frames, linear as in HBM,
 *   2. once every thread holds its points in registers, the padded exchange buffer,
 *   3. as soon as the exchange has been read back, the TMA destination of the NEXT group --
 *      that copy flies while the second pass, the epilogue and the stores of this group run.
 * Results leave through the registers (coalesced streaming stores). */

/* A CTA of the pipelined kernel holds kUnits independent pipelines ("units": one warp, or the two
 * warps of an N = 4096 frame), 128 threads in all.  They share only the read-only table values:
 * what each thread needs from the twiddle tables does not depend on the frame, so it is fetched
 * once per launch and parked in shared memory [value][thread of unit] -- with the carve-out this
 * kernel needs there is next to no L1 left, and per-frame table loads were seen to miss it and
 * stall every iteration on an L2 round trip (profiles/r1_ncu_hotspots_rfft_fwd_c.txt). */
template <class BODY, class PL> struct PipeSmem {
    typedef typename PL::Arith::elem elem;
    static constexpr int kUnitThreads = PL::kThreads;
    static constexpr int kUnits = 128 / kUnitThreads;
    static constexpr int kCtaThreads = kUnits * kUnitThreads;
    /* per unit: F*kFrameElems >= F*N (linear input, then padded exchange) + the scratch areas */
    static constexpr int kBufBytes = (PL::kSmemElems * (int)sizeof(elem) + 127) & ~127;
    static constexpr int kHoistVals = (int)(sizeof(typename BODY::Hoist) / sizeof(elem));
    static constexpr int kHoistBytes = kHoistVals * kUnitThreads * (int)sizeof(elem);
    static constexpr int kBytes = kUnits * kBufBytes + kHoistBytes + kUnits * 8;      /* + one mbarrier per unit */
    /* Two CTAs (8 warps) per SM and up to 255 registers.  Measured: capping registers to fit three or
     * four CTAs costs 5-30 % (the rfft epilogues want > 200 registers; a spill is ruinous here because
     * with this kernel's shared-memory carve-out next to no L1 is left, so every reload is an L2 round
     * trip), and the extra warps buy nothing once the loads are off the warps' critical path. */
    static constexpr int kMinBlocks = KU_PIPE_MINB;
};

/* barrier among the threads of one unit */
template <class PL> __device__ __forceinline__ void unit_sync(int unit)
{
    if constexpr (PL::kThreads <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(unit + 1), "r"(PL::kThreads) : "memory");
}

#ifndef KU_PIPE_PREFETCH
#define KU_PIPE_PREFETCH 0           /* A/B: L2 prefetch of the group after the one being copied in */
#endif

template <class BODY, class PL>
__global__ void __launch_bounds__(PipeSmem<BODY, PL>::kCtaThreads, PipeSmem<BODY, PL>::kMinBlocks)
frame_kernel_pipe(typename BODY::Args base, uint64_t nFrames)
{
    static_assert(PL::NP == 2, "the pipelined kernel is written for two-pass plans");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename BODY::elem elem;
    typedef PipeSmem<BODY, PL> SM;
    static_assert(sizeof(typename BODY::Hoist) == SM::kHoistVals * sizeof(elem), "Hoist must be a flat set of table values");

    const int unit = threadIdx.x / SM::kUnitThreads, ut = threadIdx.x % SM::kUnitThreads;
    const int fl = ut / PL::T, i = ut % PL::T;
    elem *buf = reinterpret_cast<elem *>(smem_raw + unit * SM::kBufBytes);
    elem *parked = reinterpret_cast<elem *>(smem_raw + SM::kUnits * SM::kBufBytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + SM::kUnits * SM::kBufBytes + SM::kHoistBytes) + unit;

    const uint64_t nGroups = (nFrames + PL::F - 1) / PL::F;
    const uint64_t stride = (uint64_t)gridDim.x * SM::kUnits;
    constexpr uint32_t kGroupElems = PL::F * PL::N;

    if (ut == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (unit == 0) {
        typename BODY::Hoist h0;
        BODY::hoist(h0, base, i);
        const elem *hp = reinterpret_cast<const elem *>(&h0);
#pragma unroll
        for (int s = 0; s < SM::kHoistVals; s++) parked[s * SM::kUnitThreads + ut] = hp[s];
    }
    __syncthreads();

    auto group_bytes = [&](uint64_t g) -> uint32_t {
        const uint64_t left = nFrames - g * PL::F;
        return (uint32_t)((left < (uint64_t)PL::F ? left : (uint64_t)PL::F) * PL::N * sizeof(elem));
    };
    auto fetch = [&](uint64_t g) {          /* one thread: start the copy of group g into the unit's buffer */
        const uint32_t bytes = group_bytes(g);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(buf, base.in + g * (uint64_t)kGroupElems, bytes, bar);
#if KU_PIPE_PREFETCH
        if (g + stride < nGroups)           /* and ask L2 for the group after it */
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base.in + (g + stride) * (uint64_t)kGroupElems), "r"(group_bytes(g + stride)) : "memory");
#endif
    };
    uint64_t g = (uint64_t)blockIdx.x * SM::kUnits + unit;
    if (ut == 0 && g < nGroups) fetch(g);
    uint32_t parity = 0u;
    elem *sm = buf + fl * PL::kFrameElems;                 /* exchange area of this thread's frame */
    for (; g < nGroups; g += stride) {
        const uint64_t frame = g * PL::F + fl;
        const bool valid = frame < nFrames;
        typename BODY::Args a = BODY::for_frame(base, valid ? frame : 0);
        a.in = buf + fl * PL::N;                           /* the staged copy of this frame */
        if constexpr (BODY::kImageOut) a.out = buf + fl * PL::N;   /* ... and, at the end, the image of its result */
        BODY::set_scratch(a, buf + PL::F * PL::kFrameElems + fl * PL::kSpecial);
        typename BODY::Regs r;

        const uint64_t gn = g + stride;
        const elem *pk = parked + ut;                      /* this thread's column of parked table values */
        mbar_wait(bar, parity);
        parity ^= 1u;
        if constexpr (BODY::kHasPre) {
            if (valid) BODY::pre_pk(a, sm, i, pk);
            unit_sync<PL>(unit);
        }
        if (valid) BODY::phase0_in_pk(r, a, sm, i, pk);
        unit_sync<PL>(unit);                               /* all inputs are in registers */
        if (valid) BODY::phase0_out(r, sm, i);
        unit_sync<PL>(unit);
        if (valid) BODY::last_in(r, sm, i);
        unit_sync<PL>(unit);                               /* the exchange has been read back: the buffer is free */
        if constexpr (!BODY::kImageOut) {
            /* results leave through the registers; the next group's copy flies meanwhile */
            if (ut == 0 && gn < nGroups) {
                fence_proxy_async();
                fetch(gn);
            }
            if (valid) BODY::last_out_pk(r, a, i, pk);
            if constexpr (BODY::kHasPost) {
                unit_sync<PL>(unit);
                if (valid) BODY::post_pk(a, sm, i, pk);
            }
        } else {
            /* results are assembled as a frame image in the buffer and leave with one bulk store */
            if (valid) BODY::last_out_pk(r, a, i, pk);
            if constexpr (BODY::kHasPost) {
                unit_sync<PL>(unit);
                if (valid) BODY::post_pk(a, sm, i, pk);
            }
            fence_proxy_async();                           /* image writes (generic proxy) -> visible to the TMA */
            unit_sync<PL>(unit);
            if (ut == 0) {
                bulk_s2g(base.out + g * (uint64_t)kGroupElems, buf, group_bytes(g));
                bulk_wait_read_all();                      /* the store has read the image: the buffer is free again */
                if (gn < nGroups) fetch(gn);
            }
        }
    }
}

/* ------------------------------------------------------------------ one thread per frame (N <= 64)
 *
 * For short frames no thread layout gives coalesced register loads (a thread owns at least a
 * quarter of a 128..512-byte frame), so the TMA does the gather: every lane issues ONE bulk copy
 * of its own frame into a padded shared-memory slot (slot stride = frame bytes + 16, which makes
 * the 16-byte vector accesses of 8 consecutive lanes hit 32 different banks), transforms the frame
 * in registers -- single pass, no exchange -- writes the result back into the slot and sends it
 * home with one bulk store.  No LDG/STG at all on the data path. */
/* bytes a thread-per-frame body brings in / sends home per frame, where its result goes, and how its arguments
 * are pointed at the slot (defaults: a frame of N elements in place of itself) */
template <class BODY, class PL, class = void> struct TinyTraits {
    typedef typename PL::Arith::elem elem;
    static constexpr int kIn = PL::N * (int)sizeof(elem), kOut = kIn;
    static __device__ __forceinline__ void *home(const typename BODY::Args &a) { return a.out; }
    static __device__ __forceinline__ void bind(typename BODY::Args &a, void *slot)
    {
        a.in = reinterpret_cast<const elem *>(slot);
        a.out = reinterpret_cast<elem *>(slot);
    }
};
template <class BODY, class PL> struct TinyTraits<BODY, PL, std::void_t<decltype(BODY::kTinyInBytes)>> {
    static constexpr int kIn = BODY::kTinyInBytes, kOut = BODY::kTinyOutBytes;     /* kOut = 0: the body stores its (tiny) result itself */
    static __device__ __forceinline__ void *home(const typename BODY::Args &a) { return BODY::tiny_home(a); }
    static __device__ __forceinline__ void bind(typename BODY::Args &a, void *slot) { BODY::tiny_bind(a, slot); }
};

template <class BODY, class PL> struct TinySmem {
    typedef TinyTraits<BODY, PL> TT;
    static_assert(TT::kIn % 16 == 0 && TT::kOut % 16 == 0, "bulk copies move multiples of 16 bytes");
    static constexpr int kSlotBytes = (TT::kIn > TT::kOut ? TT::kIn : TT::kOut) + 16;
    /* as many warps per CTA as keep the CTA's slots around 110 KB (two CTAs per SM) */
    static constexpr int kWarps = (kSlotBytes * 128 <= 112 * 1024) ? 4 : ((kSlotBytes * 64 <= 112 * 1024) ? 2 : 1);
    static constexpr int kCtaThreads = 32 * kWarps, kFramesPerCta = kCtaThreads;
    static constexpr int kBytes = kCtaThreads * kSlotBytes + kWarps * 8;
};

/* As KU_PREFETCH, for the thread-per-frame kernel: two to eleven CTAs per SM, each of which starts by waiting for its own
 * bulk copies.  Measured per unit (profiles/r2_aw_tiny_prefetch.txt, % of the HBM peak without -> with): cfft_q31 N = 16
 * 86.3 -> 99.5, cfft_f32 N = 16 87.1 -> 89.3, cfft + magnitude / peak N = 16 77.6 / 72.5 -> 82.3 / 93.5, cfft_q15 N = 32 / 64
 * 62.1 / 51.3 -> 63.6 / 56.9, rfft_q31 forward real N = 64 80.7 -> 85.4, rfft_q15 forward / inverse real N = 64, 128 +1 ... +3.
 * The units that already run at the HBM peak lose 7-28 points with it (f32 N = 32 / 64, every f32 real FFT, cfft_q31 N = 32):
 * the prefetches are a second access stream into DRAM.  Off for those. */
#ifndef KU_TINY_PREFETCH
#if ((KU_OP == 0 || KU_OP == 1 || KU_OP == 9) && KU_N == 16) || (KU_OP == 2 && (KU_N == 32 || KU_N == 64)) || (KU_OP == 5 && KU_N == 32) || \
    ((KU_OP == 7 || KU_OP == 8) && (KU_N == 32 || KU_N == 64))
#define KU_TINY_PREFETCH 1
#else
#define KU_TINY_PREFETCH 0
#endif
#endif

template <class BODY, class PL>
__global__ void __launch_bounds__(TinySmem<BODY, PL>::kCtaThreads) frame_kernel_tiny(typename BODY::Args base, uint64_t nFrames, uint32_t residentCtas)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef TinySmem<BODY, PL> SM;
    typedef typename SM::TT TT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + SM::kCtaThreads * SM::kSlotBytes) + warp;
    void *slot = smem_raw + threadIdx.x * SM::kSlotBytes;
    const uint64_t frame = (uint64_t)blockIdx.x * SM::kFramesPerCta + threadIdx.x;
    const bool valid = frame < nFrames;
    if constexpr (KU_TINY_PREFETCH > 0) {
        const uint64_t ahead = frame + (uint64_t)KU_TINY_PREFETCH * residentCtas * SM::kFramesPerCta;
        if (ahead < nFrames && (TT::kIn >= 128 || (threadIdx.x & 1) == 0)) {      /* 64-byte frames: one request per line */
            const char *p = reinterpret_cast<const char *>(BODY::for_frame(base, ahead).in);
#pragma unroll
            for (int off = 0; off < TT::kIn; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
        }
    }

    if (lane == 0) {
        mbar_init(bar, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    typename BODY::Args a = BODY::for_frame(base, valid ? frame : 0);
    void *home = TT::home(a);
    if (valid) {
        mbar_expect_tx(bar, TT::kIn);                         /* arrive + this lane's bytes */
        bulk_g2s(slot, a.in, TT::kIn, bar);
    } else {
        mbar_arrive(bar);
    }
    mbar_wait(bar, 0);
    if (valid) {
        TT::bind(a, slot);
        typename BODY::Regs r;
        BODY::template phase<0>(r, a, nullptr, 0);
        if constexpr (TT::kOut > 0) {
            fence_proxy_async();                              /* the slot's generic-proxy writes -> visible to the TMA */
            bulk_s2g(home, slot, TT::kOut);
        }
    }
    bulk_wait_read_all();                                     /* shared memory must outlive the store's reads */
}

#define KU_TRY(call)                                                                  \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) return shim_fail(CMSISDSP_CUDA_ERR_RUNTIME, #call, e_); \
    } while (0)

/* kernels that need more than the default 48 KiB of dynamic shared memory opt in once per device */
template <class BODY, class PL> static int prepare()
{
    if (PL::kSmemBytes <= 48 * 1024) return CMSISDSP_CUDA_OK;
    static bool done[64] = {};
    int dev = 0;
    KU_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return shim_fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "device index out of range", cudaSuccess);
    if (!done[dev]) {
        KU_TRY(cudaFuncSetAttribute(frame_kernel<BODY, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PL::kSmemBytes));
        done[dev] = true;
    }
    return CMSISDSP_CUDA_OK;
}

template <class BODY, class PL>
static int launch(const typename BODY::Args &args, uint64_t nFrames, cudaStream_t st)
{
    if (nFrames == 0) return CMSISDSP_CUDA_OK;
    const uint64_t ctas = (nFrames + PL::F - 1) / PL::F;
    if (ctas > 0x7fffffffull) return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "batch too large for one launch", cudaSuccess);
    int rc = prepare<BODY, PL>();
    if (rc) return rc;
    uint32_t resident = 0;
    if (KU_PREFETCH > 0) {
        static std::atomic<uint32_t> cached[64];          /* per device; several host threads may get here at once (same value) */
        int dev = 0;
        KU_TRY(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64) {
            resident = cached[dev].load(std::memory_order_relaxed);
            if (!resident) {
                int occ = 0, sms = 0;
                KU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, frame_kernel<BODY, PL>, PL::kThreads, PL::kSmemBytes));
                KU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
                resident = (uint32_t)(occ > 0 ? occ : 1) * (uint32_t)(sms > 0 ? sms : 1);
                cached[dev].store(resident, std::memory_order_relaxed);
            }
        }
    }
    frame_kernel<BODY, PL><<<(unsigned)ctas, PL::kThreads, PL::kSmemBytes, st>>>(args, nFrames, resident);
    shim_count_launch();
    KU_TRY(cudaGetLastError());
    return CMSISDSP_CUDA_OK;
}

template <class BODY, class PL> static int facts_of(KernelFacts *f)
{
    int rc = prepare<BODY, PL>();
    if (rc) return rc;
    cudaFuncAttributes fa;
    KU_TRY(cudaFuncGetAttributes(&fa, frame_kernel<BODY, PL>));
    int occ = 0;
    KU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, frame_kernel<BODY, PL>, PL::kThreads, PL::kSmemBytes));
    f->threads = PL::kThreads;
    f->frames = PL::F;
    f->smem = PL::kSmemBytes;
    f->regs = fa.numRegs;
    f->ctasPerSm = occ;
    return CMSISDSP_CUDA_OK;
}

static int num_sms()
{
    static int n[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!n[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

/* resident CTAs per SM of the pipelined kernel (also raises its dynamic shared memory limit once) */
template <class BODY, class PL> static int pipe_occupancy(int *occOut)
{
    static int occ[64] = {};
    int dev = 0;
    KU_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return shim_fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "device index out of range", cudaSuccess);
    if (!occ[dev]) {
        KU_TRY(cudaFuncSetAttribute(frame_kernel_pipe<BODY, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PipeSmem<BODY, PL>::kBytes));
        int o = 0;
        KU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, frame_kernel_pipe<BODY, PL>, PipeSmem<BODY, PL>::kCtaThreads, PipeSmem<BODY, PL>::kBytes));
        if (o < 1) return shim_fail(CMSISDSP_CUDA_ERR_RUNTIME, "pipelined kernel does not fit on an SM", cudaSuccess);
        int cap = KU_PIPE_MAXOCC;
        if (const char *e = getenv("CMSISDSP_CUDA_PIPE_MAXOCC")) {          /* A/B runs: cap every pipelined unit */
            if (atoi(e) >= 1) cap = atoi(e);
        }
        occ[dev] = o < cap ? o : cap;
    }
    *occOut = occ[dev];
    return CMSISDSP_CUDA_OK;
}

template <class BODY, class PL>
static int launch_pipe(const typename BODY::Args &args, uint64_t nFrames, cudaStream_t st)
{
    if (nFrames == 0) return CMSISDSP_CUDA_OK;
    int occ = 0;
    int rc = pipe_occupancy<BODY, PL>(&occ);
    if (rc) return rc;
    typedef PipeSmem<BODY, PL> SM;
    const uint64_t groups = (nFrames + PL::F - 1) / PL::F;
    const uint64_t ctas = (groups + SM::kUnits - 1) / SM::kUnits;
    const uint64_t slots = (uint64_t)occ * (uint64_t)num_sms();
    const unsigned grid = (unsigned)(ctas < slots ? ctas : slots);
    frame_kernel_pipe<BODY, PL><<<grid, SM::kCtaThreads, SM::kBytes, st>>>(args, nFrames);
    shim_count_launch();
    KU_TRY(cudaGetLastError());
    return CMSISDSP_CUDA_OK;
}

template <class BODY, class PL> static int tiny_prepare(int *occOut)
{
    static int occ[64] = {};
    int dev = 0;
    KU_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return shim_fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "device index out of range", cudaSuccess);
    if (!occ[dev]) {
        KU_TRY(cudaFuncSetAttribute(frame_kernel_tiny<BODY, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TinySmem<BODY, PL>::kBytes));
        int o = 0;
        KU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, frame_kernel_tiny<BODY, PL>, TinySmem<BODY, PL>::kCtaThreads, TinySmem<BODY, PL>::kBytes));
        if (o < 1) return shim_fail(CMSISDSP_CUDA_ERR_RUNTIME, "thread-per-frame kernel does not fit on an SM", cudaSuccess);
        occ[dev] = o;
    }
    if (occOut) *occOut = occ[dev];
    return CMSISDSP_CUDA_OK;
}
template <class BODY, class PL>
static int launch_tiny(const typename BODY::Args &args, uint64_t nFrames, cudaStream_t st)
{
    if (nFrames == 0) return CMSISDSP_CUDA_OK;
    int occ = 0;
    int rc = tiny_prepare<BODY, PL>(&occ);
    if (rc) return rc;
    const uint64_t ctas = (nFrames + TinySmem<BODY, PL>::kFramesPerCta - 1) / TinySmem<BODY, PL>::kFramesPerCta;
    if (ctas > 0x7fffffffull) return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "batch too large for one launch", cudaSuccess);
    frame_kernel_tiny<BODY, PL><<<(unsigned)ctas, TinySmem<BODY, PL>::kCtaThreads, TinySmem<BODY, PL>::kBytes, st>>>(args, nFrames,
                                                                                                                       (uint32_t)(occ * num_sms()));
    shim_count_launch();
    KU_TRY(cudaGetLastError());
    return CMSISDSP_CUDA_OK;
}
template <class BODY, class PL> static int facts_of_tiny(KernelFacts *f)
{
    int occ = 0;
    int rc = tiny_prepare<BODY, PL>(&occ);
    if (rc) return rc;
    cudaFuncAttributes fa;
    KU_TRY(cudaFuncGetAttributes(&fa, frame_kernel_tiny<BODY, PL>));
    f->threads = TinySmem<BODY, PL>::kCtaThreads;
    f->frames = TinySmem<BODY, PL>::kFramesPerCta;
    f->smem = TinySmem<BODY, PL>::kBytes;
    f->regs = fa.numRegs;
    f->ctasPerSm = occ;
    return CMSISDSP_CUDA_OK;
}

template <class BODY, class PL> static int facts_of_pipe(KernelFacts *f)
{
    int occ = 0;
    int rc = pipe_occupancy<BODY, PL>(&occ);
    if (rc) return rc;
    cudaFuncAttributes fa;
    KU_TRY(cudaFuncGetAttributes(&fa, frame_kernel_pipe<BODY, PL>));
    f->threads = PipeSmem<BODY, PL>::kCtaThreads;
    f->frames = PL::F * PipeSmem<BODY, PL>::kUnits;
    f->smem = PipeSmem<BODY, PL>::kBytes;
    f->regs = fa.numRegs;
    f->ctasPerSm = occ;
    return CMSISDSP_CUDA_OK;
}

template <class PL> static size_t twiddles_of(const void *base, void *hostOut)
{
    typedef typename PL::Arith::elem elem;
    typedef typename PL::Arith::telem telem;
    if (hostOut) PL::build_twiddles((const elem *)base, (telem *)hostOut);
    return (size_t)PL::kTwEntries + 1;
}

/* ------------------------------------------------------------------ the pair of this unit */

/* The TMA-fed flavour (KF_PIPE) of a unit is the persistent pipelined kernel for the two-pass f32
 * plans (its CTA unit is one warp, T <= 32, or one frame) and the thread-per-frame kernel for the
 * single-pass plans (N <= 64). */
template <class P> struct PipeOf {
    static constexpr bool kTiny = (P::NP == 1) && (P::T == 1);
    static constexpr bool kPipe = (P::NP == 2) && IsF32<typename P::Arith::elem>::value && (P::E >= 32 || (KU_OP >= 3 && P::kSpecial > 0));
    static constexpr bool kHas = kTiny || kPipe;
    /* default flavour of this unit, from the A/B sweeps in profiles/ (CMSISDSP_CUDA_KERNEL overrides):
     * the TMA-fed kernels win wherever they exist, except the inverse rfft at complex length 512 */
    static constexpr bool kPrefer = kHas && !(KU_OP == 4 && KU_N == 512);
    typedef typename P::template with_frames<(P::T >= 32 ? 1 : 32 / P::T)> type;
};
static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }   /* bulk copies need 16-byte aligned sources */

#if KU_OP <= 2 || KU_OP == 10   /* complex FFT, in place */

#if KU_OP == 10
typedef PlanCfftF64<KU_N>::type PL;
typedef PL::Arith AR;
#elif KU_OP == 0
typedef ArithF32 AR;
typedef PlanCfftF32<KU_N>::type PL;
#elif KU_OP == 1
typedef ArithQ31 AR;
typedef PlanCfftFix<ArithQ31, KU_N>::type PL;
#else
typedef ArithQ15 AR;
typedef PlanCfftFix<ArithQ15, KU_N>::type PL;
#endif
typedef PipeOf<PL> PIPE;

template <bool INV, bool PERM>
static int cfft_go(const void *in, void *out, uint64_t nFrames, const void *tw, const void *aux, int shl1, int flavour, cudaStream_t st)
{
    typedef AR::elem elem;
    typedef AR::telem telem;
    if constexpr (PIPE::kTiny) {
        if (flavour == KF_PIPE && aligned16(in)) {
            typedef TinyCfftBody<PL, INV, PERM> BODY;
            typename BODY::Args a{(const elem *)in, (elem *)out, (const telem *)tw, (const uint16_t *)aux, 1.0f / (float)PL::N, shl1};
            return launch_tiny<BODY, PL>(a, nFrames, st);
        }
    }
    if constexpr (PIPE::kPipe) {
        if (flavour == KF_PIPE && aligned16(in)) {
            typedef CfftBody<PIPE::type, INV, PERM, true> BODY;
            typename BODY::Args a{(const elem *)in, (elem *)out, (const telem *)tw, (const uint16_t *)aux, 1.0f / (float)PL::N, shl1};
            return launch_pipe<BODY, PIPE::type>(a, nFrames, st);
        }
    }
    typedef CfftBody<PL, INV, PERM> BODY;
    typename BODY::Args a{(const elem *)in, (elem *)out, (const telem *)tw, (const uint16_t *)aux, 1.0f / (float)PL::N, shl1};
    return launch<BODY, PL>(a, nFrames, st);
}
#if KU_OP == 0
/* arm_cfft_f32 with a window multiply fused into the load (aux2 = N window values): the direct kernel, natural order */
template <bool INV>
static int cfft_win_go(const void *in, void *out, uint64_t nFrames, const void *tw, const void *win, cudaStream_t st)
{
    typedef CfftBody<PL, INV, false, false, false, true> BODY;
    typename BODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, nullptr, 1.0f / (float)PL::N, 0, nullptr, (const float *)win};
    return launch<BODY, PL>(a, nFrames, st);
}
#endif
static int ku_launch(const void *in, void *out, uint64_t nFrames, int inv, const void *tw, const void *aux, const void *aux2, int shl1, int flavour, cudaStream_t st)
{
#if KU_OP == 0
    if (aux2) {
        if (aux) return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "windowed cfft_f32: natural-order output only", cudaSuccess);
        return inv ? cfft_win_go<true>(in, out, nFrames, tw, aux2, st) : cfft_win_go<false>(in, out, nFrames, tw, aux2, st);
    }
#else
    (void)aux2;
#endif
    if (inv) return aux ? cfft_go<true, true>(in, out, nFrames, tw, aux, shl1, flavour, st) : cfft_go<true, false>(in, out, nFrames, tw, aux, shl1, flavour, st);
    return aux ? cfft_go<false, true>(in, out, nFrames, tw, aux, shl1, flavour, st) : cfft_go<false, false>(in, out, nFrames, tw, aux, shl1, flavour, st);
}
static int ku_facts(KernelFacts *f, int flavour)
{
    if constexpr (PIPE::kTiny) {
        if (flavour == KF_PIPE) return facts_of_tiny<TinyCfftBody<PL, false, false>, PL>(f);
    }
    if constexpr (PIPE::kPipe) {
        if (flavour == KF_PIPE) return facts_of_pipe<CfftBody<PIPE::type, false, false, true>, PIPE::type>(f);
    }
    return facts_of<CfftBody<PL, false>, PL>(f);
}
typedef PL TWPLAN;                 /* plan whose pass-ordered twiddle table the shim uploads */

#elif KU_OP == 3   /* arm_rfft_fast_f32 forward, KU_N = complex length */

typedef PlanRfftFwd<KU_N>::type PL;
#if KU_N <= 64
typedef PlanCfftF32<KU_N>::type TPL;   /* the single-pass complex plan of the thread-per-frame kernel (needs no twiddle table) */
struct PIPE { static constexpr bool kHas = true, kPrefer = true; };
#else
typedef PipeOf<PL> PIPE;
#endif
static int ku_launch(const void *in, void *out, uint64_t nFrames, int, const void *tw, const void *aux, const void *aux2, int, int flavour, cudaStream_t st)
{
    if (aux2) {           /* window multiply fused into the load (aux2 = fftLenReal window values): the direct kernel */
        typedef RfftFwdBody<PL, false, true> WBODY;
        WBODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, (const cf32 *)aux, nullptr, (const cf32 *)aux2};
        return launch<WBODY, PL>(a, nFrames, st);
    }
#if KU_N <= 64
    if (flavour == KF_PIPE && aligned16(in) && aligned16(out)) {
        typedef TinyRfftFwdBody<TPL> BODY;
        BODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, (const cf32 *)aux};
        return launch_tiny<BODY, TPL>(a, nFrames, st);
    }
#else
    if constexpr (PIPE::kPipe) {
        /* kImageOut: the spectrum leaves with a bulk store, which needs a 16-byte aligned destination too */
        if (flavour == KF_PIPE && aligned16(in) && aligned16(out)) {
            typedef RfftFwdBody<PIPE::type, true> BODY;
            BODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, (const cf32 *)aux};
            return launch_pipe<BODY, PIPE::type>(a, nFrames, st);
        }
    }
#endif
    typedef RfftFwdBody<PL> BODY;
    BODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, (const cf32 *)aux};
    return launch<BODY, PL>(a, nFrames, st);
}
static int ku_facts(KernelFacts *f, int flavour)
{
#if KU_N <= 64
    if (flavour == KF_PIPE) return facts_of_tiny<TinyRfftFwdBody<TPL>, TPL>(f);
#else
    if constexpr (PIPE::kPipe) {
        if (flavour == KF_PIPE) return facts_of_pipe<RfftFwdBody<PIPE::type, true>, PIPE::type>(f);
    }
#endif
    return facts_of<RfftFwdBody<PL>, PL>(f);
}
typedef PL TWPLAN;

#elif KU_OP == 4   /* arm_rfft_fast_f32 inverse */

typedef PlanRfftInv<KU_N>::type PL;
#if KU_N <= 64
typedef PlanCfftF32<KU_N>::type TPL;
struct PIPE { static constexpr bool kHas = true, kPrefer = true; };
#else
typedef PipeOf<PL> PIPE;
#endif
static int ku_launch(const void *in, void *out, uint64_t nFrames, int, const void *tw, const void *aux, const void *, int, int flavour, cudaStream_t st)
{
#if KU_N <= 64
    if (flavour == KF_PIPE && aligned16(in) && aligned16(out)) {
        typedef TinyRfftInvBody<TPL> BODY;
        BODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, (const cf32 *)aux, 1.0f / (float)PL::N};
        return launch_tiny<BODY, TPL>(a, nFrames, st);
    }
#else
    if constexpr (PIPE::kPipe) {
        if (flavour == KF_PIPE && aligned16(in)) {
            typedef RfftInvBody<PIPE::type, true> BODY;
            BODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, (const cf32 *)aux, 1.0f / (float)PL::N};
            return launch_pipe<BODY, PIPE::type>(a, nFrames, st);
        }
    }
#endif
    typedef RfftInvBody<PL> BODY;
    BODY::Args a{(const cf32 *)in, (cf32 *)out, (const cf32 *)tw, (const cf32 *)aux, 1.0f / (float)PL::N};
    return launch<BODY, PL>(a, nFrames, st);
}
static int ku_facts(KernelFacts *f, int flavour)
{
#if KU_N <= 64
    if (flavour == KF_PIPE) return facts_of_tiny<TinyRfftInvBody<TPL>, TPL>(f);
#else
    if constexpr (PIPE::kPipe) {
        if (flavour == KF_PIPE) return facts_of_pipe<RfftInvBody<PIPE::type, true>, PIPE::type>(f);
    }
#endif
    return facts_of<RfftInvBody<PL>, PL>(f);
}
typedef PL TWPLAN;

#elif KU_OP == 9   /* arm_cfft_f32 with a fused spectrum epilogue: out = magnitudes (shl1 = SpectrumMode 0 / 1) or
                    * out = peak values, aux = peak indices (shl1 = SPEC_PEAK) */

typedef PlanCfftF32<KU_N>::type PL;
typedef PipeOf<PL> PIPEOF;
struct PIPE { static constexpr bool kHas = PIPEOF::kHas, kPrefer = PIPEOF::kHas; };

template <bool INV, int MODE>
static int mag_go(const void *in, void *out, uint64_t nFrames, const void *tw, const void *aux, int flavour, cudaStream_t st)
{
    if constexpr (PIPEOF::kTiny) {
        if (flavour == KF_PIPE && aligned16(in) && (MODE == SPEC_PEAK || aligned16(out))) {
            typedef TinyCfftMagBody<PL, INV, MODE> BODY;
            typename BODY::Args a{};
            a.in = (const cf32 *)in; a.tw = (const cf32 *)tw; a.scale = 1.0f / (float)PL::N;
            a.mag = (float *)out; a.peakVal = (float *)out; a.peakIdx = (uint32_t *)aux;
            return launch_tiny<BODY, PL>(a, nFrames, st);
        }
    }
    if constexpr (PIPEOF::kPipe) {
        typedef CfftMagBody<PIPEOF::type, INV, MODE, true> BODY;
        if constexpr (!BODY::kCross) {          /* a frame wider than a warp meets through the exchange buffer: direct kernel */
            if (flavour == KF_PIPE && aligned16(in)) {
                typename BODY::Args a{};
                a.in = (const cf32 *)in; a.tw = (const cf32 *)tw; a.scale = 1.0f / (float)PL::N;
                a.mag = (float *)out; a.peakVal = (float *)out; a.peakIdx = (uint32_t *)aux;
                return launch_pipe<BODY, PIPEOF::type>(a, nFrames, st);
            }
        }
    }
    typedef CfftMagBody<PL, INV, MODE> BODY;
    typename BODY::Args a{};
    a.in = (const cf32 *)in; a.tw = (const cf32 *)tw; a.scale = 1.0f / (float)PL::N;
    a.mag = (float *)out; a.peakVal = (float *)out; a.peakIdx = (uint32_t *)aux;
    return launch<BODY, PL>(a, nFrames, st);
}
static int ku_launch(const void *in, void *out, uint64_t nFrames, int inv, const void *tw, const void *aux, const void *, int mode, int flavour, cudaStream_t st)
{
    switch (mode) {
    case SPEC_MAG: return inv ? mag_go<true, SPEC_MAG>(in, out, nFrames, tw, aux, flavour, st) : mag_go<false, SPEC_MAG>(in, out, nFrames, tw, aux, flavour, st);
    case SPEC_MAG_SQUARED: return inv ? mag_go<true, SPEC_MAG_SQUARED>(in, out, nFrames, tw, aux, flavour, st) : mag_go<false, SPEC_MAG_SQUARED>(in, out, nFrames, tw, aux, flavour, st);
    case SPEC_PEAK: return inv ? mag_go<true, SPEC_PEAK>(in, out, nFrames, tw, aux, flavour, st) : mag_go<false, SPEC_PEAK>(in, out, nFrames, tw, aux, flavour, st);
    }
    return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "spectrum mode must be 0 (mag), 1 (mag squared) or 2 (peak)", cudaSuccess);
}
static int ku_facts(KernelFacts *f, int flavour)
{
    if constexpr (PIPEOF::kTiny) {
        if (flavour == KF_PIPE) return facts_of_tiny<TinyCfftMagBody<PL, false, SPEC_MAG>, PL>(f);
    }
    if constexpr (PIPEOF::kPipe) {
        if (flavour == KF_PIPE) return facts_of_pipe<CfftMagBody<PIPEOF::type, false, SPEC_MAG, true>, PIPEOF::type>(f);
    }
    return facts_of<CfftMagBody<PL, false, SPEC_MAG>, PL>(f);
}
typedef PL TWPLAN;

#elif KU_OP == 11 || KU_OP == 12   /* arm_rfft_fast_f64 forward (11) / inverse (12), KU_N = complex length = fftLenRFFT / 2 */

typedef PlanCfftF64<KU_N>::type PL;
struct PIPE { static constexpr bool kHas = false, kPrefer = false; };
#if KU_OP == 11
typedef RfftF64FwdBody<PL> BODY;
#else
typedef RfftF64InvBody<PL> BODY;
#endif
/* in -> out (never aliased), tw = the twiddles of the KU_N-point f64 CFFT plan, aux = twiddleCoefF64_rfft (device) */
static int ku_launch(const void *in, void *out, uint64_t nFrames, int, const void *tw, const void *aux, const void *, int, int, cudaStream_t st)
{
#if KU_OP == 11
    BODY::Args a{(const cf64 *)in, (cf64 *)out, (const cf64 *)tw, (const cf64 *)aux};
#else
    BODY::Args a{(const cf64 *)in, (cf64 *)out, (const cf64 *)tw, (const cf64 *)aux, 1.0f / (float)PL::N};
#endif
    return launch<BODY, PL>(a, nFrames, st);
}
static int ku_facts(KernelFacts *f, int) { return facts_of<BODY, PL>(f); }
typedef PL TWPLAN;

#else              /* arm_rfft_q31 (5 forward, 6 inverse) / arm_rfft_q15 (7, 8); KU_N = complex length = fftLenReal / 2 */

#if KU_OP <= 6
typedef ArithQ31 AR;
#else
typedef ArithQ15 AR;
#endif
typedef PlanCfftFix<AR, KU_N>::type PL;
/* the TMA-fed flavour of these units is the thread-per-frame kernel (complex length <= 64) */
struct PIPE { static constexpr bool kHas = (PL::NP == 1 && PL::T == 1), kPrefer = kHas; };
#if KU_OP == 5 || KU_OP == 7
static constexpr bool kKuInverse = false;
template <bool PERM> struct FixBody { typedef RfftFixFwdBody<PL, PERM> type; };
template <bool PERM> static typename FixBody<PERM>::type::Args ku_args(const void *in, void *out, const void *tw, const void *aux, const void *aux2, int shl1)
{
    return typename FixBody<PERM>::type::Args{(const AR::elem *)in, (AR::elem *)out, (const AR::telem *)tw, (const ci32x4 *)aux, shl1, (const uint16_t *)aux2};
}
#else
static constexpr bool kKuInverse = true;
template <bool PERM> struct FixBody { typedef CfftBody<PL, true, PERM, false, true> type; };
template <bool PERM> static typename FixBody<PERM>::type::Args ku_args(const void *in, void *out, const void *tw, const void *aux, const void *aux2, int shl1)
{
    return typename FixBody<PERM>::type::Args{(const AR::elem *)in, (AR::elem *)out, (const AR::telem *)tw, (const uint16_t *)aux2, 0.0f, shl1, (const ci32x4 *)aux};
}
#endif
typedef FixBody<false>::type BODY;
/* in -> out (never aliased), tw = the pass-ordered twiddles of the KU_N-point CFFT plan, aux = split coefficients,
 * aux2 = the unordered layout of the complex transform when bitReverseFlagR == 0 (else null) */
/* (templates on the direction only so that the thread-per-frame body is not instantiated for the longer plans) */
template <bool KI, bool PERM>
static int fix_go2(const void *in, void *out, uint64_t nFrames, const void *tw, const void *aux, const void *aux2, int shl1, int flavour, cudaStream_t st)
{
    if constexpr (PIPE::kHas) {
        if (flavour == KF_PIPE && aligned16(in) && aligned16(out)) {
            typedef TinyRfftFixBody<PL, KI, PERM> TBODY;
            return launch_tiny<TBODY, PL>(typename TBODY::Args{(const AR::elem *)in, (AR::elem *)out, (const AR::telem *)tw, (const ci32x4 *)aux, shl1}, nFrames, st);
        }
    }
    return launch<typename FixBody<PERM>::type, PL>(ku_args<PERM>(in, out, tw, aux, aux2, shl1), nFrames, st);
}
template <bool KI>
static int fix_go(const void *in, void *out, uint64_t nFrames, const void *tw, const void *aux, const void *aux2, int shl1, int flavour, cudaStream_t st)
{
    return aux2 ? fix_go2<KI, true>(in, out, nFrames, tw, aux, aux2, shl1, flavour, st) : fix_go2<KI, false>(in, out, nFrames, tw, aux, aux2, shl1, flavour, st);
}
template <bool KI> static int fix_facts(KernelFacts *f, int flavour)
{
    if constexpr (PIPE::kHas) {
        if (flavour == KF_PIPE) return facts_of_tiny<TinyRfftFixBody<PL, KI>, PL>(f);
    }
    return facts_of<BODY, PL>(f);
}
static int ku_launch(const void *in, void *out, uint64_t nFrames, int, const void *tw, const void *aux, const void *aux2, int shl1, int flavour, cudaStream_t st)
{
    return fix_go<kKuInverse>(in, out, nFrames, tw, aux, aux2, shl1, flavour, st);
}
static int ku_facts(KernelFacts *f, int flavour) { return fix_facts<kKuInverse>(f, flavour); }
typedef PL TWPLAN;

#endif

#define KU_CAT3(a, b, c) a##b##_##c
#define KU_NAME(op, n) KU_CAT3(ku_entry_, op, n)
namespace b200fft {
extern const KernelEntry KU_NAME(KU_OP, KU_N);
const KernelEntry KU_NAME(KU_OP, KU_N) = {ku_launch, twiddles_of<TWPLAN>, sizeof(TWPLAN::Arith::telem), ku_facts, PIPE::kHas, PIPE::kPrefer};
}
