/*
 * mfcc_unit.cu -- arm_mfcc_f32 front end fused into ONE kernel per frame batch
 * (reference: Source/TransformFunctions/arm_mfcc_f32.c:88-174, arm_mfcc_init_f32.c:91-121):
 *
 *   absmax-normalise -> window -> arm_rfft_fast_f32 (forward) -> |.| (Nyquist dropped) -> * max
 *   -> mel filter bank (packed taps) -> + 1e-6 -> log -> DCT matrix
 *
 * A frame costs one HBM read of fftLen floats and one write of nbDctOutputs floats; everything in
 * between lives in registers and in the frame's shared-memory buffer, which is in turn the staging
 * area of the windowed samples, the exchange buffer of the two-pass FFT, the packed spectrum, the
 * magnitudes.  The FFT is the forward rfft body of fft_body.cuh, unchanged (STAGED = true: it
 * reads its input from and assembles its output in shared memory).  Frames may overlap (hop <
 * fftLen): frame f starts `stride` floats after frame f-1.
 *
 * Two kernels:
 *   mfcc_kernel_pipe  (default; source 16-byte aligned and hop a multiple of 4 floats) -- persistent
 *       CTAs of four warps; a warp owns 32/T frames at a time, each fetched by ONE bulk async copy
 *       (TMA) into the warp's buffer, the next group's copies flying while this group is being
 *       transformed.  The spectrum is never materialised: the rfft split stage hands every bin to a
 *       sink that stores its magnitude; the mel filters are evaluated tap-interleaved over the T
 *       lanes of a frame (conflict-free, balanced to one tap per filter).  All coefficient tables
 *       (window, taps, DCT, parked twiddles) live in shared memory.
 *   mfcc_kernel       (any even hop / 8-byte aligned source) -- one CTA per frame group, direct loads.
 */
#include <cuda_runtime.h>
#include <stdlib.h>

#include <vector>

#include "../../../include/cmsisdsp_cuda.h"
#include "fft_plans.cuh"
#include "kernel_entry.h"
#include "bulk_copy.cuh"

using namespace b200fft;

/* Mel schedule of the pipelined kernel, passed by value (constant bank, indexed by the warp-uniform
 * row counter): row r = 2T consecutive bins starting at bin (rows[r] & 0xffff) against the 2T
 * zero-padded coefficients coefsP[r*2T ..]; (rows[r] >> 16) = filter whose sum is complete after
 * this row, 0xffff = none. */
constexpr int kMfccMaxRows = 128;
struct MfccRows { uint32_t r[kMfccMaxRows]; };

struct MfccDev {
    uint32_t fftLen, nbMel, nbDct;
    float *dct, *coefs, *window;          /* device copies of the caller's coefficient arrays */
    uint32_t *pos, *len, *off;            /* off[f] = start of filter f in coefs */
    const cf32 *tw, *twr;                 /* tables of the rfft plan (owned by the plan cache) */
    int device;
    /* pipelined kernel: filters as {pos, len, off, 0}, DCT matrix transposed to [filter][output] */
    float *dctT, *coefsP;
    uint32_t nRows;
    MfccRows rows;
    int pipeSmem, pipeOcc;                /* dynamic shared memory of the pipelined kernel, resident CTAs per SM (0: does not fit) */
};

struct MfccArgs {
    const float *src;
    uint64_t stride;                      /* floats between frame starts */
    float *dst;
    const float *dct, *coefs, *window;
    const uint32_t *pos, *len, *off;
    const cf32 *tw, *twr;
    uint32_t nbMel, nbDct;
    const float *dctT, *coefsP;           /* pipelined kernel: DCT matrix transposed, zero-padded tap rows */
    uint32_t nRows;
};


template <int T> __device__ __forceinline__ float group_max(float v)
{
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int NC>
__global__ void __launch_bounds__(128) mfcc_kernel(MfccArgs a, uint64_t nFrames)   /* every forward rfft plan has <= 128 threads per CTA */
{
    typedef typename PlanRfftFwd<NC>::type PL;
    typedef RfftFwdBody<PL, true> BODY;
    constexpr int T = PL::T, E = PL::E, F = PL::F;
    static_assert(T <= 32 && (32 % T) == 0, "a frame's threads sit inside one warp");
    static_assert(PL::kThreads <= 128, "launch bounds");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cf32 *smem = reinterpret_cast<cf32 *>(smem_raw);
    float *melAll = reinterpret_cast<float *>(smem + PL::kSmemElems);

    const int tid = threadIdx.x, fl = tid / T, i = tid % T;
    const uint64_t frame = (uint64_t)blockIdx.x * F + fl;
    const bool valid = frame < nFrames;
    cf32 *buf = smem + fl * PL::kFrameElems;                 /* staging -> exchange -> spectrum -> magnitudes */
    float *mel = melAll + fl * a.nbMel;

    /* 1. load, |max|, normalise, window (arm_mfcc_f32.c:104-112) */
    const float2 *src = reinterpret_cast<const float2 *>(a.src + (valid ? frame : 0) * a.stride);
    const float2 *win = reinterpret_cast<const float2 *>(a.window);
    float2 v[E];
    float mx = 0.0f;
#pragma unroll
    for (int m = 0; m < E; m++) {
        v[m] = valid ? __ldcs(src + i + T * m) : make_float2(0.f, 0.f);
        mx = fmaxf(mx, fmaxf(fabsf(v[m].x), fabsf(v[m].y)));
    }
    mx = group_max<T>(mx);
    const float inv = (mx != 0.0f) ? __fdiv_rn(1.0f, mx) : 1.0f;
#pragma unroll
    for (int m = 0; m < E; m++) {
        const float2 w = win[i + T * m];
        buf[i + T * m] = cf32{__fmul_rn(__fmul_rn(v[m].x, inv), w.x), __fmul_rn(__fmul_rn(v[m].y, inv), w.y)};
    }
    __syncwarp();      /* a frame is private to its T lanes of one warp */

    /* 2. forward real FFT, shared memory to shared memory (arm_mfcc_f32.c:137) */
    typename BODY::Args ra{buf, buf, a.tw, a.twr, smem + F * PL::kFrameElems + fl * PL::kSpecial};
    {
        typename BODY::Regs r;
        BODY::phase0_in(r, ra, buf, i);
        __syncwarp();      /* a frame is private to its T lanes of one warp */
        BODY::phase0_out(r, buf, i);
        __syncwarp();      /* a frame is private to its T lanes of one warp */
        BODY::last_in(r, buf, i);
        __syncwarp();      /* a frame is private to its T lanes of one warp */
        BODY::last_out(r, ra, i);                            /* regular bins -> packed spectrum in buf */
    }
    __syncwarp();      /* a frame is private to its T lanes of one warp */
    BODY::post(ra, buf, i);                                  /* the 2R special bins */
    __syncwarp();      /* a frame is private to its T lanes of one warp */

    /* 3. magnitudes of bins 0..NC-1, Nyquist (packed into bin 0) dropped, times max (:138-146) */
    float mag[E];
#pragma unroll
    for (int m = 0; m < E; m++) {
        const int c = i + T * m;
        cf32 z = buf[c];
        if (c == 0) z.y = 0.0f;
        const float s = sqrtf(__fadd_rn(__fmul_rn(z.x, z.x), __fmul_rn(z.y, z.y)));
        mag[m] = (mx != 0.0f) ? __fmul_rn(s, mx) : s;
    }
    __syncwarp();      /* a frame is private to its T lanes of one warp */
    float *magbuf = reinterpret_cast<float *>(buf);
#pragma unroll
    for (int m = 0; m < E; m++) magbuf[i + T * m] = mag[m];
    __syncwarp();      /* a frame is private to its T lanes of one warp */

    /* 4. mel filter bank, + 1e-6, log (:150-165): a lane owns whole filters and sums their taps in the
     * reference's order */
    for (uint32_t f = i; f < a.nbMel; f += T) {
        const uint32_t n = a.len[f];
        const float *c = a.coefs + a.off[f], *mg = magbuf + a.pos[f];
        float s = 0.0f;
        for (uint32_t t = 0; t < n; t++) s = __fadd_rn(s, __fmul_rn(mg[t], c[t]));
        mel[f] = logf(__fadd_rn(s, 1.0e-6f));
    }
    __syncwarp();

    /* 5. DCT matrix (:167-171) */
    if (valid) {
        for (uint32_t r = i; r < a.nbDct; r += T) {
            const float *row = a.dct + r * a.nbMel;
            float s = 0.0f;
            for (uint32_t f = 0; f < a.nbMel; f++) s = __fadd_rn(s, __fmul_rn(row[f], mel[f]));
            a.dst[frame * a.nbDct + r] = s;
        }
    }
}

/* ------------------------------------------------------------------ persistent TMA-fed kernel */

__device__ __forceinline__ float sqrt_fast(float x)
{
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));      /* max relative error 2^-23 */
    return y;
}

/* the rfft split stage hands every finished bin here: only |X[k]| is kept (arm_mfcc_f32.c:138-146;
 * bin 0 carries (DC, Nyquist) and the Nyquist part is dropped) */
struct MagSink {
    float *mag;
    __device__ __forceinline__ void put(int k, cf32 v, bool pred) const
    {
        if (pred) mag[k] = sqrt_fast(fmaf(v.x, v.x, v.y * v.y));
    }
    __device__ __forceinline__ void put_dc_nyquist(cf32 v) const { mag[0] = fabsf(v.x); }
};

template <int NC> struct MfccPipe {
    typedef typename PlanRfftFwd<NC>::type P0;
    typedef typename P0::template with_frames<(P0::T >= 32 ? 1 : 32 / P0::T)> PL;   /* one warp = one unit */
    typedef RfftFwdBody<PL, true> BODY;
    typedef Engine<PL> Eng;
    static constexpr int T = PL::T, F = PL::F, E = PL::E;
#ifndef MFCC_UNITS
#define MFCC_UNITS 4
#endif
    static constexpr int kUnits = MFCC_UNITS, kCtaThreads = 32 * kUnits;
    static_assert(PL::kThreads == 32, "a unit of the pipelined MFCC kernel is one warp");
    static_assert(PL::kFrameElems % 2 == 0, "frame slots must be 16-byte aligned (bulk copy destination)");
    static constexpr int kFrameBytes = 2 * NC * (int)sizeof(float);
    static constexpr int kBufBytes = (PL::kSmemElems * (int)sizeof(cf32) + 127) & ~127;    /* F frame slots + F scratch areas */
    static constexpr int kRowTaps = 2 * T;                                                 /* a mel row: one float2 of bins per lane */
    /* magnitudes take over the frame's slot once the exchange has been read back (NC floats + the 2T
     * the last row of a filter may touch, finite leftovers times zero coefficients), shifted by fl*T
     * floats so that the frames of a warp land in different banks */
    static_assert(2 * PL::kFrameElems >= NC + kRowTaps + (F - 1) * T, "magnitudes must fit in the frame slot");
#ifndef MFCC_QUAD
#define MFCC_QUAD 2                  /* measured 4 / 2 / 1: 31.5 / 32.1 / 31.6 % at fftLen 1024 (profiles/r2_k_mfcc_quad.txt) */
#endif
    static constexpr int kQuad = (T >= MFCC_QUAD) ? MFCC_QUAD : 1;                         /* lanes summed by shuffles before a partial sum is stored */
    static constexpr int kPartStride = T / kQuad + 1;                                      /* per-filter partial sums, conflict-free both ways */
    /* resident CTAs per SM the register allocation aims at (E <= 32: 4 CTAs = 16 warps; above: shared memory allows 2) */
#ifdef MFCC_MINBLOCKS
    static constexpr int kMinBlocks = MFCC_MINBLOCKS;
#else
    static constexpr int kMinBlocks = (E <= 32) ? (NC >= 256 ? 3 : 4) : 2;
#endif
         /* measured: 3 is 2.5 % faster than 4 at fftLen 512 / 1024 (same 128 registers) */
    static constexpr int kHoistVals = (int)(sizeof(typename BODY::Hoist) / sizeof(cf32));
    /* byte offsets inside the CTA's dynamic shared memory */
    static constexpr int oBuf = 0;
    static constexpr int oWin = oBuf + kUnits * kBufBytes;
    static constexpr int oPark = oWin + 2 * NC * 4;
    static constexpr int oBar = oPark + kHoistVals * 32 * 8;
    static constexpr int oDyn = oBar + kUnits * 8;              /* then: coefsP[nRows*2T], dctT[nbMel*nbDct], mel[kUnits*F*nbMel], part[kUnits*F*nbMel*(T+1)] */
    static int smem_bytes(uint32_t nbMel, uint32_t nbDct, uint32_t nRows)
    {
        return oDyn + (int)(nRows * kRowTaps * 4 + ((nbMel * nbDct + 3) & ~3u) * 4 + kUnits * F * nbMel * 4 + kUnits * F * nbMel * kPartStride * 4);
    }
};

template <int NC>
__global__ void __launch_bounds__(MfccPipe<NC>::kCtaThreads, MfccPipe<NC>::kMinBlocks) mfcc_kernel_pipe(const MfccArgs a, const uint64_t nFrames, const __grid_constant__ MfccRows rows)
{
    typedef MfccPipe<NC> MP;
    typedef typename MP::PL PL;
    typedef typename MP::BODY BODY;
    typedef typename MP::Eng Eng;
    constexpr int T = MP::T, F = MP::F, E = MP::E;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int unit = threadIdx.x >> 5, ut = threadIdx.x & 31;
    const int fl = ut / T, i = ut % T;
    cf32 *buf = reinterpret_cast<cf32 *>(smem_raw + MP::oBuf + unit * MP::kBufBytes);
    cf32 *slot = buf + fl * PL::kFrameElems;                   /* staged frame, then the frame's exchange area */
    cf32 *scratch = buf + F * PL::kFrameElems + fl * PL::kSpecial;
    float *mag = reinterpret_cast<float *>(slot) + fl * T;    /* after the exchange: |X[k]|, k < NC */
    const cf32 *win = reinterpret_cast<const cf32 *>(smem_raw + MP::oWin);
    cf32 *parked = reinterpret_cast<cf32 *>(smem_raw + MP::oPark);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + MP::oBar) + unit;
    float *coefsP = reinterpret_cast<float *>(smem_raw + MP::oDyn);
    float *dctT = coefsP + a.nRows * MP::kRowTaps;
    float *mel = dctT + ((a.nbMel * a.nbDct + 3) & ~3u) + (unit * F + fl) * a.nbMel;
    float *part = dctT + ((a.nbMel * a.nbDct + 3) & ~3u) + MP::kUnits * F * a.nbMel + (unit * F + fl) * a.nbMel * MP::kPartStride;

    /* ---- once per CTA: tables into shared memory ---- */
    if (ut == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        cf32 *w = reinterpret_cast<cf32 *>(smem_raw + MP::oWin);
        const cf32 *gw = reinterpret_cast<const cf32 *>(a.window);
        for (int k = threadIdx.x; k < NC; k += MP::kCtaThreads) w[k] = gw[k];
        for (uint32_t k = threadIdx.x; k < a.nRows * MP::kRowTaps; k += MP::kCtaThreads) coefsP[k] = a.coefsP[k];
        for (uint32_t k = threadIdx.x; k < a.nbMel * a.nbDct; k += MP::kCtaThreads) dctT[k] = a.dctT[k];
        float *bufAll = reinterpret_cast<float *>(smem_raw + MP::oBuf);          /* no NaN patterns in never-written padding */
        for (int k = threadIdx.x; k < MP::kUnits * MP::kBufBytes / 4; k += MP::kCtaThreads) bufAll[k] = 0.0f;
    }
    if (unit == 0) {
        typename BODY::Hoist h0;
        typename BODY::Args ha{nullptr, nullptr, a.tw, a.twr, nullptr};
        BODY::hoist(h0, ha, i);
        const cf32 *hp = reinterpret_cast<const cf32 *>(&h0);
#pragma unroll
        for (int s = 0; s < MP::kHoistVals; s++) parked[s * 32 + ut] = hp[s];
    }
    __syncthreads();

    const uint64_t nGroups = (nFrames + F - 1) / F;
    const uint64_t stride = (uint64_t)gridDim.x * MP::kUnits;
    auto fetch = [&](uint64_t g) {          /* one thread: start the copies of group g's frames into the unit's slots */
        const uint64_t left = nFrames - g * F;
        const int n = left < (uint64_t)F ? (int)left : F;
        mbar_expect_tx(bar, (uint32_t)(n * MP::kFrameBytes));
        for (int f = 0; f < n; f++)
            bulk_g2s(buf + f * PL::kFrameElems, a.src + (g * F + f) * a.stride, MP::kFrameBytes, bar);
    };
    uint64_t g = (uint64_t)blockIdx.x * MP::kUnits + unit;
    if (ut == 0 && g < nGroups) fetch(g);
    uint32_t parity = 0u;
    const cf32 *pk = parked + ut;
    constexpr int kH0 = BODY::kH0, kNS = BODY::kNS, NB = BODY::NB;
    typedef typename PL::P0 PS0;

    for (; g < nGroups; g += stride) {
        const uint64_t frame = g * F + fl;
        const bool valid = frame < nFrames;
        typename Eng::Regs r;
        mbar_wait(bar, parity);
        parity ^= 1u;

        /* 1. frame into registers in the order of the first pass, |max|, normalise, window (arm_mfcc_f32.c:104-112) */
        float mx = 0.0f;
#pragma unroll
        for (int b = 0; b < E / PS0::R; b++)
#pragma unroll
            for (int e = 0; e < PS0::R; e++) {
                const cf32 v = slot[Eng::template in_index<0>(i, b, e)];
                r.v[b * PS0::R + e] = v;
                mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
            }
        if (!valid) mx = 0.0f;                                  /* stale buffer contents of a frame past the end */
        mx = group_max<T>(mx);
        const float inv = (mx != 0.0f) ? __fdiv_rn(1.0f, mx) : 1.0f;
#pragma unroll
        for (int b = 0; b < E / PS0::R; b++)
#pragma unroll
            for (int e = 0; e < PS0::R; e++) {
                const cf32 w = win[Eng::template in_index<0>(i, b, e)];
                const cf32 v = r.v[b * PS0::R + e];
                const float2 n2 = __fmul2_rn(__fmul2_rn(make_float2(v.x, v.y), make_float2(inv, inv)), make_float2(w.x, w.y));
                r.v[b * PS0::R + e] = cf32{n2.x, n2.y};
            }
        /* 2. forward real FFT (arm_mfcc_f32.c:137): first pass, exchange, last pass */
        {
            typename Eng::template TwRegs<0> t;
#pragma unroll
            for (int s = 0; s < kH0; s++) t.w[s] = pk[s * 32];
            Eng::template compute_pre<0, false>(r, t);
        }
        __syncwarp();                                           /* every lane has consumed the staged frames */
        BODY::phase0_out(r, slot, i);
        __syncwarp();
        BODY::last_in(r, slot, i);
        __syncwarp();                                           /* the exchange has been read back: the slots now take the magnitudes */
        /* 3. last pass + split stage; every bin goes to the magnitude sink (:138-146) */
        {
            cf32 ptw[NB / 2];
#pragma unroll
            for (int m = 0; m < NB / 2; m++) ptw[m] = pk[(kH0 + kNS + m) * 32];
            Eng::template compute<1, false>(r, a.tw, i);
            BODY::split_store(r, scratch, i, ptw, MagSink{mag});
        }
        __syncwarp();
        {
            cf32 stw[kNS];
#pragma unroll
            for (int q = 0; q < kNS; q++) stw[q] = pk[(kH0 + q) * 32];
            BODY::special_bins(scratch, i, stw, MagSink{mag});
        }
        __syncwarp();
#ifndef MFCC_ABLATE
#define MFCC_ABLATE 0                /* timing experiments only (wrong results): 1 no mel loop, 2 no partial sums / log, 4 no DCT */
#endif
        /* 4. mel filter bank (:150-160): a row = 2T consecutive bins, one float2 per lane, against
         * zero-padded coefficients; the row schedule sits in the constant bank, so the loop is
         * branch-uniform and free of dependent loads.  At the end of a filter the lane sums are
         * added four lanes at a time and stored as part[filter][lane / 4]. */
        {
            const float *mrow = mag + 2 * i, *crow = coefsP + 2 * i;
            float acc = 0.0f;
            auto flush = [&](uint32_t f) {
                float v = acc;
                if (MP::kQuad >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (MP::kQuad >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
                if ((i & (MP::kQuad - 1)) == 0) part[f * MP::kPartStride + i / MP::kQuad] = v;
                acc = 0.0f;
            };
            uint32_t rr = (MFCC_ABLATE & 1) ? a.nRows : 0;
            for (; rr + 4 <= a.nRows; rr += 4) {
                uint32_t d[4];
                float2 m[4], c[4];
#pragma unroll
                for (int u = 0; u < 4; u++) d[u] = rows.r[rr + u];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    m[u] = *reinterpret_cast<const float2 *>(mrow + (d[u] & 0xffffu));
                    c[u] = *reinterpret_cast<const float2 *>(crow + (rr + u) * MP::kRowTaps);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    acc = fmaf(m[u].y, c[u].y, fmaf(m[u].x, c[u].x, acc));
                    if ((d[u] >> 16) != 0xffffu) flush(d[u] >> 16);
                }
            }
            for (; rr < a.nRows; rr++) {
                const uint32_t d = rows.r[rr];
                const float2 m = *reinterpret_cast<const float2 *>(mrow + (d & 0xffffu));
                const float2 c = *reinterpret_cast<const float2 *>(crow + rr * MP::kRowTaps);
                acc = fmaf(m.y, c.y, fmaf(m.x, c.x, acc));
                if ((d >> 16) != 0xffffu) flush(d >> 16);
            }
        }
        __syncwarp();                                           /* magnitudes consumed: the slots are free for the next group */
        {
            const uint64_t gn = g + stride;
            if (ut == 0 && gn < nGroups) {
                fence_proxy_async();
                fetch(gn);
            }
        }
        /* 5. sum the partial sums, times max, + 1e-6, log (:146,161-165) */
        for (uint32_t f = (MFCC_ABLATE & 2) ? a.nbMel : i; f < a.nbMel; f += T) {
            const float *p = part + f * MP::kPartStride;
            float s[T / MP::kQuad];
#pragma unroll
            for (int j = 0; j < T / MP::kQuad; j++) s[j] = p[j];
#pragma unroll
            for (int w = T / MP::kQuad / 2; w > 0; w >>= 1)
#pragma unroll
                for (int j = 0; j < w; j++) s[j] += s[j + w];
            mel[f] = logf(fmaf(s[0], (mx != 0.0f) ? mx : 1.0f, 1.0e-6f));
        }
        __syncwarp();
        /* 6. DCT matrix (:167-171) */
        if (valid) {
            for (uint32_t q = i; q < a.nbDct; q += T) {
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
                const float *dq = dctT + q;
                uint32_t f = (MFCC_ABLATE & 4) ? a.nbMel - 1 : 0;
                for (; f + 4 <= a.nbMel; f += 4) {
                    s0 = fmaf(dq[f * a.nbDct], mel[f], s0);
                    s1 = fmaf(dq[(f + 1) * a.nbDct], mel[f + 1], s1);
                    s2 = fmaf(dq[(f + 2) * a.nbDct], mel[f + 2], s2);
                    s3 = fmaf(dq[(f + 3) * a.nbDct], mel[f + 3], s3);
                }
                for (; f < a.nbMel; f++) s0 = fmaf(dq[f * a.nbDct], mel[f], s0);
                a.dst[frame * a.nbDct + q] = (s0 + s1) + (s2 + s3);
            }
        }
        __syncwarp();                                           /* mel[] is rewritten by the next group */
    }
}

#define MF_TRY(call)                                                                  \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) return shim_fail(CMSISDSP_CUDA_ERR_RUNTIME, #call, e_); \
    } while (0)

template <int NC> static int mfcc_launch(const MfccArgs &a, uint64_t nFrames, cudaStream_t st)
{
    typedef typename PlanRfftFwd<NC>::type PL;
    const int smem = PL::kSmemBytes + PL::F * (int)a.nbMel * (int)sizeof(float);
    static bool raised[64] = {};
    int dev = 0;
    MF_TRY(cudaGetDevice(&dev));
    if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !raised[dev]) {
        MF_TRY(cudaFuncSetAttribute(mfcc_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        raised[dev] = true;
    }
    const uint64_t ctas = (nFrames + PL::F - 1) / PL::F;
    if (ctas > 0x7fffffffull) return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "batch too large for one launch", cudaSuccess);
    mfcc_kernel<NC><<<(unsigned)ctas, PL::kThreads, smem, st>>>(a, nFrames);
    shim_count_launch();
    MF_TRY(cudaGetLastError());
    return CMSISDSP_CUDA_OK;
}

template <class X> static int to_device(X **d, const X *h, size_t n)
{
    MF_TRY(cudaMalloc((void **)d, n * sizeof(X)));
    const cudaError_t e = cudaMemcpy(*d, h, n * sizeof(X), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(*d);
        *d = nullptr;
        return shim_fail(CMSISDSP_CUDA_ERR_RUNTIME, "cudaMemcpy (mfcc coefficients)", e);
    }
    return CMSISDSP_CUDA_OK;
}

/* dynamic shared memory and occupancy of the pipelined kernel for this plan (0 CTAs: does not fit -> direct kernel) */
template <int NC> static int mfcc_pipe_prepare(MfccDev *p, const uint32_t *pos, const uint32_t *len, const uint32_t *off, const float *coefs)
{
    typedef MfccPipe<NC> MP;
    p->pipeOcc = 0;
    /* mel schedule: rows of 2T bins (a float2 per lane), starting on an even bin, coefficients zero-padded */
    std::vector<float> cp;
    std::vector<uint32_t> rows;
    for (uint32_t f = 0; f < p->nbMel; f++) {
        const uint32_t lead = pos[f] & 1u, start = pos[f] - lead, total = lead + len[f];
        const uint32_t n = (total + MP::kRowTaps - 1) / MP::kRowTaps;
        for (uint32_t j = 0; j < n; j++) {
            rows.push_back((start + j * MP::kRowTaps) | ((j + 1 == n ? f : 0xffffu) << 16));
            for (uint32_t u = j * MP::kRowTaps; u < (j + 1) * MP::kRowTaps; u++)
                cp.push_back((u >= lead && u < total) ? coefs[off[f] + u - lead] : 0.0f);
        }
        if (n == 0) rows.push_back(0u | (f << 16)), cp.insert(cp.end(), MP::kRowTaps, 0.0f);   /* empty filter: sum 0 */
    }
    if (rows.size() > (size_t)kMfccMaxRows) return CMSISDSP_CUDA_OK;
    p->nRows = (uint32_t)rows.size();
    for (size_t k = 0; k < rows.size(); k++) p->rows.r[k] = rows[k];
    p->pipeSmem = MP::smem_bytes(p->nbMel, p->nbDct, p->nRows);
    if (p->pipeSmem > 227 * 1024) return CMSISDSP_CUDA_OK;
    int rc = to_device(&p->coefsP, cp.data(), cp.size());
    if (rc) return rc;
    MF_TRY(cudaFuncSetAttribute(mfcc_kernel_pipe<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int occ = 0;
    MF_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mfcc_kernel_pipe<NC>, MP::kCtaThreads, p->pipeSmem));
    p->pipeOcc = occ;
    return CMSISDSP_CUDA_OK;
}

template <int NC> static int mfcc_launch_pipe(const MfccDev *p, const MfccArgs &a, uint64_t nFrames, cudaStream_t st)
{
    typedef MfccPipe<NC> MP;
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess || sms <= 0) sms = 148;
    const uint64_t groups = (nFrames + MP::F - 1) / MP::F;
    const uint64_t ctas = (groups + MP::kUnits - 1) / MP::kUnits;
    const uint64_t slots = (uint64_t)p->pipeOcc * (uint64_t)sms;
    mfcc_kernel_pipe<NC><<<(unsigned)(ctas < slots ? ctas : slots), MP::kCtaThreads, p->pipeSmem, st>>>(a, nFrames, p->rows);
    shim_count_launch();
    MF_TRY(cudaGetLastError());
    return CMSISDSP_CUDA_OK;
}

extern "C" int cmsisdsp_cuda_mfcc_plan_create(uint32_t fftLen, uint32_t nbMelFilters, uint32_t nbDctOutputs,
                                              const float *dctCoefs, const uint32_t *filterPos,
                                              const uint32_t *filterLengths, const float *filterCoefs,
                                              const float *windowCoefs, void **plan)
{
    if (!plan || !dctCoefs || !filterPos || !filterLengths || !filterCoefs || !windowCoefs)
        return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc_plan_create: null pointer", cudaSuccess);
    if (fftLen != 256 && fftLen != 512 && fftLen != 1024 && fftLen != 2048 && fftLen != 4096)
        return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc_plan_create: fftLen must be 256..4096 (power of two)", cudaSuccess);
    if (nbMelFilters < 1 || nbMelFilters > 128 || nbDctOutputs < 1 || nbDctOutputs > 128)
        return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc_plan_create: 1..128 mel filters and DCT outputs", cudaSuccess);
    std::vector<uint32_t> off(nbMelFilters);
    size_t taps = 0;
    for (uint32_t f = 0; f < nbMelFilters; f++) {
        if ((uint64_t)filterPos[f] + filterLengths[f] > fftLen / 2)
            return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc_plan_create: a mel filter reaches past bin fftLen/2-1", cudaSuccess);
        off[f] = (uint32_t)taps;
        taps += filterLengths[f];
    }
    const void *tw = nullptr;
    const float *twr = nullptr;
    int rc = shim_rfft_tables(fftLen, &tw, &twr);
    if (rc) return rc;
    MfccDev *p = new MfccDev();
    p->fftLen = fftLen; p->nbMel = nbMelFilters; p->nbDct = nbDctOutputs;
    p->tw = (const cf32 *)tw; p->twr = (const cf32 *)twr;
    MF_TRY(cudaGetDevice(&p->device));
    std::vector<float> dctT((size_t)nbMelFilters * nbDctOutputs);
    for (uint32_t f = 0; f < nbMelFilters; f++)
        for (uint32_t q = 0; q < nbDctOutputs; q++) dctT[(size_t)f * nbDctOutputs + q] = dctCoefs[(size_t)q * nbMelFilters + f];
    if ((rc = to_device(&p->dct, dctCoefs, (size_t)nbMelFilters * nbDctOutputs)) || (rc = to_device(&p->coefs, filterCoefs, taps ? taps : 1)) ||
        (rc = to_device(&p->window, windowCoefs, fftLen)) || (rc = to_device(&p->pos, filterPos, nbMelFilters)) ||
        (rc = to_device(&p->len, filterLengths, nbMelFilters)) || (rc = to_device(&p->off, off.data(), nbMelFilters)) ||
        (rc = to_device(&p->dctT, dctT.data(), dctT.size()))) {
        cmsisdsp_cuda_mfcc_plan_destroy(p);           /* frees whatever was allocated so far (null pointers are fine) */
        return rc;
    }
    switch (fftLen / 2) {
    case 128: rc = mfcc_pipe_prepare<128>(p, filterPos, filterLengths, off.data(), filterCoefs); break;
    case 256: rc = mfcc_pipe_prepare<256>(p, filterPos, filterLengths, off.data(), filterCoefs); break;
    case 512: rc = mfcc_pipe_prepare<512>(p, filterPos, filterLengths, off.data(), filterCoefs); break;
    case 1024: rc = mfcc_pipe_prepare<1024>(p, filterPos, filterLengths, off.data(), filterCoefs); break;
    default: rc = mfcc_pipe_prepare<2048>(p, filterPos, filterLengths, off.data(), filterCoefs); break;
    }
    if (rc) {
        cmsisdsp_cuda_mfcc_plan_destroy(p);
        return rc;
    }
    *plan = p;
    return CMSISDSP_CUDA_OK;
}

extern "C" int cmsisdsp_cuda_mfcc_plan_destroy(void *plan)
{
    MfccDev *p = (MfccDev *)plan;
    if (!p) return CMSISDSP_CUDA_OK;
    cudaFree(p->dct); cudaFree(p->coefs); cudaFree(p->window); cudaFree(p->pos); cudaFree(p->len); cudaFree(p->off);
    cudaFree(p->dctT); cudaFree(p->coefsP);
    delete p;
    return CMSISDSP_CUDA_OK;
}

extern "C" int cmsisdsp_cuda_mfcc_f32(const void *plan, const void *d_src, uint64_t strideFloats, void *d_dst,
                                      uint64_t nFrames, void *stream)
{
    const MfccDev *p = (const MfccDev *)plan;
    if (!p || ((!d_src || !d_dst) && nFrames)) return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc: null plan / pointer", cudaSuccess);
    if (strideFloats == 0 || (strideFloats & 1u) || ((uintptr_t)d_src & 7u))
        return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc: frame stride must be a non-zero even number of floats, source 8-byte aligned", cudaSuccess);
    int dev = -1;
    MF_TRY(cudaGetDevice(&dev));
    if (dev != p->device) return shim_fail(CMSISDSP_CUDA_ERR_NO_PLAN, "mfcc: plan belongs to another device", cudaSuccess);
    if (nFrames == 0) return CMSISDSP_CUDA_OK;
    MfccArgs a{(const float *)d_src, strideFloats, (float *)d_dst, p->dct, p->coefs, p->window, p->pos, p->len, p->off,
               p->tw, p->twr, p->nbMel, p->nbDct, p->dctT, p->coefsP, p->nRows};
    cudaStream_t st = (cudaStream_t)stream;
    /* the TMA-fed kernel needs 16-byte aligned frame starts; CMSISDSP_CUDA_KERNEL=direct forces the other one */
    if (p->pipeOcc > 0 && shim_forced_flavour() != KF_DIRECT && (strideFloats & 3u) == 0 && ((uintptr_t)d_src & 15u) == 0) {
        switch (p->fftLen / 2) {
        case 128: return mfcc_launch_pipe<128>(p, a, nFrames, st);
        case 256: return mfcc_launch_pipe<256>(p, a, nFrames, st);
        case 512: return mfcc_launch_pipe<512>(p, a, nFrames, st);
        case 1024: return mfcc_launch_pipe<1024>(p, a, nFrames, st);
        case 2048: return mfcc_launch_pipe<2048>(p, a, nFrames, st);
        }
    }
    switch (p->fftLen / 2) {
    case 128: return mfcc_launch<128>(a, nFrames, st);
    case 256: return mfcc_launch<256>(a, nFrames, st);
    case 512: return mfcc_launch<512>(a, nFrames, st);
    case 1024: return mfcc_launch<1024>(a, nFrames, st);
    case 2048: return mfcc_launch<2048>(a, nFrames, st);
    }
    return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc: unsupported fftLen", cudaSuccess);
}
