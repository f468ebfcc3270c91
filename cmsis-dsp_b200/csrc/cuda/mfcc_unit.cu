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
 */
#include <cuda_runtime.h>
#include <stdlib.h>

#include <vector>

#include "../../../include/cmsisdsp_cuda.h"
#include "fft_plans.cuh"
#include "kernel_entry.h"

using namespace b200fft;

struct MfccDev {
    uint32_t fftLen, nbMel, nbDct;
    float *dct, *coefs, *window;          /* device copies of the caller's coefficient arrays */
    uint32_t *pos, *len, *off;            /* off[f] = start of filter f in coefs */
    const cf32 *tw, *twr;                 /* tables of the rfft plan (owned by the plan cache) */
    int device;
};

struct MfccArgs {
    const float *src;
    uint64_t stride;                      /* floats between frame starts */
    float *dst;
    const float *dct, *coefs, *window;
    const uint32_t *pos, *len, *off;
    const cf32 *tw, *twr;
    uint32_t nbMel, nbDct;
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    MF_TRY(cudaMemcpy(*d, h, n * sizeof(X), cudaMemcpyHostToDevice));
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
    if ((rc = to_device(&p->dct, dctCoefs, (size_t)nbMelFilters * nbDctOutputs)) || (rc = to_device(&p->coefs, filterCoefs, taps ? taps : 1)) ||
        (rc = to_device(&p->window, windowCoefs, fftLen)) || (rc = to_device(&p->pos, filterPos, nbMelFilters)) ||
        (rc = to_device(&p->len, filterLengths, nbMelFilters)) || (rc = to_device(&p->off, off.data(), nbMelFilters))) {
        delete p;
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
               p->tw, p->twr, p->nbMel, p->nbDct};
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->fftLen / 2) {
    case 128: return mfcc_launch<128>(a, nFrames, st);
    case 256: return mfcc_launch<256>(a, nFrames, st);
    case 512: return mfcc_launch<512>(a, nFrames, st);
    case 1024: return mfcc_launch<1024>(a, nFrames, st);
    case 2048: return mfcc_launch<2048>(a, nFrames, st);
    }
    return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "mfcc: unsupported fftLen", cudaSuccess);
}
