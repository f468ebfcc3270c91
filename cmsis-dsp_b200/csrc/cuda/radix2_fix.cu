/*
 * radix2_fix.cu -- the deprecated fixed-point radix-2 transforms, batched: arm_cfft_radix2_q31 / arm_cfft_radix2_q15
 *   reference: Source/TransformFunctions/arm_cfft_radix2_q31.c:62-318, arm_cfft_radix2_q15.c:62-78,275-386,577-681,
 *              arm_bitreversal.c:121-254 (always applied: the result is in natural order)
 *
 * A different algorithm from arm_cfft_q31 / q15 (log2 N radix-2 stages with their own scaling: inputs >> 1 and the sum
 * halved again in stage 1, the sum halved in the middle stages, none in the last; rounding multiply-accumulates for
 * q31, truncating int16 products for q15), so it has its own kernel rather than an adapter.  It is the reference's loop
 * nest laid over a CTA: the frame sits in shared memory (32-bit lanes, q15 sign-extended), every stage hands its N/2
 * butterflies out to the frame's threads, a barrier between stages, and the store reads position bitrev(k) for
 * output k.  HBM is touched once per point each way; this is a completeness path (deprecated API), not a tuned one.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/cmsisdsp_cuda.h"
#include "kernel_entry.h"

namespace {

__device__ __forceinline__ int32_t rhi(int32_t x, int32_t y) { return (int32_t)(((int64_t)x * y + 0x80000000LL) >> 32); }
__device__ __forceinline__ int32_t rhi_acc(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) + (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
__device__ __forceinline__ int32_t rhi_sub(int32_t a, int32_t x, int32_t y)
{
    return (int32_t)((int64_t)(((uint64_t)(int64_t)a << 32) - (uint64_t)((int64_t)x * y) + 0x80000000ULL) >> 32);
}
__device__ __forceinline__ int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
__device__ __forceinline__ int32_t w16(int32_t v) { return (int32_t)(int16_t)(uint16_t)(uint32_t)v; }    /* store to an int16 variable */

/* one radix-2 DIF butterfly on (a, b) = (x[i], x[l]); STAGE: 0 first, 1 middle, 2 last */
template <bool Q15, bool INV, int STAGE> __device__ __forceinline__ void bfly(int2 &a, int2 &b, int2 w)
{
    int32_t xt, yt;
    if (!Q15) {
        if (STAGE == 0) {
            xt = wsub(a.x >> 1, b.x >> 1); yt = wsub(a.y >> 1, b.y >> 1);
            a = make_int2(wadd(a.x >> 1, b.x >> 1) >> 1, wadd(b.y >> 1, a.y >> 1) >> 1);
        } else {
            xt = wsub(a.x, b.x); yt = wsub(a.y, b.y);
            a = (STAGE == 1) ? make_int2(wadd(a.x, b.x) >> 1, wadd(b.y, a.y) >> 1) : make_int2(wadd(a.x, b.x), wadd(b.y, a.y));
        }
        if (STAGE == 2) { b = make_int2(xt, yt); return; }
        int32_t p0 = rhi(xt, w.x), p1 = rhi(yt, w.x);
        if (!INV) { p0 = rhi_acc(p0, yt, w.y); p1 = rhi_sub(p1, xt, w.y); }
        else      { p0 = rhi_sub(p0, yt, w.y); p1 = rhi_acc(p1, xt, w.y); }
        b = make_int2(p0, p1);
    } else {
        if (STAGE == 0) {
            xt = w16((a.x >> 1) - (b.x >> 1)); yt = w16((a.y >> 1) - (b.y >> 1));
            a = make_int2(w16(((a.x >> 1) + (b.x >> 1)) >> 1), w16(((b.y >> 1) + (a.y >> 1)) >> 1));
        } else {
            xt = w16(a.x - b.x); yt = w16(a.y - b.y);
            a = (STAGE == 1) ? make_int2(w16((a.x + b.x) >> 1), w16((b.y + a.y) >> 1)) : make_int2(w16(a.x + b.x), w16(b.y + a.y));
        }
        if (STAGE == 2) { b = make_int2(xt, yt); return; }
        const int32_t xc = w16((xt * w.x) >> 16), ys = w16((yt * w.y) >> 16), yc = w16((yt * w.x) >> 16), xs = w16((xt * w.y) >> 16);
        b = !INV ? make_int2(w16(xc + ys), w16(yc - xs)) : make_int2(w16(xc - ys), w16(yc + xs));
    }
}

template <bool Q15> struct Elem { typedef int2 type; };
template <> struct Elem<true> { typedef short2 type; };

/* TF threads per frame, FPC frames per CTA; tw: N/2 (cos, +sin) pairs, entry k = W_N^k as 32-bit values */
template <bool Q15, bool INV>
__global__ void radix2_kernel(typename Elem<Q15>::type *data, uint64_t nFrames, int N, int logN, int TF, const int2 *__restrict__ tw)
{
    extern __shared__ int2 sm_all[];
    typedef typename Elem<Q15>::type elem;
    const int fl = threadIdx.x / TF, t = threadIdx.x % TF, FPC = blockDim.x / TF;
    const uint64_t frame = (uint64_t)blockIdx.x * FPC + fl;
    const bool valid = frame < nFrames;
    int2 *sm = sm_all + (size_t)fl * N;
    elem *p = data + (valid ? frame : 0) * (uint64_t)N;
    if (valid)
        for (int k = t; k < N; k += TF) {
            const elem e = p[k];
            sm[k] = make_int2((int32_t)e.x, (int32_t)e.y);
        }
    __syncthreads();
    int n2 = N;
    for (int s = 0; s < logN; s++) {
        const int n1 = n2;
        n2 >>= 1;
        if (valid)
            for (int b = t; b < N / 2; b += TF) {
                const int j = b & (n2 - 1), i = (b / n2) * n1 + j, l = i + n2;
                int2 x = sm[i], y = sm[l];
                const int2 w = tw[j << s];
                if (s == 0) bfly<Q15, INV, 0>(x, y, w);
                else if (s < logN - 1) bfly<Q15, INV, 1>(x, y, w);
                else bfly<Q15, INV, 2>(x, y, w);
                sm[i] = x;
                sm[l] = y;
            }
        __syncthreads();
    }
    if (valid)
        for (int k = t; k < N; k += TF) {
            const int2 v = sm[__brev((unsigned)k) >> (32 - logN)];
            elem e;
            e.x = (decltype(e.x))v.x;
            e.y = (decltype(e.y))v.y;
            p[k] = e;
        }
}

}  // namespace

namespace b200fft {
/* used by cmsisdsp_cuda.cu */
int shim_radix2_launch(int type, void *d_p, uint32_t N, uint64_t nFrames, int inv, const void *tw, cudaStream_t st)
{
    if (nFrames == 0) return CMSISDSP_CUDA_OK;
    int logN = 0;
    while ((1u << logN) < N) logN++;
    const int TF = (int)(N / 2 < 256 ? N / 2 : 256);
    const int FPC = 256 / TF;
    const uint64_t ctas = (nFrames + FPC - 1) / FPC;
    if (ctas > 0x7fffffffull) return shim_fail(CMSISDSP_CUDA_ERR_ARGUMENT, "batch too large for one launch", cudaSuccess);
    const size_t smem = (size_t)FPC * N * sizeof(int2);
    const dim3 grid((unsigned)ctas), block((unsigned)(TF * FPC));
    const int2 *t = (const int2 *)tw;
    if (type == CMSISDSP_CUDA_Q15) {
        if (inv) radix2_kernel<true, true><<<grid, block, smem, st>>>((short2 *)d_p, nFrames, (int)N, logN, TF, t);
        else radix2_kernel<true, false><<<grid, block, smem, st>>>((short2 *)d_p, nFrames, (int)N, logN, TF, t);
    } else {
        if (inv) radix2_kernel<false, true><<<grid, block, smem, st>>>((int2 *)d_p, nFrames, (int)N, logN, TF, t);
        else radix2_kernel<false, false><<<grid, block, smem, st>>>((int2 *)d_p, nFrames, (int)N, logN, TF, t);
    }
    shim_count_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return shim_fail(CMSISDSP_CUDA_ERR_RUNTIME, "radix2 kernel launch", e);
    return CMSISDSP_CUDA_OK;
}
}  // namespace b200fft
