/*
 * cmsisdsp_cuda.cu -- libcmsisdsp_cuda.so: sm_100a kernels for the batched CMSIS-DSP FFT
 * hot path plus the C-ABI declared in include/cmsisdsp_cuda.h.
 *
 * One launch processes a whole batch of independent frames: a CTA holds PL::F frames,
 * PL::T threads each (16 points per thread), and runs the phases of the plan's body
 * (fft_body.cuh) with __syncthreads() between them.  HBM is touched exactly once per
 * point on the way in and once on the way out.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "cmsisdsp_cuda.h"
#include "fft_plans.cuh"

using namespace b200fft;

/* ------------------------------------------------------------------ error plumbing */

static thread_local char g_err[256] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (e != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof g_err, "%s", what);
    return code;
}
#define CU_TRY(call)                                                       \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return fail(CMSISDSP_CUDA_ERR_RUNTIME, #call, e_); \
    } while (0)

/* ------------------------------------------------------------------ kernel */

/* barrier between two phases of a frame: when a frame's T threads sit inside one warp the
 * exchange is warp-private and __syncwarp() is enough (no CTA-wide stall) */
template <class PL> __device__ __forceinline__ void frame_sync()
{
    if constexpr (PL::T <= 32) __syncwarp();
    else __syncthreads();
}

template <class BODY, class PL>
__global__ void __launch_bounds__(PL::kThreads) frame_kernel(typename BODY::Args base, uint64_t nFrames)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename BODY::elem elem;
    const int tid = threadIdx.x;
    const int fl = tid / PL::T, i = tid % PL::T;
    const uint64_t frame = (uint64_t)blockIdx.x * PL::F + fl;
    const bool valid = frame < nFrames;
    elem *sm = reinterpret_cast<elem *>(smem_raw) + fl * PL::kFrameElems;
    const typename BODY::Args a = BODY::for_frame(base, valid ? frame : 0);
    typename BODY::Regs r;

    if (valid) BODY::template phase<0>(r, a, sm, i);
    if constexpr (BODY::kPhases > 1) {
        frame_sync<PL>();
        if (valid) BODY::template phase<1>(r, a, sm, i);
    }
    if constexpr (BODY::kPhases > 2) {
        frame_sync<PL>();
        if (valid) BODY::template phase<2>(r, a, sm, i);
        frame_sync<PL>();
        if (valid) BODY::template phase<3>(r, a, sm, i);
    }
}

/* ------------------------------------------------------------------ persistent TMA-staged kernel
 *
 * One CTA per resident slot, looping over frame groups (PL::F consecutive frames).  The
 * next group is fetched with ONE bulk async copy (cp.async.bulk, the 1-D TMA path) into the
 * other half of a double buffer while the current group is being transformed, so HBM reads
 * never wait on the butterflies: bytes in flight per SM = resident CTAs x group size,
 * independent of register pressure.  Completion is tracked by one mbarrier per buffer.
 * Results leave through the registers (coalesced streaming stores), as in frame_kernel. */

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

template <class PL> struct StagedSmem {
    typedef typename PL::Arith::elem elem;
    static constexpr int kStageElems = PL::F * PL::N;                  /* one group, linear, as in HBM */
    static constexpr int kStageBytes = kStageElems * (int)sizeof(elem);
    static constexpr int kExchBytes = ((PL::NP > 1 ? PL::F * PL::kFrameElems * (int)sizeof(elem) : 0) + 15) & ~15;
    static constexpr int kBytes = 2 * kStageBytes + kExchBytes + 16;   /* + two mbarriers */
};

template <class BODY, class PL>
__global__ void __launch_bounds__(PL::kThreads) frame_kernel_staged(typename BODY::Args base, uint64_t nFrames)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename BODY::elem elem;
    typedef StagedSmem<PL> SM;
    elem *stage0 = reinterpret_cast<elem *>(smem_raw);
    elem *exch = reinterpret_cast<elem *>(smem_raw + 2 * SM::kStageBytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * SM::kStageBytes + SM::kExchBytes);

    const int tid = threadIdx.x;
    const int fl = tid / PL::T, i = tid % PL::T;
    const uint64_t nGroups = (nFrames + PL::F - 1) / PL::F;
    elem *sm = exch + fl * PL::kFrameElems;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto group_bytes = [&](uint64_t g) -> uint32_t {
        const uint64_t left = nFrames - g * PL::F;
        return (uint32_t)((left < (uint64_t)PL::F ? left : (uint64_t)PL::F) * PL::N * sizeof(elem));
    };
    uint64_t g = blockIdx.x;
    if (tid == 0 && g < nGroups) {
        const uint32_t bytes = group_bytes(g);
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(stage0, base.in + g * (uint64_t)SM::kStageElems, bytes, &bar[0]);
    }
    uint32_t parity[2] = {0u, 0u};
    int cur = 0;
    for (; g < nGroups; g += gridDim.x, cur ^= 1) {
        const uint64_t gn = g + gridDim.x;
        if (tid == 0 && gn < nGroups) {          /* prefetch the next group into the other buffer */
            const uint32_t bytes = group_bytes(gn);
            mbar_expect_tx(&bar[cur ^ 1], bytes);
            bulk_g2s(stage0 + (cur ^ 1) * SM::kStageElems, base.in + gn * (uint64_t)SM::kStageElems, bytes, &bar[cur ^ 1]);
        }
        const uint64_t frame = g * PL::F + fl;
        const bool valid = frame < nFrames;
        typename BODY::Args a = BODY::for_frame(base, valid ? frame : 0);
        a.in = stage0 + cur * SM::kStageElems + fl * PL::N;      /* the staged copy of this frame */
        typename BODY::Regs r;

        mbar_wait(&bar[cur], parity[cur]);
        parity[cur] ^= 1u;
        if (valid) BODY::template phase<0>(r, a, sm, i);
        if constexpr (BODY::kPhases > 1) {
            __syncthreads();
            if (valid) BODY::template phase<1>(r, a, sm, i);
        }
        if constexpr (BODY::kPhases > 2) {
            __syncthreads();
            if (valid) BODY::template phase<2>(r, a, sm, i);
            __syncthreads();
            if (valid) BODY::template phase<3>(r, a, sm, i);
        }
        /* everyone is done with stage[cur] and the exchange buffer before the next iteration
         * refills the former (prefetch of g + 2*gridDim) and overwrites the latter */
        __syncthreads();
    }
}

static int g_numSMs[64] = {0};
static int num_sms()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!g_numSMs[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_numSMs[dev] = n;
    }
    return g_numSMs[dev];
}

/* resident CTAs per SM of the staged kernel (also raises its dynamic shared memory limit once) */
template <class BODY, class PL> static int staged_occupancy(int *occOut)
{
    static int occ[64] = {0};
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "device index out of range");
    if (!occ[dev]) {
        CU_TRY(cudaFuncSetAttribute(frame_kernel_staged<BODY, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, StagedSmem<PL>::kBytes));
        int o = 0;
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, frame_kernel_staged<BODY, PL>, PL::kThreads, StagedSmem<PL>::kBytes));
        if (o < 1) return fail(CMSISDSP_CUDA_ERR_RUNTIME, "staged kernel does not fit on an SM");
        occ[dev] = o;
    }
    *occOut = occ[dev];
    return CMSISDSP_CUDA_OK;
}

template <class BODY, class PL>
static int launch_staged(const typename BODY::Args &args, uint64_t nFrames, cudaStream_t st)
{
    if (nFrames == 0) return CMSISDSP_CUDA_OK;
    int occ = 0;
    int rc = staged_occupancy<BODY, PL>(&occ);
    if (rc) return rc;
    const uint64_t groups = (nFrames + PL::F - 1) / PL::F;
    const uint64_t slots = (uint64_t)occ * (uint64_t)num_sms();
    const unsigned grid = (unsigned)(groups < slots ? groups : slots);
    frame_kernel_staged<BODY, PL><<<grid, PL::kThreads, StagedSmem<PL>::kBytes, st>>>(args, nFrames);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CU_TRY(cudaGetLastError());
    return CMSISDSP_CUDA_OK;
}

/* kernel flavour: 0 = direct loads (frame_kernel), 1 = persistent TMA-staged (default for the
 * lengths listed in use_staged()).  CMSISDSP_CUDA_KERNEL=direct|staged overrides, for A/B runs. */
static int kernel_override()
{
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("CMSISDSP_CUDA_KERNEL");
        v = !e ? -1 : (!strcmp(e, "direct") ? 0 : (!strcmp(e, "staged") ? 1 : -1));
    }
    return v;
}
static bool use_staged(uint32_t complexLen, const void *in)
{
    if (((uintptr_t)in & 15u) != 0) return false;          /* bulk copies need 16-byte aligned sources */
    const int o = kernel_override();
    if (o >= 0) return o == 1;
    (void)complexLen;
    return false;
}

template <class BODY, class PL>
static int launch(const typename BODY::Args &args, uint64_t nFrames, cudaStream_t st)
{
    if (nFrames == 0) return CMSISDSP_CUDA_OK;
    const uint64_t ctas = (nFrames + PL::F - 1) / PL::F;
    if (ctas > 0x7fffffffull) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "batch too large for one launch");
    frame_kernel<BODY, PL><<<(unsigned)ctas, PL::kThreads, PL::kSmemBytes, st>>>(args, nFrames);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CU_TRY(cudaGetLastError());
    return CMSISDSP_CUDA_OK;
}

template <class BODY, class PL>
static int kinfo(int *threads, int *frames, int *smem, int *regs, int *ctasPerSm)
{
    cudaFuncAttributes fa;
    CU_TRY(cudaFuncGetAttributes(&fa, frame_kernel<BODY, PL>));
    int occ = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, frame_kernel<BODY, PL>, PL::kThreads, PL::kSmemBytes));
    if (threads) *threads = PL::kThreads;
    if (frames) *frames = PL::F;
    if (smem) *smem = PL::kSmemBytes;
    if (regs) *regs = fa.numRegs;
    if (ctasPerSm) *ctasPerSm = occ;
    return CMSISDSP_CUDA_OK;
}

/* ------------------------------------------------------------------ plan cache */

static const uint32_t kLens[9] = {16, 32, 64, 128, 256, 512, 1024, 2048, 4096};
static int len_index(uint32_t n)
{
    for (int i = 0; i < 9; i++)
        if (kLens[i] == n) return i;
    return -1;
}

struct DevPlan {
    void *tw = nullptr;           /* pass-ordered twiddles of the cfft plan (Plan::build_twiddles) */
    void *tw_rfwd = nullptr;      /* f32 only: same for the rfft forward / inverse plans of complex length N */
    void *tw_rinv = nullptr;
    uint16_t *perm = nullptr;     /* destination position of X[k] when bitReverseFlag == 0 */
};
struct DevState {
    DevPlan plan[3][9];
    float *twr[9] = {};           /* rfft twiddles, indexed by len_index(real length) */
};
static const int kMaxDev = 64;
static DevState g_dev[kMaxDev];
static std::mutex g_mu;

static int cur_device(int *dev)
{
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "cudaGetDevice", e);
    if (*dev < 0 || *dev >= kMaxDev) return fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "device index out of range");
    return CMSISDSP_CUDA_OK;
}

#define FOR_ALL_N(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)
#define FOR_RFFT_NC(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

/* re-order the reference-layout twiddles into the plan's pass order and upload them */
template <class PL> static int upload_pass_ordered(const void *base, void **dOut)
{
    typedef typename PL::Arith::elem elem;
    std::vector<elem> host((size_t)PL::kTwEntries + 1);
    PL::build_twiddles((const elem *)base, host.data());
    void *d = nullptr;
    CU_TRY(cudaMalloc(&d, host.size() * sizeof(elem)));
    CU_TRY(cudaMemcpy(d, host.data(), host.size() * sizeof(elem), cudaMemcpyHostToDevice));
    *dOut = d;
    return CMSISDSP_CUDA_OK;
}

static int build_tables(int type, uint32_t fftLen, const void *base, DevPlan &p)
{
    int rc = CMSISDSP_CUDA_ERR_ARGUMENT;
    switch (fftLen) {
#define CASE(n)                                                                                         \
    case n:                                                                                             \
        if (type == CMSISDSP_CUDA_F32) rc = upload_pass_ordered<PlanCfftF32<n>::type>(base, &p.tw);     \
        else if (type == CMSISDSP_CUDA_Q31) rc = upload_pass_ordered<PlanCfftFix<ArithQ31, n>::type>(base, &p.tw); \
        else rc = upload_pass_ordered<PlanCfftFix<ArithQ15, n>::type>(base, &p.tw);                     \
        break;
        FOR_ALL_N(CASE)
#undef CASE
    }
    if (rc || type != CMSISDSP_CUDA_F32) return rc;
    switch (fftLen) {
#define CASE(nc)                                                                       \
    case nc:                                                                           \
        rc = upload_pass_ordered<PlanRfftFwd<nc>::type>(base, &p.tw_rfwd);             \
        if (!rc) rc = upload_pass_ordered<PlanRfftInv<nc>::type>(base, &p.tw_rinv);    \
        break;
        FOR_RFFT_NC(CASE)
#undef CASE
    }
    return rc;
}

extern "C" int cmsisdsp_cuda_plan_upload(int type, uint32_t fftLen, const void *pTwiddle,
                                         const uint16_t *pBitRevTable, uint16_t bitRevLength)
{
    const int li = len_index(fftLen);
    if (type < 0 || type > 2 || li < 0 || !pTwiddle || (!pBitRevTable && bitRevLength))
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "plan_upload: bad type / length / pointer");
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    DevPlan &p = g_dev[dev].plan[type][li];
    if (p.tw) return CMSISDSP_CUDA_OK;

    /* out[k] = scrambled[P[k]] after the swap list  =>  X[k] lives at scrambled position P[k] */
    std::vector<uint16_t> perm(fftLen);
    for (uint32_t k = 0; k < fftLen; k++) perm[k] = (uint16_t)k;
    for (uint32_t i = 0; i + 1 < bitRevLength; i += 2) {
        const uint32_t a = pBitRevTable[i] >> 3, b = pBitRevTable[i + 1] >> 3;
        if (a >= fftLen || b >= fftLen) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "plan_upload: bit-reversal entry out of range");
        const uint16_t t = perm[a]; perm[a] = perm[b]; perm[b] = t;
    }
    uint16_t *dperm = nullptr;
    CU_TRY(cudaMalloc((void **)&dperm, fftLen * sizeof(uint16_t)));
    CU_TRY(cudaMemcpy(dperm, perm.data(), fftLen * sizeof(uint16_t), cudaMemcpyHostToDevice));
    DevPlan np;
    np.perm = dperm;
    rc = build_tables(type, fftLen, pTwiddle, np);
    if (rc) return rc;
    p = np;
    return CMSISDSP_CUDA_OK;
}

extern "C" int cmsisdsp_cuda_rfft_plan_upload(uint32_t fftLenReal, const float *pTwiddleRFFT)
{
    const int li = len_index(fftLenReal);
    if (li < 1 || !pTwiddleRFFT) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_plan_upload: bad length / pointer");
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_dev[dev].twr[li]) return CMSISDSP_CUDA_OK;
    float *d = nullptr;
    CU_TRY(cudaMalloc((void **)&d, fftLenReal * sizeof(float)));
    CU_TRY(cudaMemcpy(d, pTwiddleRFFT, fftLenReal * sizeof(float), cudaMemcpyHostToDevice));
    g_dev[dev].twr[li] = d;
    return CMSISDSP_CUDA_OK;
}

extern "C" int cmsisdsp_cuda_plan_ready(int type, uint32_t fftLen)
{
    const int li = len_index(fftLen);
    int dev;
    if (type < 0 || type > 2 || li < 0 || cur_device(&dev)) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    return g_dev[dev].plan[type][li].tw != nullptr;
}
extern "C" int cmsisdsp_cuda_rfft_plan_ready(uint32_t fftLenReal)
{
    const int li = len_index(fftLenReal);
    int dev;
    if (li < 1 || cur_device(&dev)) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    return g_dev[dev].twr[li] != nullptr && g_dev[dev].plan[0][li - 1].tw != nullptr;
}

static int get_plan(int type, uint32_t fftLen, DevPlan *out)
{
    const int li = len_index(fftLen);
    if (li < 0) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported fftLen (16..4096, power of two)");
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    *out = g_dev[dev].plan[type][li];
    if (!out->tw) return fail(CMSISDSP_CUDA_ERR_NO_PLAN, "no plan uploaded for this (device, type, fftLen)");
    return CMSISDSP_CUDA_OK;
}

/* ------------------------------------------------------------------ transforms */

template <class AR, class PL>
static int cfft_launch(void *d_p, uint64_t nFrames, bool inv, const DevPlan &pl, bool bitrev, int shl1, cudaStream_t st)
{
    typedef typename AR::elem elem;
    const bool staged = use_staged(PL::N, d_p);
#define GO(INV_, STG_)                                                                                                       \
    {                                                                                                                        \
        typedef CfftBody<PL, INV_, STG_> BODY;                                                                               \
        typename BODY::Args a{(const elem *)d_p, (elem *)d_p, (const elem *)pl.tw, bitrev ? nullptr : pl.perm, 1.0f / (float)PL::N, shl1}; \
        if constexpr (STG_) return launch_staged<BODY, PL>(a, nFrames, st);                                                  \
        else return launch<BODY, PL>(a, nFrames, st);                                                                        \
    }
    if (inv) { if (staged) GO(true, true) else GO(true, false) }
    if (staged) GO(false, true) else GO(false, false)
#undef GO
}

static int cfft_any(int type, void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{
    if (!d_p && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    DevPlan pl;
    int rc = get_plan(type, fftLen, &pl);
    if (rc) return rc;
    const bool inv = (ifftFlag == 1), br = (bitReverseFlag != 0);
    cudaStream_t st = (cudaStream_t)stream;
    int lg = 0;
    while ((1u << lg) < fftLen) lg++;
    const int shl1 = lg & 1;     /* N = 2*4^m: final << 1 (fixed point only) */
    switch (fftLen) {
#define CASE(n)                                                                                                   \
    case n:                                                                                                       \
        if (type == CMSISDSP_CUDA_F32) return cfft_launch<ArithF32, PlanCfftF32<n>::type>(d_p, nFrames, inv, pl, br, 0, st);          \
        if (type == CMSISDSP_CUDA_Q31) return cfft_launch<ArithQ31, PlanCfftFix<ArithQ31, n>::type>(d_p, nFrames, inv, pl, br, shl1, st); \
        return cfft_launch<ArithQ15, PlanCfftFix<ArithQ15, n>::type>(d_p, nFrames, inv, pl, br, shl1, st);
        FOR_ALL_N(CASE)
#undef CASE
    }
    return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported fftLen");
}

extern "C" int cmsisdsp_cuda_cfft_f32(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{ return cfft_any(CMSISDSP_CUDA_F32, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag, stream); }
extern "C" int cmsisdsp_cuda_cfft_q31(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{ return cfft_any(CMSISDSP_CUDA_Q31, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag, stream); }
extern "C" int cmsisdsp_cuda_cfft_q15(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{ return cfft_any(CMSISDSP_CUDA_Q15, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag, stream); }

extern "C" int cmsisdsp_cuda_rfft_fast_f32(const void *d_p, void *d_out, uint32_t fftLenReal, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{
    if ((!d_p || !d_out) && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    if (d_p == d_out && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fast: p and pOut must not alias");
    const int li = len_index(fftLenReal);
    if (li < 1) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported rfft length (32..4096, power of two)");
    DevPlan pl;
    int rc = get_plan(CMSISDSP_CUDA_F32, fftLenReal / 2, &pl);
    if (rc) return rc;
    int dev;
    rc = cur_device(&dev);
    if (rc) return rc;
    const float *twr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        twr = g_dev[dev].twr[li];
    }
    if (!twr) return fail(CMSISDSP_CUDA_ERR_NO_PLAN, "no rfft plan uploaded for this (device, fftLen)");
    cudaStream_t st = (cudaStream_t)stream;
    switch (fftLenReal / 2) {
#define CASE(nc)                                                                                              \
    case nc:                                                                                                  \
        if (!ifftFlag) {                                                                                      \
            typedef PlanRfftFwd<nc>::type PL;                                                                 \
            RfftFwdBody<PL>::Args a{(const cf32 *)d_p, (cf32 *)d_out, (const cf32 *)pl.tw_rfwd, (const cf32 *)twr}; \
            if (use_staged(nc, d_p)) return launch_staged<RfftFwdBody<PL, true>, PL>(RfftFwdBody<PL, true>::Args{a.in, a.out, a.tw, a.twr}, nFrames, st); \
            return launch<RfftFwdBody<PL>, PL>(a, nFrames, st);                                               \
        } else {                                                                                              \
            typedef PlanRfftInv<nc>::type PL;                                                                 \
            RfftInvBody<PL>::Args a{(const cf32 *)d_p, (cf32 *)d_out, (const cf32 *)pl.tw_rinv, (const cf32 *)twr, 1.0f / (float)nc}; \
            if (use_staged(nc, d_p)) return launch_staged<RfftInvBody<PL, true>, PL>(RfftInvBody<PL, true>::Args{a.in, a.out, a.tw, a.twr, a.scale}, nFrames, st); \
            return launch<RfftInvBody<PL>, PL>(a, nFrames, st);                                               \
        }
        FOR_RFFT_NC(CASE)
#undef CASE
    }
    return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported rfft length");
}

extern "C" int cmsisdsp_cuda_kernel_info(int op, uint32_t fftLen, int *threads, int *frames, int *smem, int *regs, int *ctasPerSm)
{
    const uint32_t n = (op >= 3) ? fftLen / 2 : fftLen;
    switch (n) {
#define CASE(nn)                                                                                                   \
    case nn:                                                                                                       \
        if (op == 0) return kinfo<CfftBody<PlanCfftF32<nn>::type, false>, PlanCfftF32<nn>::type>(threads, frames, smem, regs, ctasPerSm); \
        if (op == 1) return kinfo<CfftBody<PlanCfftFix<ArithQ31, nn>::type, false>, PlanCfftFix<ArithQ31, nn>::type>(threads, frames, smem, regs, ctasPerSm); \
        if (op == 2) return kinfo<CfftBody<PlanCfftFix<ArithQ15, nn>::type, false>, PlanCfftFix<ArithQ15, nn>::type>(threads, frames, smem, regs, ctasPerSm); \
        break;
        FOR_ALL_N(CASE)
#undef CASE
    }
    switch (n) {
#define CASE(nc)                                                                                                   \
    case nc:                                                                                                       \
        if (op == 3) return kinfo<RfftFwdBody<PlanRfftFwd<nc>::type>, PlanRfftFwd<nc>::type>(threads, frames, smem, regs, ctasPerSm); \
        if (op == 4) return kinfo<RfftInvBody<PlanRfftInv<nc>::type>, PlanRfftInv<nc>::type>(threads, frames, smem, regs, ctasPerSm); \
        break;
        FOR_RFFT_NC(CASE)
#undef CASE
    }
    return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "kernel_info: unsupported (op, fftLen)");
}

/* ------------------------------------------------------------------ plumbing */

extern "C" int cmsisdsp_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
extern "C" int cmsisdsp_cuda_set_device(int device) { CU_TRY(cudaSetDevice(device)); return 0; }
extern "C" int cmsisdsp_cuda_get_device(void) { int d = -1; if (cudaGetDevice(&d) != cudaSuccess) return -1; return d; }
extern "C" int cmsisdsp_cuda_malloc(void **p, size_t bytes) { if (!p) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null"); CU_TRY(cudaMalloc(p, bytes)); return 0; }
extern "C" int cmsisdsp_cuda_free(void *p) { CU_TRY(cudaFree(p)); return 0; }
extern "C" int cmsisdsp_cuda_host_alloc(void **p, size_t bytes) { if (!p) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null"); CU_TRY(cudaHostAlloc(p, bytes, cudaHostAllocDefault)); return 0; }
extern "C" int cmsisdsp_cuda_host_free(void *p) { CU_TRY(cudaFreeHost(p)); return 0; }
extern "C" int cmsisdsp_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{ CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream)); return 0; }
extern "C" int cmsisdsp_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{ CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream)); return 0; }
extern "C" int cmsisdsp_cuda_stream_create(void **stream)
{ if (!stream) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null"); cudaStream_t s; CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); *stream = (void *)s; return 0; }
extern "C" int cmsisdsp_cuda_stream_destroy(void *stream) { CU_TRY(cudaStreamDestroy((cudaStream_t)stream)); return 0; }
extern "C" int cmsisdsp_cuda_stream_synchronize(void *stream) { CU_TRY(cudaStreamSynchronize((cudaStream_t)stream)); return 0; }

extern "C" int cmsisdsp_cuda_is_device_pointer(const void *ptr)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, ptr);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(CMSISDSP_CUDA_ERR_RUNTIME, "cudaPointerGetAttributes", e); }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? 1 : 0;
}

struct Timer { cudaEvent_t e0, e1; };
extern "C" int cmsisdsp_cuda_timer_begin(void **timer, void *stream)
{
    if (!timer) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null");
    Timer *t = new Timer;
    CU_TRY(cudaEventCreate(&t->e0));
    CU_TRY(cudaEventCreate(&t->e1));
    CU_TRY(cudaEventRecord(t->e0, (cudaStream_t)stream));
    *timer = t;
    return 0;
}
extern "C" int cmsisdsp_cuda_timer_end(void *timer, void *stream, float *ms)
{
    Timer *t = (Timer *)timer;
    if (!t) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null");
    CU_TRY(cudaEventRecord(t->e1, (cudaStream_t)stream));
    CU_TRY(cudaEventSynchronize(t->e1));
    float v = 0;
    CU_TRY(cudaEventElapsedTime(&v, t->e0, t->e1));
    if (ms) *ms = v;
    cudaEventDestroy(t->e0); cudaEventDestroy(t->e1);
    delete t;
    return 0;
}

extern "C" const char *cmsisdsp_cuda_last_error(void) { return g_err; }
extern "C" uint64_t cmsisdsp_cuda_launch_count(void) { return g_launches.load(); }
