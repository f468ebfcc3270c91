/*
 * cmsisdsp_cuda.cu -- libcmsisdsp_cuda.so: the C ABI declared in include/cmsisdsp_cuda.h.
 *
 * This file holds the device/stream plumbing, the per-device plan cache (pass-ordered
 * twiddle tables + output permutations) and the dispatch to the per-(op, length) kernel
 * objects (kernel_unit.cu, one object per pair).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "cmsisdsp_cuda.h"
#include "kernel_entry.h"

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
namespace b200fft {
int shim_fail(int code, const char *what, cudaError_t e) { return fail(code, what, e); }
void shim_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}
#define CU_TRY(call)                                                       \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return fail(CMSISDSP_CUDA_ERR_RUNTIME, #call, e_); \
    } while (0)

/* ------------------------------------------------------------------ kernel table */

#define FOR_ALL_N(X, op) X(op, 16) X(op, 32) X(op, 64) X(op, 128) X(op, 256) X(op, 512) X(op, 1024) X(op, 2048) X(op, 4096)
#define FOR_RFFT_NC(X, op) X(op, 16) X(op, 32) X(op, 64) X(op, 128) X(op, 256) X(op, 512) X(op, 1024) X(op, 2048)
#define DECL(op, n) extern const KernelEntry ku_entry_##op##_##n;
namespace b200fft {
FOR_ALL_N(DECL, 0) FOR_ALL_N(DECL, 1) FOR_ALL_N(DECL, 2) FOR_RFFT_NC(DECL, 3) FOR_RFFT_NC(DECL, 4)
FOR_ALL_N(DECL, 5) FOR_ALL_N(DECL, 6) FOR_ALL_N(DECL, 7) FOR_ALL_N(DECL, 8) FOR_ALL_N(DECL, 9) FOR_ALL_N(DECL, 10)
FOR_RFFT_NC(DECL, 11) FOR_RFFT_NC(DECL, 12)
}
#undef DECL
#define REF(op, n) &ku_entry_##op##_##n,
/* [op][index of the COMPLEX length 16..4096]; the rfft ops have no 4096-point complex plan */
static const KernelEntry *const kEntries[OP_COUNT][9] = {
    {FOR_ALL_N(REF, 0)}, {FOR_ALL_N(REF, 1)}, {FOR_ALL_N(REF, 2)}, {FOR_RFFT_NC(REF, 3) nullptr}, {FOR_RFFT_NC(REF, 4) nullptr},
    {FOR_ALL_N(REF, 5)}, {FOR_ALL_N(REF, 6)}, {FOR_ALL_N(REF, 7)}, {FOR_ALL_N(REF, 8)}, {FOR_ALL_N(REF, 9)}, {FOR_ALL_N(REF, 10)}, {FOR_RFFT_NC(REF, 11) nullptr}, {FOR_RFFT_NC(REF, 12) nullptr}};
#undef REF

static const uint32_t kLens[9] = {16, 32, 64, 128, 256, 512, 1024, 2048, 4096};
static int len_index(uint32_t n)
{
    for (int i = 0; i < 9; i++)
        if (kLens[i] == n) return i;
    return -1;
}

/* kernel flavour (kernel_entry.h): the unit's measured preference, unless
 * CMSISDSP_CUDA_KERNEL=direct|pipe forces a flavour (for A/B measurements) */
static std::atomic<int> g_flavour{-2};           /* -2: not read yet, -1: per-unit default, else forced */
namespace b200fft {
int shim_forced_flavour()
{
    int forced = g_flavour.load(std::memory_order_relaxed);
    if (forced == -2) {
        const char *e = getenv("CMSISDSP_CUDA_KERNEL");
        forced = !e ? -1 : (!strcmp(e, "direct") ? KF_DIRECT : (!strcmp(e, "pipe") ? KF_PIPE : -1));
        g_flavour.store(forced, std::memory_order_relaxed);
    }
    return forced;
}
}
static int choose_flavour(const KernelEntry *ke)
{
    const int forced = shim_forced_flavour();
    if (forced >= 0) return (forced == KF_PIPE && !ke->hasPipe) ? KF_DIRECT : forced;
    return ke->preferPipe ? KF_PIPE : KF_DIRECT;
}
extern "C" int cmsisdsp_cuda_set_kernel_flavour(int flavour)
{
    if (flavour < -1 || flavour > KF_PIPE) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "set_kernel_flavour: -1 (default), 0 (direct) or 1 (pipelined)");
    g_flavour.store(flavour, std::memory_order_relaxed);
    return CMSISDSP_CUDA_OK;
}

/* ------------------------------------------------------------------ plan cache
 *
 * Device-resident tables, keyed by (device, table kind, length) AND by the CONTENT of the caller's host tables: an
 * instance is plain data that may point at any table (the reference reads whatever the struct points at), so a second
 * instance with other values for the same length gets its own device copy instead of silently running on the first
 * one's.  Up to kSlots distinct tables per key.  Recognising a table costs a pointer compare plus a 32-sample
 * fingerprint; the full 64-bit hash is taken only when the pointer or the fingerprint is new.  Which slot a thread's
 * next transform uses is thread-local (set by the upload call that precedes it), so two host threads with different
 * tables for the same length do not disturb each other.  Device copies are never freed: they live as long as the
 * process (a kernel of another thread may still be reading them). */

static const int kSlots = 4;
static const int kMaxDev = 64;

static uint64_t fnv1a(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull)
{
    const unsigned char *p = (const unsigned char *)data;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * 1099511628211ull;
    }
    for (; i < bytes; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}
/* 32 four-byte samples spread over the table (tables are at least 64 bytes) */
static uint32_t fingerprint(const void *data, size_t bytes)
{
    const unsigned char *p = (const unsigned char *)data;
    const size_t words = bytes / 4, step = words >= 32 ? words / 32 : 1;
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < words; i += step) {
        uint32_t w;
        memcpy(&w, p + 4 * i, 4);
        h = (h ^ w) * 16777619u;
    }
    return h ^ (uint32_t)bytes;
}

struct DevPlan {
    void *tw = nullptr;           /* pass-ordered twiddles of the cfft plan (Plan::build_twiddles) */
    void *tw_rfwd = nullptr;      /* f32 only: same for the rfft forward / inverse plans of complex length N */
    void *tw_rinv = nullptr;
    uint16_t *perm = nullptr;     /* destination position of X[k] when bitReverseFlag == 0 */
    bool permIsBitrev = false;    /* ... and it is the plain binary bit reversal (every fixed-point preset) */
};
struct Slot {
    bool used = false;
    const void *host[2] = {nullptr, nullptr};   /* the caller's tables this slot was built from */
    uint32_t extra = 0;                         /* further key material (bitRevLength, twidCoefRModifier) */
    uint32_t fp = 0;
    uint64_t hash = 0;
    DevPlan plan;                               /* cfft kinds */
    void *table = nullptr;                      /* rfft twiddles / split coefficients */
};
enum TableKind { TK_PLAN_F32 = 0, TK_PLAN_Q31, TK_PLAN_Q15, TK_PLAN_F64, TK_TWR_F32, TK_TWR_F64, TK_RCOEF_Q31, TK_RCOEF_Q15,
                 TK_R2TW_Q31, TK_R2TW_Q15, TK_WIN_F32, TK_COUNT };
struct DevState {
    Slot slot[TK_COUNT][9][kSlots];
    uint16_t *bitrev[9] = {};     /* plain binary bit reversal of 0..N-1 (cfft_f32 in bit-reversed output order) */
};
static DevState g_dev[kMaxDev];
static std::mutex g_mu;
static thread_local uint8_t t_cur[kMaxDev][TK_COUNT][9];     /* slot the calling thread's transforms use */

static int cur_device(int *dev)
{
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "cudaGetDevice", e);
    if (*dev < 0 || *dev >= kMaxDev) return fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "device index out of range");
    return CMSISDSP_CUDA_OK;
}

/* Find the slot built from these host tables, or claim a free one.  Returns the slot index (>= 0) with *fresh = true
 * when the caller has to build it, or a negative error.  Call with g_mu held. */
static int find_slot(Slot *slots, const void *h0, size_t bytes0, const void *h1, size_t bytes1, uint32_t extra, bool *fresh,
                     uint32_t *fpOut, uint64_t *hashOut)
{
    const uint32_t fp = fingerprint(h0, bytes0);
    *fresh = false;
    for (int s = 0; s < kSlots; s++)
        if (slots[s].used && slots[s].host[0] == h0 && slots[s].host[1] == h1 && slots[s].extra == extra && slots[s].fp == fp) return s;
    uint64_t h = fnv1a(h0, bytes0);
    if (h1 && bytes1) h = fnv1a(h1, bytes1, h);
    h = (h ^ extra) * 1099511628211ull;
    for (int s = 0; s < kSlots; s++)
        if (slots[s].used && slots[s].hash == h) {           /* same content at another address */
            slots[s].host[0] = h0; slots[s].host[1] = h1; slots[s].extra = extra; slots[s].fp = fp;
            return s;
        }
    for (int s = 0; s < kSlots; s++)
        if (!slots[s].used) {
            *fresh = true;
            *fpOut = fp;
            *hashOut = h;
            return s;
        }
    return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "more than 4 distinct tables for one (device, type, length)");
}
static void claim(Slot &sl, const void *h0, const void *h1, uint32_t extra, uint32_t fp, uint64_t hash)
{
    sl.host[0] = h0; sl.host[1] = h1; sl.extra = extra; sl.fp = fp; sl.hash = hash;
    sl.used = true;
}

/* re-order the reference-layout twiddles into the plan's pass order and upload them */
static int upload_pass_ordered(const KernelEntry *ke, const void *base, void **dOut)
{
    const size_t n = ke->twiddles(base, nullptr);
    std::vector<unsigned char> host(n * ke->elemBytes, 0);
    ke->twiddles(base, host.data());
    void *d = nullptr;
    CU_TRY(cudaMalloc(&d, host.size()));
    const cudaError_t e = cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d);
        return fail(CMSISDSP_CUDA_ERR_RUNTIME, "cudaMemcpy (twiddles)", e);
    }
    *dOut = d;
    return CMSISDSP_CUDA_OK;
}
static int upload_raw(const void *host, size_t bytes, void **dOut)
{
    void *d = nullptr;
    CU_TRY(cudaMalloc(&d, bytes));
    const cudaError_t e = cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d);
        return fail(CMSISDSP_CUDA_ERR_RUNTIME, "cudaMemcpy (table)", e);
    }
    *dOut = d;
    return CMSISDSP_CUDA_OK;
}
static void free_plan(DevPlan &p)
{
    cudaFree(p.tw); cudaFree(p.tw_rfwd); cudaFree(p.tw_rinv); cudaFree(p.perm);
    p = DevPlan();
}

/* kernel op of the complex FFT of a data type */
static int cfft_op(int type) { return type == CMSISDSP_CUDA_F64 ? OP_CFFT_F64 : type; }

static int build_tables(int type, int li, const void *base, DevPlan &p)
{
    int rc = upload_pass_ordered(kEntries[cfft_op(type)][li], base, &p.tw);
    if (rc || type != CMSISDSP_CUDA_F32 || !kEntries[OP_RFFT_FWD][li]) return rc;
    rc = upload_pass_ordered(kEntries[OP_RFFT_FWD][li], base, &p.tw_rfwd);
    if (!rc) rc = upload_pass_ordered(kEntries[OP_RFFT_INV][li], base, &p.tw_rinv);
    return rc;
}
static uint32_t bitrev_of(uint32_t k, uint32_t n)
{
    uint32_t r = 0;
    for (uint32_t m = n >> 1; m; m >>= 1, k >>= 1) r = (r << 1) | (k & 1u);
    return r;
}
/* bytes of the reference-layout twiddle table of (type, fftLen) */
static size_t twiddle_bytes(int type, uint32_t n)
{
    switch (type) {
    case CMSISDSP_CUDA_F32: return (size_t)2 * n * sizeof(float);
    case CMSISDSP_CUDA_F64: return (size_t)2 * n * sizeof(double);
    case CMSISDSP_CUDA_Q31: return (size_t)(3 * n / 2) * sizeof(int32_t);
    default: return (size_t)(3 * n / 2) * sizeof(int16_t);
    }
}

extern "C" int cmsisdsp_cuda_plan_upload(int type, uint32_t fftLen, const void *pTwiddle,
                                         const uint16_t *pBitRevTable, uint16_t bitRevLength)
{
    const int li = len_index(fftLen);
    if (type < 0 || type > CMSISDSP_CUDA_F64 || li < 0 || !pTwiddle || (!pBitRevTable && bitRevLength))
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "plan_upload: bad type / length / pointer");
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    Slot *slots = g_dev[dev].slot[TK_PLAN_F32 + type][li];
    bool fresh;
    uint32_t fp = 0;
    uint64_t hash = 0;
    const int s = find_slot(slots, pTwiddle, twiddle_bytes(type, fftLen), pBitRevTable, (size_t)bitRevLength * sizeof(uint16_t),
                            bitRevLength, &fresh, &fp, &hash);
    if (s < 0) return s;
    t_cur[dev][TK_PLAN_F32 + type][li] = (uint8_t)s;
    if (!fresh) return CMSISDSP_CUDA_OK;

    /* out[k] = scrambled[P[k]] after the swap list  =>  X[k] lives at scrambled position P[k] */
    std::vector<uint16_t> perm(fftLen);
    for (uint32_t k = 0; k < fftLen; k++) perm[k] = (uint16_t)k;
    for (uint32_t i = 0; i + 1 < bitRevLength; i += 2) {
        const uint32_t a = pBitRevTable[i] >> 3, b = pBitRevTable[i + 1] >> 3;
        if (a >= fftLen || b >= fftLen) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "plan_upload: bit-reversal entry out of range");
        const uint16_t t = perm[a]; perm[a] = perm[b]; perm[b] = t;
    }
    DevPlan np;
    np.permIsBitrev = true;
    for (uint32_t k = 0; k < fftLen; k++) np.permIsBitrev = np.permIsBitrev && perm[k] == bitrev_of(k, fftLen);
    rc = upload_raw(perm.data(), fftLen * sizeof(uint16_t), (void **)&np.perm);
    if (!rc) rc = build_tables(type, li, pTwiddle, np);
    if (rc) {
        free_plan(np);
        return rc;
    }
    slots[s].plan = np;
    claim(slots[s], pTwiddle, pBitRevTable, bitRevLength, fp, hash);
    return CMSISDSP_CUDA_OK;
}

/* a raw table (rfft twiddles): upload once per content, select for the calling thread */
static int table_upload(int kind, int li, const void *host, size_t bytes)
{
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    Slot *slots = g_dev[dev].slot[kind][li];
    bool fresh;
    uint32_t fp = 0;
    uint64_t hash = 0;
    const int s = find_slot(slots, host, bytes, nullptr, 0, 0, &fresh, &fp, &hash);
    if (s < 0) return s;
    t_cur[dev][kind][li] = (uint8_t)s;
    if (!fresh) return CMSISDSP_CUDA_OK;
    rc = upload_raw(host, bytes, &slots[s].table);
    if (rc) return rc;
    claim(slots[s], host, nullptr, 0, fp, hash);
    return CMSISDSP_CUDA_OK;
}
/* the calling thread's current slot of (kind, li) on the current device; null when nothing was uploaded */
static const Slot *cur_slot(int kind, int li)
{
    int dev;
    if (cur_device(&dev)) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    const Slot *sl = &g_dev[dev].slot[kind][li][t_cur[dev][kind][li]];
    return sl->used ? sl : nullptr;
}

extern "C" int cmsisdsp_cuda_rfft_plan_upload(uint32_t fftLenReal, const float *pTwiddleRFFT)
{
    const int li = len_index(fftLenReal);
    if (li < 1 || !pTwiddleRFFT) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_plan_upload: bad length / pointer");
    return table_upload(TK_TWR_F32, li, pTwiddleRFFT, fftLenReal * sizeof(float));
}

extern "C" int cmsisdsp_cuda_plan_ready(int type, uint32_t fftLen)
{
    const int li = len_index(fftLen);
    if (type < 0 || type > CMSISDSP_CUDA_F64 || li < 0) return 0;
    return cur_slot(TK_PLAN_F32 + type, li) != nullptr;
}
extern "C" int cmsisdsp_cuda_rfft_plan_ready(uint32_t fftLenReal)
{
    const int li = len_index(fftLenReal);
    if (li < 1) return 0;
    return cur_slot(TK_TWR_F32, li) != nullptr && cur_slot(TK_PLAN_F32, li - 1) != nullptr;
}

extern "C" int cmsisdsp_cuda_rfft_fix_plan_upload(int type, uint32_t fftLenReal, const void *pTwiddleAReal,
                                                  const void *pTwiddleBReal, uint32_t twidCoefRModifier)
{
    const int li = len_index(fftLenReal / 2);
    if ((type != CMSISDSP_CUDA_Q31 && type != CMSISDSP_CUDA_Q15) || (fftLenReal & 1u) || li < 0 || !pTwiddleAReal || !pTwiddleBReal ||
        twidCoefRModifier == 0 || (uint64_t)twidCoefRModifier * fftLenReal > 8192u)
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fix_plan_upload: bad type / length / pointer / modifier");
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    const int kind = type == CMSISDSP_CUDA_Q31 ? TK_RCOEF_Q31 : TK_RCOEF_Q15;
    const size_t scalar = type == CMSISDSP_CUDA_Q31 ? 4 : 2;
    /* bin k reads entries 2*k*modifier and 2*k*modifier + 1 of both tables (arm_rfft_q31.c:272-273,333-334):
     * the tables are hashed over the span that is read */
    const uint32_t L2 = fftLenReal / 2;
    const size_t span = ((size_t)2 * (L2 - 1) * twidCoefRModifier + 2) * scalar;
    std::lock_guard<std::mutex> lk(g_mu);
    Slot *slots = g_dev[dev].slot[kind][li];
    bool fresh;
    uint32_t fp = 0;
    uint64_t hash = 0;
    const int s = find_slot(slots, pTwiddleAReal, span, pTwiddleBReal, span, twidCoefRModifier, &fresh, &fp, &hash);
    if (s < 0) return s;
    t_cur[dev][kind][li] = (uint8_t)s;
    if (!fresh) return CMSISDSP_CUDA_OK;
    std::vector<int32_t> coef((size_t)L2 * 4);
    for (uint32_t k = 0; k < L2; k++) {
        const size_t e = (size_t)2 * k * twidCoefRModifier;
        if (type == CMSISDSP_CUDA_Q31) {
            const int32_t *A = (const int32_t *)pTwiddleAReal, *B = (const int32_t *)pTwiddleBReal;
            coef[4 * k] = A[e]; coef[4 * k + 1] = A[e + 1]; coef[4 * k + 2] = B[e]; coef[4 * k + 3] = B[e + 1];
        } else {
            const int16_t *A = (const int16_t *)pTwiddleAReal, *B = (const int16_t *)pTwiddleBReal;
            coef[4 * k] = A[e]; coef[4 * k + 1] = A[e + 1]; coef[4 * k + 2] = B[e]; coef[4 * k + 3] = B[e + 1];
        }
    }
    rc = upload_raw(coef.data(), coef.size() * sizeof(int32_t), &slots[s].table);
    if (rc) return rc;
    claim(slots[s], pTwiddleAReal, pTwiddleBReal, twidCoefRModifier, fp, hash);
    return CMSISDSP_CUDA_OK;
}
extern "C" int cmsisdsp_cuda_rfft_fix_plan_ready(int type, uint32_t fftLenReal)
{
    const int li = len_index(fftLenReal / 2);
    if ((type != CMSISDSP_CUDA_Q31 && type != CMSISDSP_CUDA_Q15) || li < 0) return 0;
    return cur_slot(type == CMSISDSP_CUDA_Q31 ? TK_RCOEF_Q31 : TK_RCOEF_Q15, li) != nullptr && cur_slot(TK_PLAN_F32 + type, li) != nullptr;
}

static int get_plan(int type, uint32_t fftLen, DevPlan *out)
{
    const int li = len_index(fftLen);
    if (li < 0) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported fftLen (16..4096, power of two)");
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    const Slot *sl = cur_slot(TK_PLAN_F32 + type, li);
    if (!sl) return fail(CMSISDSP_CUDA_ERR_NO_PLAN, "no plan uploaded for this (device, type, fftLen)");
    *out = sl->plan;
    return CMSISDSP_CUDA_OK;
}
static int get_table(int kind, int li, const void **out, const char *what)
{
    const Slot *sl = cur_slot(kind, li);
    if (!sl) return fail(CMSISDSP_CUDA_ERR_NO_PLAN, what);
    *out = sl->table;
    return CMSISDSP_CUDA_OK;
}

namespace b200fft {
int shim_rfft_tables(uint32_t fftLenReal, const void **twForward, const float **twRfft)
{
    const int li = len_index(fftLenReal);
    if (li < 1) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported rfft length (32..4096, power of two)");
    DevPlan pl;
    int rc = get_plan(CMSISDSP_CUDA_F32, fftLenReal / 2, &pl);
    if (rc) return rc;
    const void *twr = nullptr;
    rc = get_table(TK_TWR_F32, li, &twr, "no rfft plan uploaded for this (device, fftLen)");
    if (rc) return rc;
    *twForward = pl.tw_rfwd;
    *twRfft = (const float *)twr;
    return CMSISDSP_CUDA_OK;
}
}

/* ------------------------------------------------------------------ transforms */

/* Device buffers must be aligned to one complex element (the kernels move whole complex points: 8 bytes for
 * f32 / q31, 4 for q15, 16 for f64).  The reference takes any scalar-aligned buffer; the C front library's
 * host-pointer path always satisfies this (its staging buffers are cudaMalloc'ed), a device-pointer caller
 * with an odd scalar offset gets ERR_ARGUMENT instead of a misaligned-address fault. */
static int elem_align(int type) { return type == CMSISDSP_CUDA_F64 ? 16 : (type == CMSISDSP_CUDA_Q15 ? 4 : 8); }
static bool misaligned(const void *p, int a) { return ((uintptr_t)p & (uintptr_t)(a - 1)) != 0; }

enum OutputOrder { ORDER_NATURAL = 0, ORDER_INSTANCE = 1, ORDER_BITREV = 2 };
/* plain binary bit reversal of 0..N-1 on the current device (built on first use) */
static int bitrev_table(int li, const uint16_t **out)
{
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_dev[dev].bitrev[li]) {
        const uint32_t n = kLens[li];
        std::vector<uint16_t> t(n);
        for (uint32_t k = 0; k < n; k++) t[k] = (uint16_t)bitrev_of(k, n);
        rc = upload_raw(t.data(), n * sizeof(uint16_t), (void **)&g_dev[dev].bitrev[li]);
        if (rc) return rc;
    }
    *out = g_dev[dev].bitrev[li];
    return CMSISDSP_CUDA_OK;
}

/* d_in == d_p: in place; otherwise frames are read from d_in and written to d_p (the direct kernels never read a
 * frame after they started to write it, so the two may also be distinct buffers).
 * order: ORDER_NATURAL (bitReverseFlag = 1), ORDER_INSTANCE (bitReverseFlag = 0: the order the instance's own swap list
 * leaves behind), ORDER_BITREV (plain binary bit-reversed order: the deprecated radix-4 / radix-2 f32 functions with
 * bitReverseFlag = 0, arm_cfft_radix4_f32.c:81, arm_cfft_radix2_f32.c) */
static int cfft_io(int type, const void *d_in, void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, int order, void *stream)
{
    if ((!d_p || !d_in) && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    DevPlan pl;
    int rc = get_plan(type, fftLen, &pl);
    if (rc) return rc;
    const int li = len_index(fftLen);
    /* N = 2*4^m: final << 1 (fixed point only; arm_cfft_q31.c:803-820, arm_cfft_q15.c:810-827) */
    const int shl1 = (type == CMSISDSP_CUDA_Q31 || type == CMSISDSP_CUDA_Q15) ? ((li + 4) & 1) : 0;
    if (misaligned(d_p, elem_align(type)) || misaligned(d_in, elem_align(type)))
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "cfft: device data must be aligned to one complex element (8 bytes f32/q31, 4 q15, 16 f64)");
    const uint16_t *perm = nullptr;
    if (order == ORDER_INSTANCE) perm = pl.perm;
    else if (order == ORDER_BITREV && (rc = bitrev_table(li, &perm))) return rc;
    const KernelEntry *ke = kEntries[cfft_op(type)][li];
    return ke->launch(d_in, d_p, nFrames, ifftFlag == 1, pl.tw, perm, nullptr, shl1, choose_flavour(ke), (cudaStream_t)stream);
}
static int cfft_any(int type, void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{
    return cfft_io(type, d_p, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag ? ORDER_NATURAL : ORDER_INSTANCE, stream);
}

extern "C" int cmsisdsp_cuda_cfft_f32(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{ return cfft_any(CMSISDSP_CUDA_F32, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag, stream); }
extern "C" int cmsisdsp_cuda_cfft_q31(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{ return cfft_any(CMSISDSP_CUDA_Q31, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag, stream); }
extern "C" int cmsisdsp_cuda_cfft_q15(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{ return cfft_any(CMSISDSP_CUDA_Q15, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag, stream); }
extern "C" int cmsisdsp_cuda_cfft_f64(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, uint8_t bitReverseFlag, void *stream)
{ return cfft_any(CMSISDSP_CUDA_F64, d_p, fftLen, nFrames, ifftFlag, bitReverseFlag, stream); }
extern "C" int cmsisdsp_cuda_cfft_f32_bitrev_order(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{ return cfft_io(CMSISDSP_CUDA_F32, d_p, d_p, fftLen, nFrames, ifftFlag, ORDER_BITREV, stream); }

extern "C" int cmsisdsp_cuda_rfft_fast_f32(const void *d_p, void *d_out, uint32_t fftLenReal, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{
    if ((!d_p || !d_out) && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    if (d_p == d_out && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fast: p and pOut must not alias");
    if (misaligned(d_p, 8) || misaligned(d_out, 8)) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fast: device data must be 8-byte aligned");
    const int li = len_index(fftLenReal);
    if (li < 1) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported rfft length (32..4096, power of two)");
    DevPlan pl;
    int rc = get_plan(CMSISDSP_CUDA_F32, fftLenReal / 2, &pl);
    if (rc) return rc;
    const void *twr = nullptr;
    rc = get_table(TK_TWR_F32, li, &twr, "no rfft plan uploaded for this (device, fftLen)");
    if (rc) return rc;
    const KernelEntry *ke = kEntries[ifftFlag ? OP_RFFT_INV : OP_RFFT_FWD][li - 1];
    return ke->launch(d_p, d_out, nFrames, ifftFlag != 0, ifftFlag ? pl.tw_rinv : pl.tw_rfwd, twr, nullptr, 0, choose_flavour(ke),
                      (cudaStream_t)stream);
}

/* ------------------------------------------------------------------ arm_rfft_fast_f64 (arm_rfft_fast_f64.c:207-233)
 *
 * One fused kernel per direction (fft_body.cuh: RfftF64FwdBody / RfftF64InvBody; units ku_11_L / ku_12_L): the L-point
 * f64 CFFT with stage_rfft_f64 (:30-118) as its epilogue, merge_rfft_f64 (:121-181) as the load of the inverse.  The
 * first form of this entry point was an adapter -- complex kernel, then a stage kernel in place, two passes over HBM:
 * 45-52 % of the HBM peak (profiles/r1_f_notes.md). */
extern "C" int cmsisdsp_cuda_rfft_f64_plan_upload(uint32_t fftLenReal, const double *pTwiddleRFFT)
{
    const int li = len_index(fftLenReal);
    if (li < 1 || !pTwiddleRFFT) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_f64_plan_upload: bad length / pointer");
    return table_upload(TK_TWR_F64, li, pTwiddleRFFT, fftLenReal * sizeof(double));
}
extern "C" int cmsisdsp_cuda_rfft_f64_plan_ready(uint32_t fftLenReal)
{
    const int li = len_index(fftLenReal);
    if (li < 1) return 0;
    return cur_slot(TK_TWR_F64, li) != nullptr && cur_slot(TK_PLAN_F64, li - 1) != nullptr;
}

extern "C" int cmsisdsp_cuda_rfft_fast_f64(const void *d_p, void *d_out, uint32_t fftLenReal, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{
    if ((!d_p || !d_out) && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    if (d_p == d_out && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fast_f64: p and pOut must not alias");
    if (((uintptr_t)d_p | (uintptr_t)d_out) & 15u) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fast_f64: data must be 16-byte aligned");
    const int li = len_index(fftLenReal);
    if (li < 1) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported rfft length (32..4096, power of two)");
    const void *twr = nullptr;
    int rc = get_table(TK_TWR_F64, li, &twr, "no f64 rfft plan uploaded for this (device, fftLen)");
    if (rc) return rc;
    DevPlan pl;
    rc = get_plan(CMSISDSP_CUDA_F64, fftLenReal / 2, &pl);
    if (rc) return rc;
    const KernelEntry *ke = kEntries[ifftFlag ? OP_RFFT_F64_INV : OP_RFFT_F64_FWD][li - 1];
    return ke->launch(d_p, d_out, nFrames, ifftFlag != 0, pl.tw, twr, nullptr, 0, KF_DIRECT, (cudaStream_t)stream);
}

/* arm_cfft_f32 + spectrum epilogue: mode 0 magnitudes, 1 squared magnitudes (d_out: fftLen floats per frame),
 * 2 peak (d_out: one value, d_aux: one uint32 index per frame) */
static int cfft_spectrum(const void *d_src, void *d_out, void *d_aux, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, int mode, void *stream)
{
    if ((!d_src || !d_out || (mode == 2 && !d_aux)) && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    if (d_src == d_out && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "spectrum epilogue: source and destination must not alias");
    if (misaligned(d_src, 8) || misaligned(d_out, 4) || misaligned(d_aux, 4))
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "spectrum epilogue: source must be 8-byte, destinations 4-byte aligned");
    DevPlan pl;
    int rc = get_plan(CMSISDSP_CUDA_F32, fftLen, &pl);
    if (rc) return rc;
    const KernelEntry *ke = kEntries[OP_CFFT_MAG_F32][len_index(fftLen)];
    return ke->launch(d_src, d_out, nFrames, ifftFlag == 1, pl.tw, d_aux, nullptr, mode, choose_flavour(ke), (cudaStream_t)stream);
}
extern "C" int cmsisdsp_cuda_cfft_mag_f32(const void *d_src, void *d_mag, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag,
                                          uint8_t squared, void *stream)
{ return cfft_spectrum(d_src, d_mag, nullptr, fftLen, nFrames, ifftFlag, squared ? 1 : 0, stream); }
extern "C" int cmsisdsp_cuda_cfft_peak_f32(const void *d_src, void *d_val, void *d_idx, uint32_t fftLen, uint64_t nFrames,
                                           uint8_t ifftFlag, void *stream)
{ return cfft_spectrum(d_src, d_val, d_idx, fftLen, nFrames, ifftFlag, 2, stream); }

/* bitReverseFlagR = 0 (arm_rfft_q31.c:164,173 hand the flag to arm_cfft_q31): the complex transform inside leaves its
 * result in the instance's unordered layout -- forward, the split stage then runs over that layout as if it were
 * natural order; inverse, the result stays unordered.  Supported for instances whose unordered layout is the plain
 * binary bit reversal (every preset). */
static int rfft_fix(int type, const void *d_src, void *d_dst, uint32_t fftLenReal, uint64_t nFrames, uint8_t ifftFlagR,
                    uint8_t bitReverseFlagR, void *stream)
{
    if ((!d_src || !d_dst) && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    if (d_src == d_dst && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft: pSrc and pDst must not alias");
    if (misaligned(d_src, elem_align(type)) || misaligned(d_dst, elem_align(type)))
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft: device data must be aligned to one complex element (8 bytes q31, 4 q15)");
    const int li = (fftLenReal & 1u) ? -1 : len_index(fftLenReal / 2);
    if (li < 0) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported fixed-point rfft length (32..8192, power of two)");
    DevPlan pl;
    int rc = get_plan(type, fftLenReal / 2, &pl);
    if (rc) return rc;
    const void *coef = nullptr;
    rc = get_table(type == CMSISDSP_CUDA_Q31 ? TK_RCOEF_Q31 : TK_RCOEF_Q15, li, &coef,
                   "no fixed-point rfft plan uploaded for this (device, type, fftLenReal)");
    if (rc) return rc;
    if (!bitReverseFlagR && !pl.permIsBitrev)
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft with bitReverseFlagR = 0 needs the standard (binary) bit-reversal table");
    const int op = (type == CMSISDSP_CUDA_Q31 ? OP_RFFT_Q31_FWD : OP_RFFT_Q15_FWD) + (ifftFlagR ? 1 : 0);
    const KernelEntry *ke = kEntries[op][li];
    return ke->launch(d_src, d_dst, nFrames, ifftFlagR != 0, pl.tw, coef, bitReverseFlagR ? nullptr : pl.perm, (li + 4) & 1,
                      choose_flavour(ke), (cudaStream_t)stream);
}
extern "C" int cmsisdsp_cuda_rfft_q31(const void *d_src, void *d_dst, uint32_t fftLenReal, uint64_t nFrames, uint8_t ifftFlagR,
                                      uint8_t bitReverseFlagR, void *stream)
{ return rfft_fix(CMSISDSP_CUDA_Q31, d_src, d_dst, fftLenReal, nFrames, ifftFlagR, bitReverseFlagR, stream); }
extern "C" int cmsisdsp_cuda_rfft_q15(const void *d_src, void *d_dst, uint32_t fftLenReal, uint64_t nFrames, uint8_t ifftFlagR,
                                      uint8_t bitReverseFlagR, void *stream)
{ return rfft_fix(CMSISDSP_CUDA_Q15, d_src, d_dst, fftLenReal, nFrames, ifftFlagR, bitReverseFlagR, stream); }

/* ------------------------------------------------------------------ window multiply fused into the load (f32)
 * arm_mult_f32(pSrc, window, pSrc, fftLen) + arm_rfft_fast_f32 (as arm_mfcc_f32.c:112,137 does), or a real window over
 * the complex samples of arm_cfft_f32: the windowed frame never exists in memory. */
extern "C" int cmsisdsp_cuda_window_upload(uint32_t length, const float *pWindow)
{
    const int li = len_index(length);
    if (li < 0 || !pWindow) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "window_upload: bad length / pointer");
    return table_upload(TK_WIN_F32, li, pWindow, length * sizeof(float));
}
extern "C" int cmsisdsp_cuda_cfft_window_f32(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{
    if (!d_p && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    if (misaligned(d_p, 8)) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "cfft_window: device data must be 8-byte aligned");
    DevPlan pl;
    int rc = get_plan(CMSISDSP_CUDA_F32, fftLen, &pl);
    if (rc) return rc;
    const int li = len_index(fftLen);
    const void *win = nullptr;
    rc = get_table(TK_WIN_F32, li, &win, "no window uploaded for this (device, length)");
    if (rc) return rc;
    const KernelEntry *ke = kEntries[OP_CFFT_F32][li];
    return ke->launch(d_p, d_p, nFrames, ifftFlag == 1, pl.tw, nullptr, win, 0, KF_DIRECT, (cudaStream_t)stream);
}
extern "C" int cmsisdsp_cuda_rfft_fast_window_f32(const void *d_p, void *d_out, uint32_t fftLenReal, uint64_t nFrames, void *stream)
{
    if ((!d_p || !d_out) && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    if (d_p == d_out && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fast_window: p and pOut must not alias");
    if (misaligned(d_p, 8) || misaligned(d_out, 8)) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "rfft_fast_window: device data must be 8-byte aligned");
    const int li = len_index(fftLenReal);
    if (li < 1) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported rfft length (32..4096, power of two)");
    DevPlan pl;
    int rc = get_plan(CMSISDSP_CUDA_F32, fftLenReal / 2, &pl);
    if (rc) return rc;
    const void *twr = nullptr, *win = nullptr;
    rc = get_table(TK_TWR_F32, li, &twr, "no rfft plan uploaded for this (device, fftLen)");
    if (!rc) rc = get_table(TK_WIN_F32, li, &win, "no window uploaded for this (device, length)");
    if (rc) return rc;
    const KernelEntry *ke = kEntries[OP_RFFT_FWD][li - 1];
    return ke->launch(d_p, d_out, nFrames, 0, pl.tw_rfwd, twr, win, 0, KF_DIRECT, (cudaStream_t)stream);
}

/* ------------------------------------------------------------------ deprecated fixed-point radix-2 (radix2_fix.cu) */
namespace b200fft { int shim_radix2_launch(int type, void *d_p, uint32_t N, uint64_t nFrames, int inv, const void *tw, cudaStream_t st); }

/* pCoef / twidCoefModifier as in arm_cfft_radix2_instance_q31 / _q15: entry k * modifier of the table is W_fftLen^k */
extern "C" int cmsisdsp_cuda_radix2_plan_upload(int type, uint32_t fftLen, const void *pCoef, uint32_t twidCoefModifier)
{
    const int li = len_index(fftLen);
    if ((type != CMSISDSP_CUDA_Q31 && type != CMSISDSP_CUDA_Q15) || li < 0 || !pCoef || twidCoefModifier == 0 ||
        (uint64_t)twidCoefModifier * fftLen > 4096u)
        return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "radix2_plan_upload: bad type / length / pointer / modifier");
    int dev;
    int rc = cur_device(&dev);
    if (rc) return rc;
    const int kind = type == CMSISDSP_CUDA_Q31 ? TK_R2TW_Q31 : TK_R2TW_Q15;
    const size_t scalar = type == CMSISDSP_CUDA_Q31 ? 4 : 2;
    const uint32_t half = fftLen / 2;
    const size_t span = ((size_t)2 * (half - 1) * twidCoefModifier + 2) * scalar;      /* the part of the table that is read */
    std::lock_guard<std::mutex> lk(g_mu);
    Slot *slots = g_dev[dev].slot[kind][li];
    bool fresh;
    uint32_t fp = 0;
    uint64_t hash = 0;
    const int s = find_slot(slots, pCoef, span, nullptr, 0, twidCoefModifier, &fresh, &fp, &hash);
    if (s < 0) return s;
    t_cur[dev][kind][li] = (uint8_t)s;
    if (!fresh) return CMSISDSP_CUDA_OK;
    std::vector<int32_t> tw((size_t)half * 2);
    for (uint32_t k = 0; k < half; k++) {
        const size_t e = (size_t)2 * k * twidCoefModifier;
        if (type == CMSISDSP_CUDA_Q31) { tw[2 * k] = ((const int32_t *)pCoef)[e]; tw[2 * k + 1] = ((const int32_t *)pCoef)[e + 1]; }
        else                           { tw[2 * k] = ((const int16_t *)pCoef)[e]; tw[2 * k + 1] = ((const int16_t *)pCoef)[e + 1]; }
    }
    rc = upload_raw(tw.data(), tw.size() * sizeof(int32_t), &slots[s].table);
    if (rc) return rc;
    claim(slots[s], pCoef, nullptr, twidCoefModifier, fp, hash);
    return CMSISDSP_CUDA_OK;
}
static int radix2_fix(int type, void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{
    if (!d_p && nFrames) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "null data pointer");
    const int li = len_index(fftLen);
    if (li < 0) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "unsupported fftLen (16..4096, power of two)");
    if (misaligned(d_p, elem_align(type))) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "cfft_radix2: device data must be aligned to one complex element");
    const void *tw = nullptr;
    int rc = get_table(type == CMSISDSP_CUDA_Q31 ? TK_R2TW_Q31 : TK_R2TW_Q15, li, &tw, "no radix-2 plan uploaded for this (device, type, fftLen)");
    if (rc) return rc;
    if ((size_t)fftLen * 8 > 48 * 1024) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "cfft_radix2: frame does not fit the kernel's shared memory");
    return shim_radix2_launch(type, d_p, fftLen, nFrames, ifftFlag == 1, tw, (cudaStream_t)stream);
}
extern "C" int cmsisdsp_cuda_cfft_radix2_q31(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{ return radix2_fix(CMSISDSP_CUDA_Q31, d_p, fftLen, nFrames, ifftFlag, stream); }
extern "C" int cmsisdsp_cuda_cfft_radix2_q15(void *d_p, uint32_t fftLen, uint64_t nFrames, uint8_t ifftFlag, void *stream)
{ return radix2_fix(CMSISDSP_CUDA_Q15, d_p, fftLen, nFrames, ifftFlag, stream); }

extern "C" int cmsisdsp_cuda_kernel_info(int op, uint32_t fftLen, int *threads, int *frames, int *smem, int *regs, int *ctasPerSm)
{
    const int li = len_index(((op >= 3 && op <= 8) || op == OP_RFFT_F64_FWD || op == OP_RFFT_F64_INV) ? fftLen / 2 : fftLen);   /* real length */
    if (op < 0 || op >= OP_COUNT || li < 0 || !kEntries[op][li]) return fail(CMSISDSP_CUDA_ERR_ARGUMENT, "kernel_info: unsupported (op, fftLen)");
    KernelFacts f;
    int rc = kEntries[op][li]->facts(&f, choose_flavour(kEntries[op][li]));
    if (rc) return rc;
    if (threads) *threads = f.threads;
    if (frames) *frames = f.frames;
    if (smem) *smem = f.smem;
    if (regs) *regs = f.regs;
    if (ctasPerSm) *ctasPerSm = f.ctasPerSm;
    return CMSISDSP_CUDA_OK;
}

/* ------------------------------------------------------------------ plumbing */

extern "C" int cmsisdsp_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
extern "C" int cmsisdsp_cuda_set_device(int device) { CU_TRY(cudaSetDevice(device)); return 0; }
extern "C" int cmsisdsp_cuda_get_device(void)
{
    int d = -1;
    const cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess) {
        fail(CMSISDSP_CUDA_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback): cudaGetDevice", e);
        return -1;
    }
    return d;
}
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

/* device ordinal that owns ptr (device or managed memory), -1 for host memory, < -1 on error */
extern "C" int cmsisdsp_cuda_pointer_device(const void *ptr)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, ptr);
    if (e != cudaSuccess) { (void)cudaGetLastError(); fail(CMSISDSP_CUDA_ERR_RUNTIME, "cudaPointerGetAttributes", e); return -2; }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
}
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
