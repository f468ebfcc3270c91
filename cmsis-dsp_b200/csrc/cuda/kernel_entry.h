/*
 * kernel_entry.h -- internal interface between the C-ABI shim (cmsisdsp_cuda.cu) and the
 * per-(op, length) kernel objects (kernel_unit.cu compiled once per pair, in parallel).
 * Not part of the public ABI.
 */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200fft {

enum KernelOp { OP_CFFT_F32 = 0, OP_CFFT_Q31 = 1, OP_CFFT_Q15 = 2, OP_RFFT_FWD = 3, OP_RFFT_INV = 4,
                OP_RFFT_Q31_FWD = 5, OP_RFFT_Q31_INV = 6, OP_RFFT_Q15_FWD = 7, OP_RFFT_Q15_INV = 8, OP_CFFT_MAG_F32 = 9,
                OP_CFFT_F64 = 10, OP_RFFT_F64_FWD = 11, OP_RFFT_F64_INV = 12, OP_COUNT = 13 };

struct KernelFacts { int threads, frames, smem, regs, ctasPerSm; };

/* kernel flavours: KF_DIRECT = one CTA per frame group, loads straight into registers;
 * KF_PIPE = persistent CTAs fed by bulk async copies (two-pass f32 plans only; falls back to
 * KF_DIRECT when the flavour does not exist for the pair or the input is not 16-byte aligned) */
enum KernelFlavour { KF_DIRECT = 0, KF_PIPE = 1 };

struct KernelEntry {
    /* cfft: in == out (in place), aux = output permutation (uint16, may be null), inv = ifftFlag,
     *       shl1 = final << 1 (fixed point, N = 2*4^m)
     * rfft: in -> out, aux = twiddleCoef_rfft table (device); direction is fixed by the op
     * rfft q31/q15: in -> out, tw = twiddles of the cfft plan of the same type and complex length,
     *       aux = split coefficients (ci32x4 per bin), shl1 as for cfft
     * cfft + spectrum epilogue (f32): in -> out = magnitudes, or out = peak values and aux = peak indices; tw = the
     *       cfft plan's twiddles; the shl1 slot carries the SpectrumMode (0 mag, 1 mag squared, 2 peak) */
    /* aux2: rfft q31/q15 only -- the unordered layout of the complex transform (uint16 positions) when
     *       bitReverseFlagR == 0, else null */
    int (*launch)(const void *in, void *out, uint64_t nFrames, int inv, const void *tw, const void *aux, const void *aux2, int shl1,
                  int flavour, cudaStream_t st);
    /* number of elements of the pass-ordered twiddle table (+1 pad); fills hostOut when non-null */
    size_t (*twiddles)(const void *base, void *hostOut);
    size_t elemBytes;           /* bytes per entry of that table (Arith::telem) */
    int (*facts)(KernelFacts *out, int flavour);
    bool hasPipe;               /* a KF_PIPE flavour exists for this pair */
    bool preferPipe;            /* ... and measured faster than KF_DIRECT */
};

/* defined in cmsisdsp_cuda.cu */
int shim_fail(int code, const char *what, cudaError_t e);
void shim_count_launch();
/* device tables of the forward rfft plan of real length fftLenReal on the current device
 * (CMSISDSP_CUDA_ERR_NO_PLAN when cmsisdsp_cuda_plan_upload / _rfft_plan_upload have not run) */
int shim_rfft_tables(uint32_t fftLenReal, const void **twForward, const float **twRfft);
/* flavour forced by cmsisdsp_cuda_set_kernel_flavour / CMSISDSP_CUDA_KERNEL: KF_DIRECT, KF_PIPE or -1 (per-kernel default) */
int shim_forced_flavour();

}  // namespace b200fft
