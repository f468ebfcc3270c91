/*
 * arm_cuda_engine.h -- internal to libcmsisdsp_b200: how a batched call reaches the device(s).
 *
 * Every exec function of the library describes its work as a JOB -- per-frame byte counts, where the frames
 * live, and three callbacks that upload the instance's tables and enqueue the kernel(s) -- and hands it to
 *   arm_cuda_run_device   both buffers are device memory: the job runs on the device that owns them, in place of
 *                         the buffers, on the library's stream; the call returns when the result is there
 *   arm_cuda_run_host     host buffers: the frame range is block-partitioned over the device list (SURVEY 8(e):
 *                         device g of G gets frames [g*ceil(B/G), min(B, (g+1)*ceil(B/G))) ), one host thread per
 *                         device, each streaming its block through device staging buffers in chunks on several
 *                         streams (copy-in / kernel / copy-out of consecutive chunks overlap)
 * The reference has no counterpart (it has no threads and no devices): frames are independent, nothing is
 * exchanged between devices.
 */
#ifndef ARM_CUDA_ENGINE_H
#define ARM_CUDA_ENGINE_H

#include <stddef.h>
#include <stdint.h>
#include "arm_math_types.h"

typedef struct arm_cuda_job arm_cuda_job;
struct arm_cuda_job {
    /* frame f of the input starts at in + f*inStride and is inFrame bytes long (inFrame > inStride: overlapping
     * frames, arm_mfcc_batch_f32 with hop < fftLen); frame f of the output likewise */
    size_t inStride, inFrame, outStride, outFrame;
    /* in place: the kernel transforms the input buffer itself (out == in, outStride == inStride) */
    int inPlace;
    /* a second output array written by the same kernel (peak pick: values + indices); NULL when unused */
    char *outB;
    size_t outBStride, outBFrame;
    /* a second result copied home from the INPUT staging buffer after `post` ran on it (the legacy single-frame real
     * FFTs leave the complex transform in the source buffer): out2 == NULL when unused */
    char *out2;
    size_t out2Stride, out2Frame;
    /* make the tables of the instance resident on the CURRENT device and select them for the calling thread */
    int (*prepare)(const arm_cuda_job *job);
    /* enqueue the transform of n frames: din / dout / doutB are device buffers laid out like the host ones */
    int (*launch)(const arm_cuda_job *job, const void *din, void *dout, void *doutB, uint64_t n, void *stream);
    /* optional: enqueue the second kernel (on din, in place) whose result goes to out2 */
    int (*post)(const arm_cuda_job *job, void *din, uint64_t n, void *stream);
    const void *self;            /* the exec function's own arguments */
};

/* in / out: host pointers (out == in for in-place jobs) */
arm_status arm_cuda_run_host(const arm_cuda_job *job, const void *in, void *out, uint64_t nFrames);
arm_status arm_cuda_run_device(const arm_cuda_job *job, const void *in, void *out, uint64_t nFrames);
/* host or device, decided by the pointers (both must be of the same kind) */
arm_status arm_cuda_run(const arm_cuda_job *job, const void *in, void *out, uint64_t nFrames);

/* shim return code -> arm_status (CMSISDSP_CUDA_OK -> ARM_MATH_SUCCESS, ...) */
arm_status arm_cuda_status_of(int shimRc);
arm_status arm_cuda_set_last_status(arm_status s);

#endif
