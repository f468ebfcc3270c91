/*
 * arm_cuda_engine.c -- runs a batched job (arm_cuda_engine.h) on one device (device buffers) or fans it out over the
 * device list (host buffers).  Pure C over the C ABI of libcmsisdsp_cuda.
 *
 * Device list: arm_cuda_set_devices(), else the environment variable CMSISDSP_CUDA_DEVICES ("all" or a comma list of
 * ordinals; an ordinal may repeat: two workers then share that device), else every visible device with the calling
 * thread's current device first.  A call uses as many workers as its size warrants (at least MIN_BYTES_PER_WORKER per
 * worker), so a legacy single-frame call stays on the current device.
 * Staging: chunks of CMSISDSP_CUDA_CHUNK_MIB MiB (default 32) cycling over CMSISDSP_CUDA_NSTREAMS streams (default 3)
 * per worker; arm_cuda_set_staging() overrides both.  Chunk size and stream count were swept on a B200 (profiles/
 * r2_staging_sweep.txt).  A call's first chunks grow geometrically from CMSISDSP_CUDA_RAMP_MIB MiB (default 4; 0: every
 * chunk full-sized) and its last ones shrink the same way: a call takes the time of all its copies in one direction plus
 * the FIRST chunk's copy in and the LAST chunk's copy out, which nothing overlaps (arm_cuda_set_staging_ramp()).
 * Streams and staging buffers belong to the calling host thread (one set per worker slot); they are released when the
 * thread exits or calls arm_cuda_release().
 */
#include "arm_cuda_engine.h"
#include "dsp/transform_functions.h"
#include "cmsisdsp_cuda.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAXW 16                      /* workers of one call */
#define MAXS 8                       /* streams per worker */
#define NBUF 3                       /* staging buffers per stream: input, output, second output */
#define MIN_BYTES_PER_WORKER ((size_t)8 << 20)
#define DEFAULT_CHUNK ((size_t)32 << 20)
#define DEFAULT_NSTREAM 3
#define DEFAULT_RAMP ((size_t)4 << 20)

typedef struct {
    int device;                      /* -1: unused */
    int nstream;
    void *stream[MAXS];
    void *buf[MAXS][NBUF];
    size_t cap[MAXS][NBUF];
} dev_ctx;
typedef struct { dev_ctx w[MAXW]; } ctx_pool;

static struct {
    pthread_mutex_t mu;
    int inited;
    int ndev;                        /* 0: default (all visible devices, current first) */
    int dev[MAXW];
    size_t chunk;
    int nstream;
    size_t ramp;                     /* size of a call's first and last chunk; 0: no ramp */
} g_cfg = { PTHREAD_MUTEX_INITIALIZER, 0, 0, {0}, DEFAULT_CHUNK, DEFAULT_NSTREAM, DEFAULT_RAMP };

static pthread_key_t g_key;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static __thread arm_status g_last = ARM_MATH_SUCCESS;

arm_status arm_cuda_last_status(void) { return g_last; }
arm_status arm_cuda_set_last_status(arm_status s) { g_last = s; return s; }

arm_status arm_cuda_status_of(int rc)
{
    switch (rc) {
    case CMSISDSP_CUDA_OK: return ARM_MATH_SUCCESS;
    case CMSISDSP_CUDA_ERR_ARGUMENT: return ARM_MATH_ARGUMENT_ERROR;
    case CMSISDSP_CUDA_ERR_NO_PLAN: return ARM_MATH_CUDA_NO_PLAN;
    case CMSISDSP_CUDA_ERR_NO_DEVICE: return ARM_MATH_CUDA_NO_DEVICE;
    default: return ARM_MATH_CUDA_RUNTIME_ERROR;
    }
}

/* ------------------------------------------------------------------ configuration */

static void cfg_init_locked(void)
{
    if (g_cfg.inited) return;
    g_cfg.inited = 1;
    const char *e = getenv("CMSISDSP_CUDA_DEVICES");
    if (e && strcmp(e, "all") != 0 && g_cfg.ndev == 0) {
        int n = 0;
        while (*e && n < MAXW) {
            char *end;
            long v = strtol(e, &end, 10);
            if (end == e) break;
            if (v >= 0 && v < 1024) g_cfg.dev[n++] = (int)v;
            e = (*end == ',') ? end + 1 : end;
        }
        g_cfg.ndev = n;
    }
    e = getenv("CMSISDSP_CUDA_CHUNK_MIB");
    if (e && atol(e) > 0 && atol(e) <= 4096) g_cfg.chunk = (size_t)atol(e) << 20;
    e = getenv("CMSISDSP_CUDA_NSTREAMS");
    if (e && atoi(e) >= 1 && atoi(e) <= MAXS) g_cfg.nstream = atoi(e);
    e = getenv("CMSISDSP_CUDA_RAMP_MIB");
    if (e && atol(e) >= 0 && atol(e) <= 4096) g_cfg.ramp = (size_t)atol(e) << 20;
}

arm_status arm_cuda_set_devices(const int32_t *devices, uint32_t nDevices)
{
    if (nDevices > MAXW || (nDevices && !devices)) return ARM_MATH_ARGUMENT_ERROR;
    const int visible = cmsisdsp_cuda_device_count();
    for (uint32_t i = 0; i < nDevices; i++)
        if (devices[i] < 0 || devices[i] >= visible) return visible ? ARM_MATH_ARGUMENT_ERROR : ARM_MATH_CUDA_NO_DEVICE;
    pthread_mutex_lock(&g_cfg.mu);
    cfg_init_locked();
    g_cfg.ndev = (int)nDevices;
    for (uint32_t i = 0; i < nDevices; i++) g_cfg.dev[i] = devices[i];
    pthread_mutex_unlock(&g_cfg.mu);
    return ARM_MATH_SUCCESS;
}

/* the devices host-pointer calls fan out over, in order; returns how many (0: no device) */
static int device_list(int *out)
{
    pthread_mutex_lock(&g_cfg.mu);
    cfg_init_locked();
    int n = g_cfg.ndev;
    for (int i = 0; i < n; i++) out[i] = g_cfg.dev[i];
    pthread_mutex_unlock(&g_cfg.mu);
    if (n > 0) return n;
    const int visible = cmsisdsp_cuda_device_count();
    if (visible <= 0) return 0;
    int cur = cmsisdsp_cuda_get_device();
    if (cur < 0) cur = 0;
    n = visible < MAXW ? visible : MAXW;
    for (int i = 0; i < n; i++) out[i] = (cur + i) % visible;
    return n;
}

uint32_t arm_cuda_get_devices(int32_t *devices, uint32_t maxDevices)
{
    int d[MAXW];
    const int n = device_list(d);
    for (int i = 0; i < n && (uint32_t)i < maxDevices && devices; i++) devices[i] = d[i];
    return (uint32_t)n;
}

arm_status arm_cuda_set_staging(uint32_t chunkMiB, uint32_t nStreams)
{
    if (chunkMiB > 4096 || nStreams > MAXS) return ARM_MATH_ARGUMENT_ERROR;
    pthread_mutex_lock(&g_cfg.mu);
    cfg_init_locked();
    if (chunkMiB) g_cfg.chunk = (size_t)chunkMiB << 20;
    if (nStreams) g_cfg.nstream = (int)nStreams;
    pthread_mutex_unlock(&g_cfg.mu);
    return ARM_MATH_SUCCESS;
}

arm_status arm_cuda_set_staging_ramp(uint32_t firstChunkMiB)
{
    if (firstChunkMiB > 4096) return ARM_MATH_ARGUMENT_ERROR;
    pthread_mutex_lock(&g_cfg.mu);
    cfg_init_locked();
    g_cfg.ramp = (size_t)firstChunkMiB << 20;
    pthread_mutex_unlock(&g_cfg.mu);
    return ARM_MATH_SUCCESS;
}

static void staging_config(size_t *chunk, int *nstream, size_t *ramp)
{
    pthread_mutex_lock(&g_cfg.mu);
    cfg_init_locked();
    *chunk = g_cfg.chunk;
    *nstream = g_cfg.nstream;
    *ramp = g_cfg.ramp;
    pthread_mutex_unlock(&g_cfg.mu);
}

/* ------------------------------------------------------------------ per-thread streams and staging buffers */

static void ctx_release(dev_ctx *c)
{
    for (int s = 0; s < MAXS; s++) {
        for (int b = 0; b < NBUF; b++)
            if (c->buf[s][b]) cmsisdsp_cuda_free(c->buf[s][b]);
        if (c->stream[s]) cmsisdsp_cuda_stream_destroy(c->stream[s]);
    }
    memset(c, 0, sizeof *c);
    c->device = -1;
}
static void pool_free(void *p)
{
    ctx_pool *pool = (ctx_pool *)p;
    if (!pool) return;
    for (int w = 0; w < MAXW; w++)
        if (pool->w[w].device >= 0) ctx_release(&pool->w[w]);
    free(pool);
}
static void make_key(void) { pthread_key_create(&g_key, pool_free); }
static ctx_pool *pool_get(void)
{
    pthread_once(&g_once, make_key);
    ctx_pool *pool = (ctx_pool *)pthread_getspecific(g_key);
    if (!pool) {
        pool = (ctx_pool *)calloc(1, sizeof *pool);
        if (!pool) return 0;
        for (int w = 0; w < MAXW; w++) pool->w[w].device = -1;
        pthread_setspecific(g_key, pool);
    }
    return pool;
}
void arm_cuda_release(void)
{
    pthread_once(&g_once, make_key);
    ctx_pool *pool = (ctx_pool *)pthread_getspecific(g_key);
    if (pool) {
        pthread_setspecific(g_key, 0);
        pool_free(pool);
    }
}

/* `device` is the current device: streams are created on it */
static int ctx_ensure(dev_ctx *c, int device, int nstream)
{
    if (c->device != device) {
        if (c->device >= 0) ctx_release(c);
        c->device = device;
    }
    while (c->nstream < nstream) {
        int rc = cmsisdsp_cuda_stream_create(&c->stream[c->nstream]);
        if (rc) return rc;
        c->nstream++;
    }
    return 0;
}
static int staging(dev_ctx *c, int s, int which, size_t bytes, void **out)
{
    if (c->cap[s][which] < bytes) {
        if (c->buf[s][which]) cmsisdsp_cuda_free(c->buf[s][which]);
        c->buf[s][which] = 0;
        c->cap[s][which] = 0;
        int rc = cmsisdsp_cuda_malloc(&c->buf[s][which], bytes);
        if (rc) return rc;
        c->cap[s][which] = bytes;
    }
    *out = c->buf[s][which];
    return 0;
}

/* ------------------------------------------------------------------ host buffers: one worker per device */

typedef struct {
    const arm_cuda_job *job;
    dev_ctx *ctx;
    int device;
    uint64_t f0, f1;                 /* this worker's frames */
    const char *in;
    char *out;
    size_t chunk, ramp;
    int nstream;
    int rc;                          /* shim code of the first failure */
} work_item;

static size_t span(uint64_t n, size_t stride, size_t frame) { return n ? (size_t)(n - 1) * stride + frame : 0; }

static void *worker(void *arg)
{
    work_item *it = (work_item *)arg;
    const arm_cuda_job *job = it->job;
    dev_ctx *c = it->ctx;
    int rc = cmsisdsp_cuda_set_device(it->device);
    if (!rc) rc = ctx_ensure(c, it->device, it->nstream);
    if (!rc) rc = job->prepare(job);
    if (rc) {
        it->rc = rc;
        return 0;
    }
    size_t per = job->inStride > job->outStride ? job->inStride : job->outStride;
    if (job->outB && job->outBStride > per) per = job->outBStride;
    uint64_t perChunk = it->chunk / (per ? per : 1);
    if (perChunk == 0) perChunk = 1;
    if (perChunk > it->f1 - it->f0) perChunk = it->f1 - it->f0;
    /* chunk sizes: rampMin, 2 rampMin, 4 rampMin ... perChunk ... and never more than half of what is left (but at
     * least rampMin), so the call ends on small chunks too */
    uint64_t rampMin = it->ramp ? it->ramp / (per ? per : 1) : perChunk;
    if (rampMin == 0) rampMin = 1;
    if (rampMin > perChunk) rampMin = perChunk;
    uint64_t up = rampMin, n = 0;
    int used = 0, s = 0;
    for (uint64_t f = it->f0; f < it->f1 && !rc; f += n, s = (s + 1) % it->nstream) {
        const uint64_t left = it->f1 - f;
        uint64_t half = (left + 1) / 2;
        if (half < rampMin) half = rampMin;
        n = up < perChunk ? up : perChunk;
        if (n > half) n = half;
        if (n > left) n = left;
        if (up < perChunk) up *= 2;
        void *st = c->stream[s], *din = 0, *dout = 0, *doutB = 0;
        if (s + 1 > used) used = s + 1;
        /* the stream serialises reuse of its staging buffers */
        if ((rc = staging(c, s, 0, span(perChunk, job->inStride, job->inFrame), &din))) break;
        if (job->inPlace) dout = din;
        else if ((rc = staging(c, s, 1, span(perChunk, job->outStride, job->outFrame), &dout))) break;
        if (job->outB && (rc = staging(c, s, 2, span(perChunk, job->outBStride, job->outBFrame), &doutB))) break;
        if ((rc = cmsisdsp_cuda_memcpy_h2d(din, it->in + f * job->inStride, span(n, job->inStride, job->inFrame), st))) break;
        if ((rc = job->launch(job, din, dout, doutB, n, st))) break;
        if ((rc = cmsisdsp_cuda_memcpy_d2h(it->out + f * job->outStride, dout, span(n, job->outStride, job->outFrame), st))) break;
        if (job->outB && (rc = cmsisdsp_cuda_memcpy_d2h(job->outB + f * job->outBStride, doutB, span(n, job->outBStride, job->outBFrame), st))) break;
        if (job->out2 && job->post) {
            if ((rc = job->post(job, din, n, st))) break;
            rc = cmsisdsp_cuda_memcpy_d2h(job->out2 + f * job->out2Stride, din, span(n, job->out2Stride, job->out2Frame), st);
        }
    }
    for (int i = 0; i < used; i++) {
        const int e = cmsisdsp_cuda_stream_synchronize(c->stream[i]);
        if (e && !rc) rc = e;
    }
    it->rc = rc;
    return 0;
}

arm_status arm_cuda_run_host(const arm_cuda_job *job, const void *in, void *out, uint64_t nFrames)
{
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    int devs[MAXW];
    const int nd = device_list(devs);
    if (nd <= 0) return ARM_MATH_CUDA_NO_DEVICE;
    ctx_pool *pool = pool_get();
    if (!pool) return ARM_MATH_CUDA_RUNTIME_ERROR;
    size_t chunk, ramp;
    int nstream;
    staging_config(&chunk, &nstream, &ramp);

    size_t per = job->inStride > job->outStride ? job->inStride : job->outStride;
    const uint64_t total = nFrames * (uint64_t)(per ? per : 1);
    uint64_t W = (total + MIN_BYTES_PER_WORKER - 1) / MIN_BYTES_PER_WORKER;
    if (W > (uint64_t)nd) W = (uint64_t)nd;
    if (W > nFrames) W = nFrames;
    if (W < 1) W = 1;
    const uint64_t block = (nFrames + W - 1) / W;              /* ceil(B / G) frames per device */

    work_item item[MAXW];
    pthread_t th[MAXW];
    int started[MAXW] = {0};
    int nw = 0;
    for (uint64_t w = 0; w < W; w++) {
        const uint64_t f0 = w * block, f1 = (f0 + block < nFrames) ? f0 + block : nFrames;
        if (f0 >= f1) break;
        work_item *it = &item[nw];
        it->job = job; it->ctx = &pool->w[nw]; it->device = devs[nw]; it->f0 = f0; it->f1 = f1;
        it->in = (const char *)in; it->out = (char *)out; it->chunk = chunk; it->ramp = ramp; it->nstream = nstream; it->rc = 0;
        nw++;
    }
    const int cur = cmsisdsp_cuda_get_device();
    for (int w = 1; w < nw; w++) {
        if (pthread_create(&th[w], 0, worker, &item[w]) == 0) started[w] = 1;
        else item[w].rc = CMSISDSP_CUDA_ERR_RUNTIME;
    }
    worker(&item[0]);                                          /* the calling thread is worker 0 */
    for (int w = 1; w < nw; w++)
        if (started[w]) pthread_join(th[w], 0);
    if (cur >= 0 && cur != item[0].device) cmsisdsp_cuda_set_device(cur);
    for (int w = 0; w < nw; w++)
        if (item[w].rc) return arm_cuda_status_of(item[w].rc);
    return ARM_MATH_SUCCESS;
}

/* ------------------------------------------------------------------ device buffers */

arm_status arm_cuda_run_device(const arm_cuda_job *job, const void *in, void *out, uint64_t nFrames)
{
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    const int dev = cmsisdsp_cuda_pointer_device(in);
    if (dev < 0) return ARM_MATH_ARGUMENT_ERROR;
    const int cur = cmsisdsp_cuda_get_device();
    if (cur < 0) return ARM_MATH_CUDA_NO_DEVICE;
    int rc = 0;
    if (cur != dev) rc = cmsisdsp_cuda_set_device(dev);        /* the call runs where the data is */
    /* The legacy default stream: the transform is ordered after whatever the caller enqueued on that stream or on
     * any blocking stream (a cudaMemcpy, a kernel that produced the frames), as a C caller expects of a function
     * that takes device buffers; the library's own streams are non-blocking and would race with such work. */
    void *const st = 0;
    if (!rc) rc = job->prepare(job);
    if (!rc) rc = job->launch(job, in, out, job->outB, nFrames, st);
    if (!rc && job->out2 && job->post) rc = job->post(job, (void *)in, nFrames, st);
    if (!rc) rc = cmsisdsp_cuda_stream_synchronize(st);
    if (cur != dev) cmsisdsp_cuda_set_device(cur);
    return arm_cuda_status_of(rc);
}

arm_status arm_cuda_run(const arm_cuda_job *job, const void *in, void *out, uint64_t nFrames)
{
    if (nFrames == 0) return ARM_MATH_SUCCESS;
    if (cmsisdsp_cuda_device_count() <= 0) {
        (void)cmsisdsp_cuda_get_device();                      /* records why in cmsisdsp_cuda_last_error() */
        return ARM_MATH_CUDA_NO_DEVICE;
    }
    const int di = cmsisdsp_cuda_pointer_device(in), dq = cmsisdsp_cuda_pointer_device(out);
    const int db = job->outB ? cmsisdsp_cuda_pointer_device(job->outB) : dq;
    if (di < -1 || dq < -1 || db < -1) return ARM_MATH_CUDA_RUNTIME_ERROR;
    if (di != dq || db != dq) return ARM_MATH_ARGUMENT_ERROR;  /* all host, or all on one device */
    return di >= 0 ? arm_cuda_run_device(job, in, out, nFrames) : arm_cuda_run_host(job, in, out, nFrames);
}
