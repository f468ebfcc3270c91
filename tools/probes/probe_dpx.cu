// Semantics / throughput probes used while designing the q15 butterflies (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_dpx probe_dpx.cu && ./probe_dpx
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void sem(unsigned *o)
{
    // does the packed add inside VIADDMNMX.S16x2 keep the 17th bit?
    o[0] = __viaddmin_s16x2(0x7fff7fffu, 0x00010001u, 0x7fff7fffu);   // 32767+1 min 32767: 0x7fff if wide, 0x8000 if it wraps
    o[1] = __viaddmax_s16x2(0x80008000u, 0xffffffffu, 0x80008000u);   // -32768-1 max -32768: 0x8000 if wide, 0x7fff if it wraps
    o[2] = __viaddmin_s16x2(0x40004000u, 0x40004000u, 0x7fff7fffu);
    o[3] = __vimax3_s16x2(0x00010002u, 0xfffffffeu, 0x00030000u);
}

__device__ __forceinline__ int hiw(int a, int b)
{
    int hi;
    asm("{\n\t.reg .b64 t;\n\t.reg .b32 lo;\n\tmul.wide.s32 t, %1, %2;\n\tmov.b64 {lo, %0}, t;\n\t}" : "=r"(hi) : "r"(a), "r"(b));
    return hi;
}
template <int MODE> __global__ void tput(int *o, int k, int iters)
{
    int a = threadIdx.x, b = threadIdx.x * 3 + 1, c = threadIdx.x ^ 5, d = k;
    int e = a + 7, f = b + 9, g = c + 11, h = d + 13;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) {          // SHF
                a = (a >> 1) ^ b; b = (b >> 1) ^ c; c = (c >> 1) ^ d; d = (d >> 1) ^ a;
                e = (e >> 1) ^ f; f = (f >> 1) ^ g; g = (g >> 1) ^ h; h = (h >> 1) ^ e;
            } else if (MODE == 1) {   // IMAD.WIDE upper word
                a = hiw(a, k) ^ b; b = hiw(b, k) ^ c; c = hiw(c, k) ^ d; d = hiw(d, k) ^ a;
                e = hiw(e, k) ^ f; f = hiw(f, k) ^ g; g = hiw(g, k) ^ h; h = hiw(h, k) ^ e;
            } else if (MODE == 2) {   // IMAD 32-bit
                a = a * k + b; b = b * k + c; c = c * k + d; d = d * k + a;
                e = e * k + f; f = f * k + g; g = g * k + h; h = h * k + e;
            } else {                  // VIADDMNMX
                a = __viaddmin_s32(a, b, k); b = __viaddmin_s32(b, c, k); c = __viaddmin_s32(c, d, k); d = __viaddmin_s32(d, a, k);
                e = __viaddmin_s32(e, f, k); f = __viaddmin_s32(f, g, k); g = __viaddmin_s32(g, h, k); h = __viaddmin_s32(h, e, k);
            }
        }
    }
    o[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + e + f + g + h;
}

template <int MODE> void run(const char *name, int *d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, blocks = 148 * 8, threads = 256;
    tput<MODE><<<blocks, threads>>>(d, 1 << 30, 10);
    cudaEventRecord(e0);
    tput<MODE><<<blocks, threads>>>(d, 1 << 30, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * 16 * 8;        // "main" ops (each iteration also has 8 side ops in modes 0, 1)
    printf("%-12s %.3f ms  %.1f Gop/s per SM per clock-ish: %.2f thread-ops/clk/SM (at 1.965 GHz)\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    unsigned *o, h[4];
    cudaMalloc(&o, 64);
    sem<<<1, 1>>>(o);
    cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost);
    printf("viaddmin_s16x2(0x7fff,1,0x7fff)=%08x  viaddmax_s16x2(0x8000,-1,0x8000)=%08x  viaddmin(0x4000+0x4000,0x7fff)=%08x vimax3=%08x\n", h[0], h[1], h[2], h[3]);
    int *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("shf+lop", d);
    run<1>("imad.wide+add", d);
    run<2>("imad", d);
    run<3>("viaddmnmx", d);
    return 0;
}
