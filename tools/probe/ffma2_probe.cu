// Probe: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) on this GPU.  Not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long f2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 0.5f); }
    unsigned long long ps = pk(s, s), pc = pk(0.25f, 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(0.25f)); }
            else if (MODE == 1) p[i] = f2(p[i], ps, pc);
            else { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(0.25f)); p[i] = f2(p[i], ps, pc); }
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i] + (float)(p[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* o; cudaMalloc(&o, 148 * 1024 * 8 * 4);
    const int iters = 20000;
    for (int threads : {256, 512, 1024}) for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&] { if (mode == 0) k<0><<<148, threads>>>(o, iters, 1.0001f); else if (mode == 1) k<1><<<148, threads>>>(o, iters, 1.0001f); else k<2><<<148, threads>>>(o, iters, 1.0001f); };
        run(); cudaDeviceSynchronize();
        cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double winstr = (double)iters * 8 * (mode == 2 ? 2 : 1) * threads / 32;   // warp-instructions per SM
        printf("threads/SM %4d mode %s: %.3f ms, %.2f warp-instr/ns/SM (at 1.965 GHz: %.2f per clk), fma/clk/SM %.0f\n", threads,
               mode == 0 ? "FFMA " : mode == 1 ? "FFMA2" : "mixed", ms, winstr / (ms * 1e6), winstr / (ms * 1e6) / 1.965,
               winstr / (ms * 1e6) / 1.965 * 32 * (mode == 0 ? 1 : mode == 1 ? 2 : 1.5));
    }
    return 0;
}
