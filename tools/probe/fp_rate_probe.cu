// Probe: issue rate of 3-register fp32 instructions (FFMA / FMUL / FADD) vs packed FFMA2 / FMUL2 / FADD2 and vs the
// immediate form, per SM.  Not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE> __global__ void k(float* out, int iters, const float* in) {
    float a[8], b[8], c[8]; float2 p[8], q[8], r[8];
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i];
        p[i] = make_float2(a[i], b[i]); q[i] = make_float2(b[i], c[i]); r[i] = make_float2(c[i], a[i]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));        // FFMA 3-reg
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(0.25f));      // FFMA imm
            if (MODE == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));                       // FMUL
            if (MODE == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));                       // FADD
            if (MODE == 4) p[i] = __ffma2_rn(p[i], q[i], r[i]);                                                   // FFMA2 3-reg
            if (MODE == 5) p[i] = __fmul2_rn(p[i], q[i]);                                                         // FMUL2
            if (MODE == 6) p[i] = __fadd2_rn(p[i], q[i]);                                                         // FADD2
            if (MODE == 7) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));      // FFMA + IADD mix
                             asm volatile("add.s32 %0, %0, %1;" : "+r"(*(int*)&c[(i + 1) & 7]) : "r"(i)); }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *o, *in; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
    const int iters = 20000; const char* names[] = {"FFMA 3-reg", "FFMA imm", "FMUL", "FADD", "FFMA2 3-reg", "FMUL2", "FADD2", "FFMA+IADD"};
    for (int threads : {256, 512}) for (int mode = 0; mode < 8; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&] { switch (mode) { case 0: k<0><<<148, threads>>>(o, iters, in); break; case 1: k<1><<<148, threads>>>(o, iters, in); break;
            case 2: k<2><<<148, threads>>>(o, iters, in); break; case 3: k<3><<<148, threads>>>(o, iters, in); break; case 4: k<4><<<148, threads>>>(o, iters, in); break;
            case 5: k<5><<<148, threads>>>(o, iters, in); break; case 6: k<6><<<148, threads>>>(o, iters, in); break; default: k<7><<<148, threads>>>(o, iters, in); } };
        run(); cudaDeviceSynchronize(); cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double w = (double)iters * 8 * threads / 32 * (mode == 7 ? 2 : 1);
        printf("thr/SM %4d %-12s %.3f ms  %.2f warp-instr/clk/SM (1.965 GHz)\n", threads, names[mode], ms, w / (ms * 1e6) / 1.965);
    }
    return 0;
}
