// Probe: issue rate of packed fp32x2 instructions whose multiplier is a broadcast scalar (uniform register / immediate),
// alone and interleaved with scalar 3-register FFMA.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o packed_probe packed_const_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
struct Cf { float c0, c1, c2, c3; };
template <int MODE> __global__ void k(float* out, int iters, Cf c, const float* in) {
    float2 p[8], q[8]; float a[8], b[8], d[8];
    for (int i = 0; i < 8; ++i) {
        p[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 8]); q[i] = make_float2(in[threadIdx.x + i + 16], in[threadIdx.x + i + 24]);
        a[i] = in[threadIdx.x + 32 + i]; b[i] = in[threadIdx.x + 40 + i]; d[i] = in[threadIdx.x + 48 + i];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 3) p[i] = __ffma2_rn(p[i], make_float2(c.c0, c.c0), q[i]);       // FFMA2 pair, UR scalar, pair
            if (MODE == 1) p[i] = __fadd2_rn(p[i], q[i]);                                             // FADD2
            if (MODE == 2) p[i] = __ffma2_rn(p[i], make_float2(0.5f, 0.5f), q[i]);                    // FFMA2 with immediate
            if (MODE == 3 || MODE == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(d[i]));  // scalar 3-reg FFMA
            if (MODE == 5) { p[i] = __fadd2_rn(p[i], q[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(d[i])); }
            if (MODE == 6) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c.c1)); }            // scalar FFMA with UR
            if (MODE == 8 || MODE == 9) p[i] = __ffma2_rn(p[i], make_float2(c.c0, c.c0), q[i]);
            if (MODE == 8 || MODE == 9 || MODE == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(int*)&b[i]) : "r"(it), "r"(i));
            if (MODE == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(int*)&d[i]) : "r"(it), "r"(i + 1));
            if (MODE == 10) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[i]), "f"(c.c1));
            if (MODE == 11) { p[i] = __ffma2_rn(p[i], make_float2(c.c0, c.c0), q[i]); asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+f"(b[i])); }
            if (MODE == 7) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c.c1));
                             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d[i]) : "f"(b[i]), "f"(c.c2)); }            // two scalar = one packed
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y + d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *o, *in; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
    Cf c{1.0001f, 0.999f, 1.0002f, 0.5f};
    const int iters = 20000;
    const char* names[] = {"FFMA2 ur", "FADD2", "FFMA2 imm", "FFMA2ur+FFMA", "FFMA 3reg", "FADD2+FFMA", "FFMA ur", "2xFFMA ur", "FFMA2ur+LOP3", "FFMA2ur+2LOP3", "FFMAur+LOP3", "FFMA2ur+SHFL"};
    const int per[] = {1, 1, 1, 2, 1, 2, 1, 2, 2, 3, 2, 2};
    for (int threads : {256}) for (int mode = 0; mode < 12; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&] { switch (mode) { case 0: k<0><<<148, threads>>>(o, iters, c, in); break; case 1: k<1><<<148, threads>>>(o, iters, c, in); break;
            case 2: k<2><<<148, threads>>>(o, iters, c, in); break; case 3: k<3><<<148, threads>>>(o, iters, c, in); break; case 4: k<4><<<148, threads>>>(o, iters, c, in); break;
            case 5: k<5><<<148, threads>>>(o, iters, c, in); break; case 6: k<6><<<148, threads>>>(o, iters, c, in); break; case 7: k<7><<<148, threads>>>(o, iters, c, in); break;
            case 8: k<8><<<148, threads>>>(o, iters, c, in); break; case 9: k<9><<<148, threads>>>(o, iters, c, in); break; case 10: k<10><<<148, threads>>>(o, iters, c, in); break;
            default: k<11><<<148, threads>>>(o, iters, c, in); } };
        run(); cudaDeviceSynchronize(); cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double w = (double)iters * 8 * threads / 32 * per[mode];
        printf("thr/SM %4d %-14s %.3f ms  %.2f warp-instr/clk/SM (1.965 GHz)\n", threads, names[mode], ms, w / (ms * 1e6) / 1.965);
    }
    return 0;
}
