// Stand-alone probe: which TMA tile-load forms work on this box.  Not part of the library.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int nfloats, int c0, int c1, int mode) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((nfloats * 4 + 127) / 128) * 128);
    const bool producer = mode == 0 ? (threadIdx.x == 0) : (threadIdx.x < 32);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (producer) {
        bool go = true;
        if (mode == 1) {
            uint32_t pred;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
            go = pred != 0;
        }
        if (go) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nfloats * 4) : "memory");
            if (RANK == 4)
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(tile)),
                             "l"(&tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(0), "r"(0) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)),
                             "l"(&tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
        }
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int t = threadIdx.x; t < nfloats; t += blockDim.x) out[t] = tile[t];
}

int main() {
    PFN_cuTensorMapEncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    const int Ny = 64, Nx = 64, Nz = 8, NV = 8;
    std::vector<float> h((size_t)NV * Nz * Nx * Ny);
    for (size_t t = 0; t < h.size(); ++t) h[t] = (float)t;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 1 << 20);
    struct V { int rank, bw, bh, mode, c0, c1; } vs[] = {
        {2, 40, 18, 0, 0, -1}, {2, 40, 18, 0, -4, 0}, {2, 40, 18, 0, -4, -1}, {4, 40, 18, 0, -4, -1}, {4, 40, 18, 0, 28, 27},
        {4, 40, 18, 0, 56, 55}, {4, 40, 18, 1, 56, 55}, {4, 40, 18, 0, 2, 3}};
    for (auto v : vs) {
        CUtensorMap m;
        cuuint64_t dims[4] = {Ny, Nx, Nz, NV};
        cuuint64_t str[3] = {Ny * 4, (cuuint64_t)Nx * Ny * 4, (cuuint64_t)Nz * Nx * Ny * 4};
        cuuint32_t box[4] = {(cuuint32_t)v.bw, (cuuint32_t)v.bh, 1, (cuuint32_t)(v.rank == 4 ? NV : 1)};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, v.rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int nf = v.bw * v.bh * (v.rank == 4 ? NV : 1);
        size_t smem = ((nf * 4 + 127) / 128) * 128 + 64;
        if (v.rank == 4) { cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<4><<<1, 128, smem>>>(m, o, nf, v.c0, v.c1, v.mode); }
        else             { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<1, 128, smem>>>(m, o, nf, v.c0, v.c1, v.mode); }
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> r0(nf);
        if (e == cudaSuccess) cudaMemcpy(r0.data(), o, nf * 4, cudaMemcpyDeviceToHost);
        // expected element (row 1, col 1) of variable 0 = value at (i=c1+1, j=c0+1)
        float expect = (v.c1 + 1 < 0 || v.c0 + 1 < 0) ? 0.f : (float)((v.c1 + 1) * Ny + (v.c0 + 1));
        printf("rank %d box %dx%d mode %d c=(%d,%d): encode=%d run=%s got[1][1]=%g expect=%g\n", v.rank, v.bw, v.bh, v.mode, v.c0, v.c1, (int)r,
               cudaGetErrorString(e), e == cudaSuccess ? r0[v.bw + 1] : -1.f, expect);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); break; }
    }
    return 0;
}
