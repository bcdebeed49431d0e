// CFL / stability scan on the device.
//
// Replaces src/on-device/utils/compute_stability.cpp: a host raster scan that builds three 8x8 flux Jacobians per
// cell and calls Eigen for their eigenvalues (:108-137, :165-181), forked once from the driver when eigen_bin_name
// != "none" (no_diffusion.cu:259-276).  Here: one HBM-bound pass over the state (32 B per cell), the spectral radius
// of the ideal-MHD flux Jacobian in closed form per direction,
//     {0, u_d, u_d +- c_a, u_d +- c_s, u_d +- c_f},  c_a^2 = B_d^2/rho,
//     c_f^2, c_s^2 = roots of x^2 - (a^2 + b^2) x + a^2 c_a^2,  a^2 = gamma p/rho,  b^2 = B^2/rho,
// (a negative root is an imaginary speed: modulus sqrt(u_d^2 - x), Eigen's complex abs in the reference, :173),
// LHS = (dt/dx)|l_x| + (dt/dy)|l_y| + (dt/dz)|l_z| (:157-163), a block reduction and one 64-bit atomicMax + one
// atomicAdd per block.  The reference's x matrix has exactly this spectrum; its y and z matrices carry sign / index
// slips (oracle/stability.py, quirk B-26) that are deliberately NOT reproduced.
#include <string.h>

#include "imhd_common.cuh"

namespace imhd {

// sqrt.approx / rcp.approx: 1 MUFU each, relative error ~1e-7 -- two orders below what a CFL bound needs, and the
// IEEE sequences would make this one-pass scan instruction-bound instead of HBM-bound
__device__ __forceinline__ float fsqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// max over x in {c_f^2, c_s^2, c_a^2} of (x >= 0 ? |u| + sqrt(x) : sqrt(u^2 - x)); both branches are monotone in x,
// so only the largest and the smallest of the three matter
__device__ __forceinline__ float radius(float ud, float cd2, float a2, float b2) {
    const float s = a2 + b2;
    const float disc = fsqrt(fmaxf(fmaf(s, s, -4.0f * a2 * cd2), 0.0f));
    const float xf = 0.5f * (s + disc), xs = 0.5f * (s - disc);
    const float hi = fmaxf(fmaxf(xf, xs), cd2), lo = fminf(fminf(xf, xs), cd2);
    const float au = fabsf(ud);
    float best = au;
    if (hi > 0.0f) best = au + fsqrt(hi);
    if (lo < 0.0f) best = fmaxf(best, fsqrt(fmaf(ud, ud, -lo)));  // imaginary speed (non-physical state)
    return (s != s || cd2 != cd2 || ud != ud) ? __int_as_float(0x7fc00000) : best;  // fmaxf/fminf drop NaNs: keep a NaN cell NaN
}

__device__ __forceinline__ float cell_lhs(const float U[8], float tx, float ty, float tz) {
    const float inv = fast_rcp(U[RHO]);
    const float u = U[MX] * inv, v = U[MY] * inv, w = U[MZ] * inv;
    const float usq = fmaf(w, w, fmaf(v, v, u * u));
    const float Bsq = fmaf(U[BZ], U[BZ], fmaf(U[BY], U[BY], U[BX] * U[BX]));
    // the scanner's pressure uses rho u^2 / 2 (compute_stability.cpp:206), unlike the solver's helper (B-1)
    const float p = kGm1f * (U[EN] - 0.5f * U[RHO] * usq - 0.5f * Bsq);
    const float a2 = (float)kGamma * p * inv, b2 = Bsq * inv;
    const float lx = radius(u, U[BX] * U[BX] * inv, a2, b2);
    const float ly = radius(v, U[BY] * U[BY] * inv, a2, b2);
    const float lz = radius(w, U[BZ] * U[BZ] * inv, a2, b2);
    return fmaf(tz, lz, fmaf(ty, ly, tx * lx));
}

// out[0]: (float bits of max LHS) << 32 | (0xFFFFFFFF - cell)  -> atomicMax picks the largest LHS and, among equal
//         ones, the first cell in the reference's scan order (k, i, j) = the smallest linear index
// out[1]: number of cells with LHS >= 1
// VEC = 4: four consecutive cells per thread and iteration through 16-byte loads (needs 16-byte aligned variable
// arrays and ncells % 4 == 0); VEC = 1 otherwise.
template <int VEC>
__global__ void __launch_bounds__(256) k_stability(const float* __restrict__ Q, long long vs, long long first,
                                                   long long ncells, float tx, float ty, float tz,
                                                   unsigned long long* __restrict__ out) {
    unsigned long long key = 0, viol = 0;
    const long long ngroups = ncells / VEC;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (long long)gridDim.x * blockDim.x) {
        float U[VEC][8];
        if (VEC == 4) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(Q + first + v * vs) + g);
                U[0][v] = q.x; U[1 % VEC][v] = q.y; U[2 % VEC][v] = q.z; U[3 % VEC][v] = q.w;
            }
        } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) U[0][v] = __ldg(Q + first + g + v * vs);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float lhs = cell_lhs(U[e], tx, ty, tz);
            if (lhs >= 1.0f) ++viol;
            if (lhs > 0.0f) {  // false for NaN (rho == 0), as `>=` / `>` are in the reference
                const unsigned long long k2 =
                    ((unsigned long long)__float_as_uint(lhs) << 32) | (0xFFFFFFFFu - (unsigned)(g * VEC + e));
                key = k2 > key ? k2 : key;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
        key = other > key ? other : key;
        viol += __shfl_xor_sync(0xFFFFFFFFu, viol, o);
    }
    __shared__ unsigned long long skey[8], sviol[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { skey[warp] = key; sviol[warp] = viol; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) {
            key = skey[q] > key ? skey[q] : key;
            viol += sviol[q];
        }
        if (key) atomicMax(out, key);
        if (viol) atomicAdd(out + 1, viol);
    }
}

}  // namespace imhd

using namespace imhd;

#include <mutex>

namespace {
std::mutex g_scan_mu;                        // the scan is synchronous; one at a time per process keeps the scratch simple
unsigned long long* g_scratch[64] = {nullptr};  // 16 bytes per device, allocated on first use, never freed
}  // namespace

extern "C" int imhd_stability_scan(const float* Q, const imhd_slab* s, imhd_stability* host_out, void* stream) {
    if (!Q || !s || !host_out) { set_error("imhd_stability_scan: null argument"); return IMHD_E_INVALID; }
    if (int e = bad_dims(s->Nx, s->Ny, s->Nz)) return e;
    if (s->nzl < 1 || s->k0 < 0 || s->k0 + s->nzl > s->Nz || s->ghosts < 0) {
        set_error("imhd_stability_scan: slab [%d,%d) outside 0..%d", s->k0, s->k0 + s->nzl, s->Nz);
        return IMHD_E_INVALID;
    }
    const long long plane = (long long)s->Nx * s->Ny, ncells = plane * s->nzl;
    if (ncells > 0xFFFFFFFFll) { set_error("imhd_stability_scan: more than 2^32 cells in one slab"); return IMHD_E_INVALID; }
    const long long vs = plane * (s->nzl + 2 * s->ghosts), first = plane * s->ghosts;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    IMHD_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 0 || dev >= 64) { set_error("imhd_stability_scan: device ordinal %d not supported", dev); return IMHD_E_INVALID; }
    std::lock_guard<std::mutex> guard(g_scan_mu);
    if (!g_scratch[dev]) IMHD_CUDA(cudaMalloc(&g_scratch[dev], 2 * sizeof(unsigned long long)));
    unsigned long long* d_out = g_scratch[dev];
    IMHD_CUDA(cudaMemsetAsync(d_out, 0, 2 * sizeof(unsigned long long), st));
    const bool vec4 = ncells % 4 == 0 && vs % 4 == 0 && first % 4 == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0;
    const long long want = (ncells / (vec4 ? 4 : 1) + 255) / 256;
    const unsigned grid = (unsigned)(want < (long long)sms * 8 ? want : (long long)sms * 8);  // 8 resident blocks of 256 per SM
    const float tx = s->dt / s->dx, ty = s->dt / s->dy, tz = s->dt / s->dz;
    if (vec4) k_stability<4><<<grid, 256, 0, st>>>(Q, vs, first, ncells, tx, ty, tz, d_out);
    else      k_stability<1><<<grid, 256, 0, st>>>(Q, vs, first, ncells, tx, ty, tz, d_out);
    IMHD_LAUNCH_CHECK(1);
    unsigned long long h[2] = {0, 0};
    IMHD_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st));
    IMHD_CUDA(cudaStreamSynchronize(st));
    const unsigned bits = (unsigned)(h[0] >> 32);
    const long long cell = h[0] ? (long long)(0xFFFFFFFFu - (unsigned)(h[0] & 0xFFFFFFFFu)) : 0;
    float mx;
    memcpy(&mx, &bits, sizeof(mx));
    host_out->max_lhs = h[0] ? mx : 0.0f;
    host_out->k = s->k0 + (int)(cell / plane);
    host_out->i = (int)((cell % plane) / s->Ny);
    host_out->j = (int)(cell % s->Ny);
    host_out->violations = h[1];
    host_out->dt_new = h[0] ? 0.1f * s->dt / mx : 0.0f;  // alpha = 0.1 (compute_stability.cpp:139-141)
    return 0;
}
