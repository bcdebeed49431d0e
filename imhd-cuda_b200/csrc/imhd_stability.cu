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
// slips (oracle/stability.py, quirk B-26).  Two modes:
//   IMHD_STABILITY_WAVE_SPEEDS      (default) the exact MHD wave speeds in all three directions -- the bound the
//                                   reference means to compute;
//   IMHD_STABILITY_REFERENCE_QUIRKS the reference's report: closed form for its (correct) x matrix, and for y and z the
//                                   spectral radius of ITS matrices (computeB / computeC, compute_stability.cpp:296-459,
//                                   slips included), eigenvalues by elimination to Hessenberg form + shifted QR per cell
//                                   in fp64 (the reference calls Eigen's fp32 solver, :165-181).
#include <string.h>

#include "imhd_common.cuh"
#include "imhd_engine.h"

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

// ---- eigenvalues of a general real 8x8 matrix: reduction to Hessenberg form by stabilised elimination, then the
// shifted QR algorithm (the classical EISPACK elmhes / hqr pair), returning the largest modulus.  1-based arrays.
__device__ double spectral_radius8(double (&a)[9][9]) {
    const int n = 8;
    for (int m = 2; m < n; ++m) {  // elmhes
        double x = 0.0;
        int i = m;
        for (int j = m; j <= n; ++j)
            if (fabs(a[j][m - 1]) > fabs(x)) { x = a[j][m - 1]; i = j; }
        if (i != m) {
            for (int j = m - 1; j <= n; ++j) { const double t = a[i][j]; a[i][j] = a[m][j]; a[m][j] = t; }
            for (int j = 1; j <= n; ++j) { const double t = a[j][i]; a[j][i] = a[j][m]; a[j][m] = t; }
        }
        if (x != 0.0)
            for (i = m + 1; i <= n; ++i) {
                double y = a[i][m - 1];
                if (y != 0.0) {
                    y /= x;
                    a[i][m - 1] = y;
                    for (int j = m; j <= n; ++j) a[i][j] -= y * a[m][j];
                    for (int j = 1; j <= n; ++j) a[j][m] += y * a[j][i];
                }
            }
    }
    for (int i = 3; i <= n; ++i)
        for (int j = 1; j <= i - 2; ++j) a[i][j] = 0.0;  // the multipliers elmhes leaves below the subdiagonal
    double best = 0.0, anorm = 0.0, t = 0.0, p = 0.0, q = 0.0, r = 0.0, x, y, z, w, sq;
    for (int i = 1; i <= n; ++i)
        for (int j = (i - 1 > 1 ? i - 1 : 1); j <= n; ++j) anorm += fabs(a[i][j]);
    if (!(anorm == anorm) || anorm > 1e300) return __longlong_as_double(0x7ff8000000000000ll);  // NaN / inf cell (rho == 0)
    int nn = n;
    while (nn >= 1) {  // hqr
        int its = 0, l;
        do {
            for (l = nn; l >= 2; --l) {
                sq = fabs(a[l - 1][l - 1]) + fabs(a[l][l]);
                if (sq == 0.0) sq = anorm;
                if (fabs(a[l][l - 1]) + sq == sq) { a[l][l - 1] = 0.0; break; }
            }
            x = a[nn][nn];
            if (l == nn) {  // one real root
                best = fmax(best, fabs(x + t));
                --nn;
            } else {
                y = a[nn - 1][nn - 1];
                w = a[nn][nn - 1] * a[nn - 1][nn];
                if (l == nn - 1) {  // two roots
                    p = 0.5 * (y - x);
                    q = p * p + w;
                    z = sqrt(fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + (p >= 0.0 ? z : -z);
                        best = fmax(best, fabs(x + z));
                        if (z != 0.0) best = fmax(best, fabs(x - w / z));
                    } else {
                        best = fmax(best, sqrt((x + p) * (x + p) + z * z));  // complex pair: Eigen's complex abs
                    }
                    nn -= 2;
                } else {
                    if (its == 60) return __longlong_as_double(0x7ff8000000000000ll);
                    if (its == 10 || its == 20) {  // exceptional shift
                        t += x;
                        for (int i = 1; i <= nn; ++i) a[i][i] -= x;
                        sq = fabs(a[nn][nn - 1]) + fabs(a[nn - 1][nn - 2]);
                        y = x = 0.75 * sq;
                        w = -0.4375 * sq * sq;
                    }
                    ++its;
                    int m;
                    for (m = nn - 2; m >= l; --m) {
                        z = a[m][m];
                        r = x - z;
                        sq = y - z;
                        p = (r * sq - w) / a[m + 1][m] + a[m][m + 1];
                        q = a[m + 1][m + 1] - z - r - sq;
                        r = a[m + 2][m + 1];
                        sq = fabs(p) + fabs(q) + fabs(r);
                        p /= sq; q /= sq; r /= sq;
                        if (m == l) break;
                        const double u = fabs(a[m][m - 1]) * (fabs(q) + fabs(r));
                        const double v = fabs(p) * (fabs(a[m - 1][m - 1]) + fabs(z) + fabs(a[m + 1][m + 1]));
                        if (u + v == v) break;
                    }
                    for (int i = m + 2; i <= nn; ++i) {
                        a[i][i - 2] = 0.0;
                        if (i != m + 2) a[i][i - 3] = 0.0;
                    }
                    for (int k = m; k <= nn - 1; ++k) {  // double QR step on rows l..nn, columns m..nn
                        if (k != m) {
                            p = a[k][k - 1];
                            q = a[k + 1][k - 1];
                            r = k != nn - 1 ? a[k + 2][k - 1] : 0.0;
                            if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        sq = sqrt(p * p + q * q + r * r);
                        sq = p >= 0.0 ? sq : -sq;
                        if (sq != 0.0) {
                            if (k == m) {
                                if (l != m) a[k][k - 1] = -a[k][k - 1];
                            } else {
                                a[k][k - 1] = -sq * x;
                            }
                            p += sq;
                            x = p / sq; y = q / sq; z = r / sq;
                            q /= p; r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = a[k][j] + q * a[k + 1][j];
                                if (k != nn - 1) { p += r * a[k + 2][j]; a[k + 2][j] -= p * z; }
                                a[k + 1][j] -= p * y;
                                a[k][j] -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * a[i][k] + y * a[i][k + 1];
                                if (k != nn - 1) { p += z * a[i][k + 2]; a[i][k + 2] -= p * r; }
                                a[i][k + 1] -= p * q;
                                a[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return best;
}

// the reference's y and z flux Jacobians B, C (computeB / computeC, compute_stability.cpp:296-373, :395-459), slips of
// quirk B-26 included, built in fp32 as there; DIR = 1 (B) or 2 (C)
template <int DIR>
__device__ void reference_jacobian(const float U[8], double (&M)[9][9]) {
    const float g = (float)kGamma;
    const float rho = U[RHO], u = U[MX] / rho, v = U[MY] / rho, w = U[MZ] / rho;
    const float Bx = U[BX], By = U[BY], Bz = U[BZ], e = U[EN];
    const float usq = u * u + v * v + w * w, Bsq = Bx * Bx + By * By + Bz * Bz, Bdotu = Bx * u + By * v + Bz * w;
    const float H = (g * e + (2 - g) * 0.5f * Bsq) / rho;  // the bracket shared by the energy rows
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) M[i][j] = 0.0;
#define IMHD_M(r, c) M[(r) + 1][(c) + 1]
    if (DIR == 1) {
        IMHD_M(0, 2) = 1.0f;
        IMHD_M(1, 0) = -u * v; IMHD_M(1, 1) = v; IMHD_M(1, 2) = u; IMHD_M(1, 4) = -By; IMHD_M(1, 5) = -Bx;
        IMHD_M(2, 0) = 0.5f * (g - 1) * usq - v * v; IMHD_M(2, 1) = (1 - g) * u; IMHD_M(2, 2) = (3 - g) * v; IMHD_M(2, 3) = (1 - g) * w;
        IMHD_M(2, 4) = (2 - g) * Bx; IMHD_M(2, 5) = -g * By; IMHD_M(2, 6) = (2 - g) * Bz; IMHD_M(2, 7) = g - 1;
        IMHD_M(3, 0) = -v * w; IMHD_M(3, 2) = w; IMHD_M(3, 3) = v; IMHD_M(3, 5) = -Bz; IMHD_M(3, 6) = -By;
        IMHD_M(4, 0) = (v * Bx - u * By) / rho; IMHD_M(4, 1) = By / rho; IMHD_M(4, 2) = -Bx / rho; IMHD_M(4, 4) = -v; IMHD_M(4, 5) = u;  // slip
        IMHD_M(6, 0) = (v * Bz - w * By) / rho; IMHD_M(6, 2) = -Bz / rho; IMHD_M(6, 3) = By / rho; IMHD_M(6, 5) = w; IMHD_M(6, 6) = -v;
        IMHD_M(7, 0) = v * ((g - 1) * usq - H) + By * Bdotu / rho;
        IMHD_M(7, 1) = (1 - g) * u * v - Bx * By / rho;
        IMHD_M(7, 2) = H + (1 - g) * (v * v + 0.5f * usq) - By * By / rho;
        IMHD_M(7, 3) = (1 - g) * v * w - By * Bz / rho;
        IMHD_M(7, 4) = (2 - g) * v * Bx - u * By; IMHD_M(7, 5) = (1 - g) * v * By + By * Bdotu;  // slip
        IMHD_M(7, 6) = (2 - g) * v * Bz - w * By; IMHD_M(7, 7) = v * g;
    } else {
        IMHD_M(0, 3) = 1.0f;
        IMHD_M(1, 0) = -u * w; IMHD_M(1, 1) = w; IMHD_M(1, 3) = u; IMHD_M(1, 4) = -Bz; IMHD_M(1, 6) = -Bx;
        IMHD_M(2, 0) = -v * w; IMHD_M(2, 2) = w; IMHD_M(2, 3) = v; IMHD_M(2, 4) = -Bz; IMHD_M(2, 6) = -Bx;  // slip
        IMHD_M(3, 0) = 0.5f * (g - 1) * usq - w * w; IMHD_M(3, 1) = (1 - g) * u; IMHD_M(3, 2) = (1 - g) * v; IMHD_M(3, 3) = (3 - g) * w;
        IMHD_M(3, 4) = (2 - g) * Bx; IMHD_M(3, 5) = (2 - g) * By; IMHD_M(3, 6) = -g * By; IMHD_M(3, 7) = g - 1;  // slip
        IMHD_M(4, 0) = (w * Bx - u * Bz) / rho; IMHD_M(4, 1) = Bz / rho; IMHD_M(4, 3) = -Bx / rho; IMHD_M(4, 4) = -w; IMHD_M(4, 6) = u;
        IMHD_M(5, 0) = (w * By - v * Bz) / rho; IMHD_M(5, 2) = Bz / rho; IMHD_M(5, 3) = -By / rho; IMHD_M(5, 5) = -w; IMHD_M(5, 6) = v;
        IMHD_M(7, 0) = w * ((g - 1) * usq - H) + Bz * Bdotu / rho;
        IMHD_M(7, 1) = (1 - g) * u * w - Bx * Bz / rho; IMHD_M(7, 2) = (1 - g) * v * w - By * Bz / rho;
        IMHD_M(7, 3) = H + (1 - g) * (w * w + 0.5f * usq) - Bz * Bz / rho;
        IMHD_M(7, 4) = (2 - g) * w * Bx - u * Bz; IMHD_M(7, 5) = (2 - g) * w * By - v * Bz; IMHD_M(7, 6) = (1 - g) * w * Bz - Bdotu;
        IMHD_M(7, 7) = w * g;
    }
#undef IMHD_M
}

template <bool QUIRKS>
__device__ __forceinline__ float cell_lhs(const float U[8], float tx, float ty, float tz) {
    const float inv = fast_rcp(U[RHO]);
    const float u = U[MX] * inv, v = U[MY] * inv, w = U[MZ] * inv;
    const float usq = fmaf(w, w, fmaf(v, v, u * u));
    const float Bsq = fmaf(U[BZ], U[BZ], fmaf(U[BY], U[BY], U[BX] * U[BX]));
    // the scanner's pressure uses rho u^2 / 2 (compute_stability.cpp:206), unlike the solver's helper (B-1)
    const float p = kGm1f * (U[EN] - 0.5f * U[RHO] * usq - 0.5f * Bsq);
    const float a2 = (float)kGamma * p * inv, b2 = Bsq * inv;
    const float lx = radius(u, U[BX] * U[BX] * inv, a2, b2);
    float ly, lz;
    if (QUIRKS) {
        double M[9][9];
        reference_jacobian<1>(U, M);
        ly = (float)spectral_radius8(M);
        reference_jacobian<2>(U, M);
        lz = (float)spectral_radius8(M);
    } else {
        ly = radius(v, U[BY] * U[BY] * inv, a2, b2);
        lz = radius(w, U[BZ] * U[BZ] * inv, a2, b2);
    }
    return fmaf(tz, lz, fmaf(ty, ly, tx * lx));
}

// out[0]: (float bits of max LHS) << 32 | (0xFFFFFFFF - cell)  -> atomicMax picks the largest LHS and, among equal
//         ones, the first cell in the reference's scan order (k, i, j) = the smallest linear index
// out[1]: number of cells with LHS >= 1
// VEC = 4: four consecutive cells per thread and iteration through 16-byte loads (needs 16-byte aligned variable
// arrays and ncells % 4 == 0); VEC = 1 otherwise.
template <int VEC, bool QUIRKS>
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
            const float lhs = cell_lhs<QUIRKS>(U[e], tx, ty, tz);
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

static int g_mode = IMHD_STABILITY_WAVE_SPEEDS;
extern "C" void imhd_stability_mode(int mode) { g_mode = mode == IMHD_STABILITY_REFERENCE_QUIRKS ? mode : IMHD_STABILITY_WAVE_SPEEDS; }

// The scan without the wait: kernel + 16-byte D2H into `h` (pinned memory if the caller wants to go on) on `stream`.
int imhd_stability_scan_async(const float* Q, const imhd_slab* s, unsigned long long h[2], void* stream) {
    if (!Q || !s || !h) { set_error("imhd_stability_scan: null argument"); return IMHD_E_INVALID; }
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
    {   // keep freed blocks in the device's stream-ordered pool across synchronisations (default: trimmed at every sync,
        // which makes each cudaMallocAsync a driver allocation: +0.37 ms per scan)
        static unsigned long long tuned = 0;
        if (!(tuned & (1ull << (dev & 63)))) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = 1ull << 26;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            tuned |= 1ull << (dev & 63);
        }
    }
    unsigned long long* d_out = nullptr;   // 16 bytes from the stream-ordered pool: no process-global scratch, nothing to leak
    IMHD_CUDA(cudaMallocAsync(&d_out, 2 * sizeof(unsigned long long), st));
    IMHD_CUDA(cudaMemsetAsync(d_out, 0, 2 * sizeof(unsigned long long), st));
    const bool vec4 = ncells % 4 == 0 && vs % 4 == 0 && first % 4 == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0;
    const long long want = (ncells / (vec4 ? 4 : 1) + 255) / 256;
    const unsigned grid = (unsigned)(want < (long long)sms * 8 ? want : (long long)sms * 8);  // 8 resident blocks of 256 per SM
    const float tx = s->dt / s->dx, ty = s->dt / s->dy, tz = s->dt / s->dz;
    if (g_mode == IMHD_STABILITY_REFERENCE_QUIRKS) {
        const long long w1 = (ncells + 255) / 256;
        k_stability<1, true><<<(unsigned)(w1 < (long long)sms * 8 ? w1 : (long long)sms * 8), 256, 0, st>>>(Q, vs, first, ncells, tx, ty, tz, d_out);
    } else if (vec4) {
        k_stability<4, false><<<grid, 256, 0, st>>>(Q, vs, first, ncells, tx, ty, tz, d_out);
    } else {
        k_stability<1, false><<<grid, 256, 0, st>>>(Q, vs, first, ncells, tx, ty, tz, d_out);
    }
    IMHD_LAUNCH_CHECK(1);
    IMHD_CUDA(cudaMemcpyAsync(h, d_out, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    IMHD_CUDA(cudaFreeAsync(d_out, st));
    return 0;
}

// The report of a finished scan (h as written by imhd_stability_scan_async, after its stream has passed the copy).
void imhd_stability_decode(const unsigned long long h[2], const imhd_slab* s, imhd_stability* host_out) {
    const long long plane = (long long)s->Nx * s->Ny;
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
}

extern "C" int imhd_stability_scan(const float* Q, const imhd_slab* s, imhd_stability* host_out, void* stream) {
    if (!host_out) { set_error("imhd_stability_scan: null argument"); return IMHD_E_INVALID; }
    unsigned long long h[2] = {0, 0};
    if (int e = imhd_stability_scan_async(Q, s, h, stream)) return e;
    IMHD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    imhd_stability_decode(h, s, host_out);
    return 0;
}
