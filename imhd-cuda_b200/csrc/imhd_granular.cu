// Parity-granular operators: one kernel group per reference kernel group, same cell sets, same
// fp32/fp64 rounding points (EXACT recipe of imhd_math.cuh; this translation unit is compiled
// with -fmad=false so no product+sum is contracted).  One thread per cell, global memory only --
// these exist to be compared call by call with the reference, not to be fast; the fused
// kernels in imhd_fused.cu are the hot path.
//
// Cell classes and quirks follow SURVEY.md A.3/A.4 and Appendix B; reference file:line are
// given at each kernel.
#include "imhd_common.cuh"

namespace imhd {

__device__ __forceinline__ void load8(const float* __restrict__ A, long long l, long long vs, float U[8]) {
#pragma unroll
    for (int v = 0; v < 8; ++v) U[v] = A[l + v * vs];
}

// -----------------------------------------------------------------------------------------------
// Predictor, all cell classes of planes k <= Nz-2.
//   generic  intRho..intE            lib/on-device/kernels_od_intvar.cu:1160-1253
//   Right / Bottom / BottomRight     lib/on-device/kernels_intvarbcs.cu:560-738, 1024-1110
//   FrontRight / FrontBottom typos   lib/on-device/kernels_intvarbcs.cu:923, 973, 1005, 1017 (B-15)
//   path B diffusion on [1,N-2]^3    lib/on-device/kernels_od_intvar.cu:130-145
// -----------------------------------------------------------------------------------------------
template <int PATH>
__global__ void __launch_bounds__(256) k_predictor(const float* __restrict__ Q, float* __restrict__ Qint, Params P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= P.Nx || j >= P.Ny || k >= P.Nz - 1) return;
    const long long vs = P.cube, l = (long long)k * P.plane + (long long)i * P.Ny + j;
    const bool bottom = (i == P.Nx - 1), right = (j == P.Ny - 1);
    const bool frontright = (k == 0 && right && !bottom), frontbottom = (k == 0 && bottom && !right);

    float U[8], f[8], g[8], h[8], dF[8], dG[8], dH[8], t[8];
    load8(Q, l, vs, U);
    const Aux<true> a = make_aux<true>(U);
    flux_indexed<true, DIR_X>(U, a, f);
    flux_indexed<true, DIR_Y>(U, a, g);
    flux_indexed<true, DIR_Z>(U, a, h);
    if (!bottom) {
        float V[8];
        load8(Q, l + P.Ny, vs, V);
        flux_indexed<true, DIR_X>(V, make_aux<true>(V), t);
#pragma unroll
        for (int v = 0; v < 8; ++v) dF[v] = t[v] - f[v];
    } else {
#pragma unroll
        for (int v = 0; v < 8; ++v) dF[v] = -f[v];
    }
    float gp_bz = 0.0f;  // G(Bz) at j+1, for the FrontBottom typo
    if (!right) {
        float V[8];
        load8(Q, l + 1, vs, V);
        flux_indexed<true, DIR_Y>(V, make_aux<true>(V), t);
        gp_bz = t[BZ];
#pragma unroll
        for (int v = 0; v < 8; ++v) dG[v] = t[v] - g[v];
    } else {
#pragma unroll
        for (int v = 0; v < 8; ++v) dG[v] = -g[v];
    }
    {
        float V[8];
        load8(Q, l + P.plane, vs, V);
        flux_indexed<true, DIR_Z>(V, make_aux<true>(V), t);
#pragma unroll
        for (int v = 0; v < 8; ++v) dH[v] = t[v] - h[v];
    }
    if (frontright) dF[EN] = f[EN] - f[EN];
    if (frontbottom) {
        dH[MZ] = t[MZ] - g[MZ];
        dH[EN] = t[EN] - g[EN];
        dG[BZ] = gp_bz - gp_bz;
    }
    float base[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) base[v] = U[v];
    if (bottom && right) base[MZ] = U[MX];

    const bool diffuse = PATH == IMHD_PATH_B && i > 0 && i < P.Nx - 1 && j > 0 && j < P.Ny - 1 && k > 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        float r = base[v] - P.tx * dF[v] - P.ty * dG[v] - P.tz * dH[v];
        if (diffuse) {
            const float* q = Q + v * vs;
            r = r + P.dt * num_diff<true>(U[v], q[l + P.Ny], q[l + 1], q[l + P.plane], q[l - P.Ny], q[l - 1],
                                          q[l - P.plane], P.D, P.dc);
        }
        Qint[l + v * vs] = r;
    }
}

// dst plane <- src plane for all 8 variables (QintBdryPBCs kernels_intvarbcs.cu:383-398; PBCs
// kernels_fluidbcs.cu:498-510).
__global__ void k_copy_plane(float* __restrict__ A, long long dst_off, long long src_off, long long plane,
                             long long vs) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plane) return;
#pragma unroll
    for (int v = 0; v < 8; ++v) A[dst_off + c + v * vs] = A[src_off + c + v * vs];
}

// -----------------------------------------------------------------------------------------------
// Corrector over the volume, in place (each thread reads only its own Q cell).
//   A: FluidAdvanceLocalNoDiff kernels_od.cu:353-525, cells i,j,k >= 1 including far faces (B-14)
//   B: FluidAdvanceLocal       kernels_od.cu:82-350,  [1,N-2]^3, + dt*numericalDiffusionLocal(Qint)
//   LaxWendroffAdv*Local kernels_od.cu:1206-1332; mixed-neighbour Bsq / Bdotu at k-1 (B-4, B-5);
//   path B rho update takes rho_int_im1 for rhovx_int_im1 (B-6).
// -----------------------------------------------------------------------------------------------
template <int PATH>
__global__ void __launch_bounds__(256) k_corrector(float* __restrict__ Q, const float* __restrict__ Qint, Params P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    const int iend = PATH == IMHD_PATH_A ? P.Nx : P.Nx - 1, jend = PATH == IMHD_PATH_A ? P.Ny : P.Ny - 1;
    const int kend = PATH == IMHD_PATH_A ? P.Nz : P.Nz - 1;
    if (i < 1 || j < 1 || k < 1 || i >= iend || j >= jend || k >= kend) return;
    const long long vs = P.cube, l = (long long)k * P.plane + (long long)i * P.Ny + j;

    float q[8], c[8], xi[8], yj[8], zk[8];
    load8(Q, l, vs, q);
    load8(Qint, l, vs, c);
    load8(Qint, l - P.Ny, vs, xi);
    load8(Qint, l - 1, vs, yj);
    load8(Qint, l - P.plane, vs, zk);

    const Aux<true> ac = make_aux<true>(c), ai = make_aux<true>(xi), aj = make_aux<true>(yj);
    // k-1 point: KE of zk, Bsq of (Bx(i-1), By(j-1), Bz(k-1)), Bdotu with rhovy(j-1)
    const float KEk = h_KE<true>(zk[RHO], zk[MX], zk[MY], zk[MZ], 0.f);
    const float Bk = h_Bsq<true>(xi[BX], yj[BY], zk[BZ]);
    const float pk = h_p<true>(zk[EN], Bk, KEk);
    const float Dk = h_Bdotu<true>(zk[RHO], zk[MX], yj[MY], zk[MZ], zk[BX], zk[BY], zk[BZ], 0.f);

    float fc[8], gc[8], hc[8], fi[8], gj[8], hk[8];
    flux_local<true, DIR_X>(c, ac.p, ac.Bsq, ac.Bdotu, 0.f, fc);
    flux_local<true, DIR_Y>(c, ac.p, ac.Bsq, ac.Bdotu, 0.f, gc);
    flux_local<true, DIR_Z>(c, ac.p, ac.Bsq, ac.Bdotu, 0.f, hc);
    flux_local<true, DIR_X>(xi, ai.p, ai.Bsq, ai.Bdotu, 0.f, fi);
    flux_local<true, DIR_Y>(yj, aj.p, aj.Bsq, aj.Bdotu, 0.f, gj);
    flux_local<true, DIR_Z>(zk, pk, Bk, Dk, 0.f, hk);
    if (PATH == IMHD_PATH_B) fi[RHO] = xi[RHO];  // B-6

#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float dF = fc[v] - fi[v], dG = gc[v] - gj[v], dH = hc[v] - hk[v];
        float r = (float)(0.5 * (q[v] + c[v]) - 0.5 * P.tx * dF - 0.5 * P.ty * dG - 0.5 * P.tz * dH);
        if (PATH == IMHD_PATH_B) {
            const float* qi = Qint + v * vs;
            r = r + P.dt * num_diff<true>(c[v], qi[l + P.Ny], qi[l + 1], qi[l + P.plane], xi[v], yj[v], zk[v], P.D, P.dc);
        }
        Q[l + v * vs] = r;
    }
}

// -----------------------------------------------------------------------------------------------
// Path B fluid boundary pass: BoundaryConditions (kernels_fluidbcs.cu:32-235), single-application
// semantics.  One thread per (i,j) of the k=0 plane:
//   interior (i,j): corrector with INDEXED fluxes of Qint, k-1 -> Nz-2 (B-10), + dt*diffusion
//                   (diffusion.cu:79-106), whole expression fp64 (:53-116)
//   i = 0, Nx-1   : wall values, all j (:164-188); j-walls are dead code (B-8)
//   (Nx-1,Ny-1)   : copy the column's k=0 values to k=Nz-1 (:227-231, B-8)
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_boundary_conditions(float* __restrict__ Q, const float* __restrict__ Qint, Params P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= P.Nx || j >= P.Ny) return;
    const long long vs = P.cube, l = (long long)i * P.Ny + j;
    if (i > 0 && i < P.Nx - 1) {
        if (j == 0 || j == P.Ny - 1) return;
        const long long lb = l + (long long)(P.Nz - 2) * P.plane;
        float c[8], V[8], f[8], g[8], h[8], t[8], dF[8], dG[8], dH[8];
        load8(Qint, l, vs, c);
        const Aux<true> a = make_aux<true>(c);
        flux_indexed<true, DIR_X>(c, a, f);
        flux_indexed<true, DIR_Y>(c, a, g);
        flux_indexed<true, DIR_Z>(c, a, h);
        load8(Qint, l - P.Ny, vs, V);
        flux_indexed<true, DIR_X>(V, make_aux<true>(V), t);
#pragma unroll
        for (int v = 0; v < 8; ++v) dF[v] = f[v] - t[v];
        load8(Qint, l - 1, vs, V);
        flux_indexed<true, DIR_Y>(V, make_aux<true>(V), t);
#pragma unroll
        for (int v = 0; v < 8; ++v) dG[v] = g[v] - t[v];
        load8(Qint, lb, vs, V);
        flux_indexed<true, DIR_Z>(V, make_aux<true>(V), t);
#pragma unroll
        for (int v = 0; v < 8; ++v) dH[v] = h[v] - t[v];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const float* qi = Qint + v * vs;
            const float nd = num_diff<true>(c[v], qi[l + P.Ny], qi[l + 1], qi[l + P.plane], qi[l - P.Ny], qi[l - 1],
                                            V[v], P.D, P.dc);
            float* q = Q + v * vs;
            q[l] = (float)(0.5 * (q[l] + c[v]) - 0.5 * P.tx * dF[v] - 0.5 * P.ty * dG[v] - 0.5 * P.tz * dH[v] +
                           P.dt * nd);
        }
        return;
    }
    // wall cells (0,j,0) and (Nx-1,j,0)
    Q[l] = 1.0f;
#pragma unroll
    for (int v = 1; v < 7; ++v) Q[l + v * vs] = 0.0f;
    // every x-thread of the reference launch re-applies e <- p(e,0,0)/(gamma-1); iterate to the
    // fixed point (reached after <= 2 applications; e == 0 on the walls of the shipped ICs)
    float e = Q[l + EN * vs];
    for (int rep = 0; rep < P.Nx; ++rep) {
        const float e2 = wall_e(e);
        if (e2 == e) break;
        e = e2;
    }
    Q[l + EN * vs] = e;
    if (i == P.Nx - 1 && j == P.Ny - 1) {
        const long long lt = l + (long long)(P.Nz - 1) * P.plane;
        Q[lt] = 1.0f;
#pragma unroll
        for (int v = 1; v < 7; ++v) Q[lt + v * vs] = 0.0f;
        Q[lt + EN * vs] = e;
    }
}

// Path A init: rigidConductingWallBCsLeftRight (kernels_fluidbcs.cu:436-464): j = 0 and Ny-1 for
// every i, 0 < k < Nz-1.  (rigidConductingWallBCsTopBottom is a no-op under the shipped launch: B-11.)
__global__ void k_wall_leftright(float* __restrict__ Q, Params P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= P.Nx || k < 1 || k >= P.Nz - 1) return;
    const long long vs = P.cube;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const long long l = (long long)k * P.plane + (long long)i * P.Ny + (side ? P.Ny - 1 : 0);
        Q[l] = 1.0f;
#pragma unroll
        for (int v = 1; v < 7; ++v) Q[l + v * vs] = 0.0f;
        Q[l + EN * vs] = wall_e(Q[l + EN * vs]);
    }
}

// -----------------------------------------------------------------------------------------------
// Grids and initial conditions (lib/on-device/initialize_od.cu:26-57, 269-345, 132-205)
// -----------------------------------------------------------------------------------------------
__global__ void k_init_axis(float* __restrict__ g, float lo, float d, int n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (unsigned)n) g[i] = lo + i * d;
}

__device__ __forceinline__ float total_energy(const float U[8], float p) {  // :327-334
    return (float)((p / kGm1) + (sqd(U[MX]) + sqd(U[MY]) + sqd(U[MZ])) / (2.0 * U[RHO]) +
                   0.5 * (sqd(U[BX]) + sqd(U[BY]) + sqd(U[BZ])));
}

// IC: 0 ScrewPinchStride(a=J0)  1 CubicBennettVortex_m0(a=k,b=A)  2 CubicBennettVortex  3 ZPinch(a=r_max_coeff)
//     4 ScrewPinch(a=J0,b=r_max_coeff)   (initialize_od.cu:269, 132, 59, 347, 207)
template <int IC>
__global__ void __launch_bounds__(256) k_init_state(float* __restrict__ Q, float a, float b, const float* __restrict__ gx,
                                                    const float* __restrict__ gy, const float* __restrict__ gz, Params P,
                                                    int kofs = 0) {
    // kofs: global index of array plane 0 (z-slab arrays of the multi-GPU engine); gz is the array's own z grid
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= P.Nx || j >= P.Ny) return;
    const long long vs = P.cube, l = (long long)k * P.plane + (long long)i * P.Ny + j;
    const float r_max = sqrtf((float)(sqd(gx[P.Nx - 1]) + sqd(gy[P.Ny - 1])));
    // 0.25 is a double literal in the stride / Bennett kernels; r_max_coeff is an fp32 argument (:219, :359)
    const float r_pinch = (IC == 3) ? a * r_max : (IC == 4) ? b * r_max : (float)(0.25 * r_max);
    const float x = gx[i], y = gy[j];
    const float r = sqrtf((float)(sqd(x) + sqd(y)));
    if (IC == 4 && !(r < r_pinch)) {  // ScrewPinch writes only rho outside the pinch (:237); the rest is left as found
        Q[l] = 0.1f;
        return;
    }
    float U[8] = {0.01f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < r_pinch) {
        const double r2 = sqd(r), rp2 = sqd(r_pinch);
        const float Br = 0.0f;
        if (IC == 0 || IC == 4) {
            const float J0 = a;
            const float Jr = 0.0f, Jphi = 0.0f;
            const float Btheta = (float)(0.5 * J0 * r * (1.0 - 0.5 * r2 / rp2));  // :313 / :243
            const double rp4 = rp2 * rp2, r4 = r2 * r2, r6 = r4 * r2;             // pow(.,4), pow(.,6)
            const float p = (float)(-0.25 * (sqd(J0) / rp4) * (r6 / 6.0 - 0.75 * rp2 * r4 + rp4 * r2));  // :316 / :246
            U[RHO] = 1.0f;
            U[MX] = Jr * x - Jphi * y / r;
            U[MY] = Jr * y + Jphi * x / r;
            U[MZ] = (float)(J0 * (1 - r2 / rp2));
            U[BX] = Br * x - Btheta * y / r;
            U[BY] = Br * y + Btheta * x / r;
            U[BZ] = 1.0f;
            U[EN] = total_energy(U, p);
        } else if (IC == 1 || IC == 2) {
            const float phi = r;
            const double phi3 = r2 * (double)phi;
            const float Btheta = (float)(-(1) * (phi3 - 3 * r2 - 6 * phi + 6 * (phi + 1) * logf(phi + 1)) /
                                         (2 * phi * (phi + 1)));                       // :177 / :101
            if (IC == 1) {
                const float z = gz[k];
                const float p = (float)(1 - phi3 / sqd(phi + 1) * (phi - 10));      // :180
                U[RHO] = (float)(1.0 + b * cosf((k + kofs) * z));  // the z loop index shadows the wavenumber (:158,183)
                U[MZ] = (float)((1) * r2 / sqd(phi + 1));
                U[BX] = Br * x - Btheta * y / phi;
                U[BY] = Br * y + Btheta * x / phi;
                U[EN] = total_energy(U, p);
            } else {
                const float p = (float)phi3;                                         // :103
                U[RHO] = 1.0f;
                U[MZ] = (float)(r2 / sqd(phi + 1));
                U[BX] = Br * x - Btheta * y / phi;
                U[BY] = Br * y + Btheta * x / phi;
                U[EN] = total_energy(U, p);
            }
        } else {  // ZPinch
            const double r3 = r2 * (double)r, r4 = r2 * r2, r6 = r4 * r2, rp4 = rp2 * rp2, rp6 = rp4 * rp2;
            const float Btheta = (float)(0.5 * (r + 0.5 * r3 / rp2));                                           // :394
            const float p = (float)(1 + 0.5 * (0.5 * r2 / rp2 + 0.375 * r4 / rp4 - (1.0 / 12.0) * r6 / rp6));  // :398
            U[RHO] = 1.0f;
            U[MZ] = (float)(1 + r2 / rp2);
            U[BX] = Br * x - Btheta * y / r;
            U[BY] = Br * y + Btheta * x / r;
            U[EN] = total_energy(U, p);
        }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) Q[l + v * vs] = U[v];
}

}  // namespace imhd

// =================================================================================================
// C ABI
// =================================================================================================
using namespace imhd;

static inline dim3 cell_grid(const Params& P, int nz) { return dim3((P.Ny + 31) / 32, (P.Nx + 7) / 8, nz); }
static const dim3 kCellBlock(32, 8, 1);

extern "C" int imhd_predictor(const float* Q, float* Qint, int path, float D, float dt, float dx, float dy,
                              float dz, int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const Params P = make_params(path, D, dt, dx, dy, dz, Nx, Ny, Nz);
    if (path == IMHD_PATH_A) k_predictor<IMHD_PATH_A><<<cell_grid(P, Nz - 1), kCellBlock, 0, st>>>(Q, Qint, P);
    else                     k_predictor<IMHD_PATH_B><<<cell_grid(P, Nz - 1), kCellBlock, 0, st>>>(Q, Qint, P);
    IMHD_LAUNCH_CHECK(1);
    k_copy_plane<<<(unsigned)((P.plane + 255) / 256), 256, 0, st>>>(Qint, (long long)(Nz - 1) * P.plane, 0, P.plane, P.cube);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_corrector(float* Q, const float* Qint, int path, float D, float dt, float dx, float dy,
                              float dz, int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const Params P = make_params(path, D, dt, dx, dy, dz, Nx, Ny, Nz);
    if (path == IMHD_PATH_A) k_corrector<IMHD_PATH_A><<<cell_grid(P, Nz), kCellBlock, 0, st>>>(Q, Qint, P);
    else                     k_corrector<IMHD_PATH_B><<<cell_grid(P, Nz), kCellBlock, 0, st>>>(Q, Qint, P);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_fluid_bcs(float* Q, const float* Qint, int path, float D, float dt, float dx, float dy,
                              float dz, int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const Params P = make_params(path, D, dt, dx, dy, dz, Nx, Ny, Nz);
    if (path == IMHD_PATH_A)
        k_copy_plane<<<(unsigned)((P.plane + 255) / 256), 256, 0, st>>>(Q, 0, (long long)(Nz - 1) * P.plane, P.plane, P.cube);
    else
        k_boundary_conditions<<<cell_grid(P, 1), kCellBlock, 0, st>>>(Q, Qint, P);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_initial_bcs(float* Q, int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const Params P = make_params(IMHD_PATH_A, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, Nz);
    k_wall_leftright<<<dim3((Nx + 31) / 32, (Nz + 7) / 8), dim3(32, 8), 0, st>>>(Q, P);
    IMHD_LAUNCH_CHECK(1);
    k_copy_plane<<<(unsigned)((P.plane + 255) / 256), 256, 0, st>>>(Q, 0, (long long)(Nz - 1) * P.plane, P.plane, P.cube);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_init_grids(float* x, float* y, float* z, float x_min, float x_max, float y_min, float y_max,
                               float z_min, float z_max, int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const float dx = (x_max - x_min) / (Nx - 1), dy = (y_max - y_min) / (Ny - 1), dz = (z_max - z_min) / (Nz - 1);
    k_init_axis<<<(Nx + 127) / 128, 128, 0, st>>>(x, x_min, dx, Nx);
    k_init_axis<<<(Ny + 127) / 128, 128, 0, st>>>(y, y_min, dy, Ny);
    k_init_axis<<<(Nz + 127) / 128, 128, 0, st>>>(z, z_min, dz, Nz);
    IMHD_LAUNCH_CHECK(3);
    return 0;
}

extern "C" int imhd_init_screwpinch_stride(float* Q, float J0, const float* x, const float* y, const float* z,
                                           int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    const Params P = make_params(0, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, Nz);
    k_init_state<0><<<cell_grid(P, Nz), kCellBlock, 0, (cudaStream_t)stream>>>(Q, J0, 0.f, x, y, z, P);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_init_cubic_bennett_vortex_m0(float* Q, float k, float A, const float* x, const float* y,
                                                 const float* z, int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    const Params P = make_params(0, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, Nz);
    k_init_state<1><<<cell_grid(P, Nz), kCellBlock, 0, (cudaStream_t)stream>>>(Q, k, A, x, y, z, P);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_init_cubic_bennett_vortex(float* Q, const float* x, const float* y, const float* z, int Nx, int Ny,
                                              int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    const Params P = make_params(0, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, Nz);
    k_init_state<2><<<cell_grid(P, Nz), kCellBlock, 0, (cudaStream_t)stream>>>(Q, 0.f, 0.f, x, y, z, P);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_init_zpinch(float* Q, float r_max_coeff, const float* x, const float* y, const float* z, int Nx,
                                int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    const Params P = make_params(0, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, Nz);
    k_init_state<3><<<cell_grid(P, Nz), kCellBlock, 0, (cudaStream_t)stream>>>(Q, r_max_coeff, 0.f, x, y, z, P);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_init_screwpinch(float* Q, float J0, float r_max_coeff, const float* x, const float* y,
                                    const float* z, int Nx, int Ny, int Nz, void* stream) {
    if (int e = bad_dims(Nx, Ny, Nz)) return e;
    const Params P = make_params(0, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, Nz);
    k_init_state<4><<<cell_grid(P, Nz), kCellBlock, 0, (cudaStream_t)stream>>>(Q, J0, r_max_coeff, x, y, z, P);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

// Internal (multi-GPU engine): initial condition `ic` (the IC ids of k_init_state) on a z-slab array of nz_array planes
// whose plane 0 is global plane kofs; z = the array's own grid.  Same kernel, same bits as the full-domain call.
int imhd_init_ic_slab(int ic, float* Q, float a, float b, const float* x, const float* y, const float* z, int Nx, int Ny,
                      int nz_array, int kofs, void* stream) {
    const Params P = make_params(0, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, nz_array);
    cudaStream_t st = (cudaStream_t)stream;
    switch (ic) {
        case 0: k_init_state<0><<<cell_grid(P, nz_array), kCellBlock, 0, st>>>(Q, a, b, x, y, z, P, kofs); break;
        case 1: k_init_state<1><<<cell_grid(P, nz_array), kCellBlock, 0, st>>>(Q, a, b, x, y, z, P, kofs); break;
        case 2: k_init_state<2><<<cell_grid(P, nz_array), kCellBlock, 0, st>>>(Q, a, b, x, y, z, P, kofs); break;
        case 3: k_init_state<3><<<cell_grid(P, nz_array), kCellBlock, 0, st>>>(Q, a, b, x, y, z, P, kofs); break;
        case 4: k_init_state<4><<<cell_grid(P, nz_array), kCellBlock, 0, st>>>(Q, a, b, x, y, z, P, kofs); break;
        default: set_error("unknown initial-condition id %d", ic); return IMHD_E_INVALID;
    }
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

// rigidConductingWallBCsLeftRight (kernels_fluidbcs.cu:436-465) on array planes [ka, kb) of a slab array
namespace imhd {
__global__ void k_wall_leftright_planes(float* __restrict__ Q, Params P, int ka, int kb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = ka + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= P.Nx || k >= kb) return;
    const long long vs = P.cube;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const long long l = (long long)k * P.plane + (long long)i * P.Ny + (side ? P.Ny - 1 : 0);
        Q[l] = 1.0f;
#pragma unroll
        for (int v = 1; v < 7; ++v) Q[l + v * vs] = 0.0f;
        Q[l + EN * vs] = wall_e(Q[l + EN * vs]);
    }
}
}  // namespace imhd
int imhd_wall_leftright_planes(float* Q, int Nx, int Ny, int nz_array, int ka, int kb, void* stream) {
    if (kb <= ka) return 0;
    const Params P = make_params(0, 0.f, 1.f, 1.f, 1.f, 1.f, Nx, Ny, nz_array);
    imhd::k_wall_leftright_planes<<<dim3((Nx + 15) / 16, (kb - ka + 15) / 16), dim3(16, 16), 0, (cudaStream_t)stream>>>(Q, P, ka, kb);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

// z grid of a slab array: plane kk holds z_min + (kofs + kk) * dz evaluated exactly like the full grid (k_init_axis)
namespace imhd {
__global__ void k_init_axis_ofs(float* __restrict__ g, float lo, float d, int n, int ofs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = lo + (unsigned)max(i + ofs, 0) * d;
}
}  // namespace imhd
int imhd_init_axis_slab(float* g, float lo, float d, int n, int ofs, void* stream) {
    imhd::k_init_axis_ofs<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, lo, d, n, ofs);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}
