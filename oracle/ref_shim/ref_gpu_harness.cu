// TEST / MEASUREMENT INFRASTRUCTURE ONLY -- not product code.
//
// Headless GPU driver for the reference's UNMODIFIED on-device kernels, compiled by nvcc for
// sm_100 from the sources where they lie under /root/reference (oracle/Makefile target
// `refgpu`; output oracle/_ref/libimhd_ref_gpu*.so).  It answers SURVEY.md 8(d) "reference GPU
// number on the same box" (the kernels to beat, recompiled for B200) and gives the parity tests
// a second pin: the reference itself, run on the B200.
//
// It reproduces the launch ORDER and the device-wide synchronisation of the two shipped drivers
// without shm / fork / HDF5 / D2H:
//   Path A  /root/reference/src/on-device/no_diffusion.cu:174-199 (prime), :288-316 (step)
//   Path B  /root/reference/src/on-device/main.cu:108-112 (prime), :200-213 (step)
// Launch GEOMETRY is a parameter:
//   geom[0..2]  thread-block dims of the volume kernels (FA_*threads / fluidblockdims / intvarblockdims)
//   geom[3]     0 = "stock": grid = SM_mult x numberOfSMs per axis exactly as the drivers compute it
//                   (no_diffusion.cu:152, main.cu:73-75) -- on a 148-SM part that is 444^3 or 148^3 blocks;
//               1 = "cover": the smallest grid that covers the domain (one-thread-per-cell kernels)
//                   or min(stock, cover) (grid-stride kernels) -- the friendliest launch the
//                   reference's kernels admit
//   geom[4..6]  SM multipliers per axis for "stock"
//   geom[7]     path B only: 1 = BoundaryConditions with z-extent 1 (deterministic),
//               0 = the driver's racy full-grid launch
// Only tests/, bench.py's reference legs and tools/ may load this library.
#include <cuda_runtime.h>

#include <algorithm>

#include "initialize_od.cuh"
#include "kernels_fluidbcs.cuh"
#include "kernels_intvarbcs.cuh"
#include "kernels_od.cuh"
#include "kernels_od_intvar.cuh"

namespace {

int g_sms = 0;

int sms() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sms;
}

unsigned cdiv(int a, int b) { return (unsigned)((a + b - 1) / b); }

struct Geo {
    dim3 block, grid_cells, grid_stride;
};

Geo make_geo(const int* g, int Nx, int Ny, int Nz) {
    Geo o;
    o.block = dim3(g[0], g[1], g[2]);
    dim3 stock(g[4] * sms(), g[5] * sms(), g[6] * sms());
    dim3 cover(cdiv(Nx, g[0]), cdiv(Ny, g[1]), cdiv(Nz, g[2]));
    if (g[3] == 0) {
        o.grid_cells = stock;
        o.grid_stride = stock;
    } else {
        o.grid_cells = cover;
        o.grid_stride = dim3(std::min(stock.x, cover.x), std::min(stock.y, cover.y), std::min(stock.z, cover.z));
    }
    return o;
}

// Boundary launches of no_diffusion.cu:133-148: numberOfSMs blocks per live axis; block dims from
// input.inp (16x16 faces, 32 lines).  With "cover" they are sized to the domain instead.
struct BGeo {
    dim3 g_lr, g_tb, g_fb, g_fr, g_fbo, g_br, b_lr, b_tb, b_fb, b_fr, b_fbo, b_br;
};

BGeo make_bgeo(int cover, int Nx, int Ny, int Nz) {
    BGeo o;
    o.b_lr = dim3(16, 1, 16);
    o.b_tb = dim3(1, 16, 16);
    o.b_fb = dim3(16, 16, 1);
    o.b_fr = dim3(32, 1, 1);
    o.b_fbo = dim3(1, 32, 1);
    o.b_br = dim3(1, 1, 32);
    unsigned S = sms();
    unsigned gx = cover ? cdiv(Nx, 16) : S, gy = cover ? cdiv(Ny, 16) : S, gz = cover ? cdiv(Nz, 16) : S;
    o.g_lr = dim3(gx, 1, gz);
    o.g_tb = dim3(1, gy, gz);
    o.g_fb = dim3(gx, gy, 1);
    o.g_fr = dim3(cover ? cdiv(Nx, 32) : S, 1, 1);
    o.g_fbo = dim3(1, cover ? cdiv(Ny, 32) : S, 1);
    o.g_br = dim3(1, 1, cover ? cdiv(Nz, 32) : S);
    return o;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, 0); }
    float stop() {
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

void qint_bdry_A(const float* Q, float* Qint, float dt, float dx, float dy, float dz, int Nx, int Ny, int Nz,
                 const BGeo& b) {
    QintBdryFrontNoDiff<<<b.g_fb, b.b_fb>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    QintBdryLeftRightNoDiff<<<b.g_lr, b.b_lr>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    QintBdryTopBottomNoDiff<<<b.g_tb, b.b_tb>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    QintBdryFrontBottomNoDiff<<<b.g_fbo, b.b_fbo>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    QintBdryFrontRightNoDiff<<<b.g_fr, b.b_fr>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    QintBdryBottomRightNoDiff<<<b.g_br, b.b_br>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    cudaDeviceSynchronize();
    QintBdryPBCs<<<b.g_fb, b.b_fb>>>(Q, Qint, Nx, Ny, Nz);
    cudaDeviceSynchronize();
}

}  // namespace

extern "C" {

int refgpu_num_sms() { return sms(); }

// One-thread-per-cell kernels silently skip cells beyond the launch extent:
// returns 1 when geometry g covers an Nx x Ny x Nz domain for those kernels.
int refgpu_covers(const int* g, int Nx, int Ny, int Nz) {
    Geo o = make_geo(g, Nx, Ny, Nz);
    return (long long)o.grid_cells.x * o.block.x >= Nx && (long long)o.grid_cells.y * o.block.y >= Ny &&
           (long long)o.grid_cells.z * o.block.z >= Nz;
}

// ic: 0 ScrewPinchStride(a=J0) 1 CubicBennettVortex_m0(a=k,b=A) 2 CubicBennettVortex 3 ZPinch(a=r_max_coeff)
//     4 ScrewPinch(a=J0,b=r_max_coeff); launch geometry of no_diffusion.cu:127-131 (8x8x8 blocks) sized to the domain
int refgpu_init(int ic, float* Q, float a, float b, const float* x, const float* y, const float* z, int Nx, int Ny,
                int Nz) {
    dim3 block(8, 8, 8), grid(cdiv(Nx, 8), cdiv(Ny, 8), cdiv(Nz, 8));
    switch (ic) {
        case 0: ScrewPinchStride<<<grid, block>>>(Q, a, x, y, z, Nx, Ny, Nz); break;
        case 1: CubicBennettVortex_m0<<<grid, block>>>(Q, a, b, x, y, z, Nx, Ny, Nz); break;
        case 2: CubicBennettVortex<<<grid, block>>>(Q, x, y, z, Nx, Ny, Nz); break;
        case 3: ZPinch<<<grid, block>>>(Q, a, x, y, z, Nx, Ny, Nz); break;
        case 4: ScrewPinch<<<grid, block>>>(Q, a, b, x, y, z, Nx, Ny, Nz); break;
        default: return -1;
    }
    cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}

int refgpu_pathA_prime(float* Q, float* Qint, float dt, float dx, float dy, float dz, int Nx, int Ny, int Nz,
                       const int* g) {
    Geo o = make_geo(g, Nx, Ny, Nz);
    BGeo b = make_bgeo(g[3], Nx, Ny, Nz);
    rigidConductingWallBCsLeftRight<<<b.g_lr, b.b_lr>>>(Q, Nx, Ny, Nz);
    rigidConductingWallBCsTopBottom<<<b.g_tb, b.b_tb>>>(Q, Nx, Ny, Nz);
    PBCs<<<b.g_fb, b.b_fb>>>(Q, Nx, Ny, Nz);
    cudaDeviceSynchronize();
    ComputeIntermediateVariablesNoDiff<<<o.grid_cells, o.block>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    cudaDeviceSynchronize();
    qint_bdry_A(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz, b);
    return (int)cudaGetLastError();
}

// ms[0..3] accumulate corrector, PBCs, predictor, Qint boundary passes; returns total ms in ms[4].
int refgpu_pathA_steps(float* Q, float* Qint, int nsteps, float dt, float dx, float dy, float dz, int Nx, int Ny,
                       int Nz, const int* g, float* ms) {
    Geo o = make_geo(g, Nx, Ny, Nz);
    BGeo b = make_bgeo(g[3], Nx, Ny, Nz);
    Timer t, all;
    for (int q = 0; q < 5; ++q) ms[q] = 0.f;
    all.start();
    for (int s = 0; s < nsteps; ++s) {
        t.start();
        FluidAdvanceLocalNoDiff<<<o.grid_cells, o.block>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
        cudaDeviceSynchronize();
        ms[0] += t.stop();
        t.start();
        PBCs<<<b.g_fb, b.b_fb>>>(Q, Nx, Ny, Nz);
        cudaDeviceSynchronize();
        ms[1] += t.stop();
        t.start();
        ComputeIntermediateVariablesNoDiff<<<o.grid_cells, o.block>>>(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
        cudaDeviceSynchronize();
        ms[2] += t.stop();
        t.start();
        qint_bdry_A(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz, b);
        ms[3] += t.stop();
    }
    ms[4] = all.stop();
    return (int)cudaGetLastError();
}

int refgpu_pathB_prime(float* Q, float* Qint, float D, float dt, float dx, float dy, float dz, int Nx, int Ny,
                       int Nz, const int* g) {
    Geo o = make_geo(g, Nx, Ny, Nz);
    ComputeIntermediateVariablesStride<<<o.grid_stride, o.block>>>(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
    cudaDeviceSynchronize();
    // one-thread-per-cell mega-kernel: needs the covering grid (main.cu launches it with the intvar grid)
    ComputeIntermediateVariablesBoundary<<<o.grid_cells, o.block>>>(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
    cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}

// ms[0..3]: corrector, BoundaryConditions, predictor, Qint boundary mega-kernel; ms[4] total.
int refgpu_pathB_steps(float* Q, float* Qint, int nsteps, float D, float dt, float dx, float dy, float dz, int Nx,
                       int Ny, int Nz, const int* g, float* ms) {
    Geo o = make_geo(g, Nx, Ny, Nz);
    dim3 bc_block = g[7] ? dim3(o.block.x, o.block.y, 1) : o.block;
    dim3 bc_grid = g[7] ? dim3(o.grid_cells.x, o.grid_cells.y, 1) : o.grid_cells;
    Timer t, all;
    for (int q = 0; q < 5; ++q) ms[q] = 0.f;
    all.start();
    for (int s = 0; s < nsteps; ++s) {
        t.start();
        FluidAdvanceLocal<<<o.grid_stride, o.block>>>(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
        cudaDeviceSynchronize();
        ms[0] += t.stop();
        t.start();
        BoundaryConditions<<<bc_grid, bc_block>>>(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
        cudaDeviceSynchronize();
        ms[1] += t.stop();
        t.start();
        ComputeIntermediateVariablesStride<<<o.grid_stride, o.block>>>(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
        cudaDeviceSynchronize();
        ms[2] += t.stop();
        t.start();
        ComputeIntermediateVariablesBoundary<<<o.grid_cells, o.block>>>(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
        cudaDeviceSynchronize();
        ms[3] += t.stop();
    }
    ms[4] = all.stop();
    return (int)cudaGetLastError();
}

}  // extern "C"
