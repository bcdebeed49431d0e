// TEST INFRASTRUCTURE ONLY -- not product code.
//
// Headless driver for the reference's UNMODIFIED on-device kernels compiled for
// the host through cuda_shim.h.  It reproduces the launch ORDER of the two
// shipped drivers without shm / fork / HDF5:
//   Path A  /root/reference/src/on-device/no_diffusion.cu:161-200 (init), :284-312 (step)
//   Path B  /root/reference/src/on-device/main.cu:102-112 (init), :196-213 (step)
// Launch GEOMETRY is chosen so every emulated launch covers the whole domain
//  and BoundaryConditions has z-extent 1, i.e.
// the deterministic single-application meaning of the racy k=0 face update.
//
// One emulated CUDA thread == one call of the kernel function with
// blockDim=(1,1,1), threadIdx=(0,0,0), blockIdx=(i,j,k).  Grid-stride kernels are
// called with gridDim=(1,1,T): host thread t owns k = t+1, t+1+T, ...
//
// Everything here is exported with C linkage and loaded through ctypes by
// oracle/oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#include "initialize_od.cuh"
#include "kernels_fluidbcs.cuh"
#include "kernels_intvarbcs.cuh"
#include "kernels_od.cuh"
#include "kernels_od_intvar.cuh"

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace {

void set_thread(unsigned i, unsigned j, unsigned k, unsigned gx, unsigned gy, unsigned gz) {
    blockDim = dim3(1, 1, 1);
    threadIdx = dim3(0, 0, 0);
    gridDim = dim3(gx, gy, gz);
    blockIdx = dim3(i, j, k);
}

// Run fn(t, T) on T host threads (t = 0..T-1).
void parallel(int T, const std::function<void(int, int)>& fn) {
    if (T <= 1) { fn(0, 1); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) pool.emplace_back(fn, t, T);
    for (auto& th : pool) th.join();
}

// One-thread-per-cell launch covering [0,Nx) x [0,Ny) x [0,Nz), z split over host threads.
template <class K>
void launch_cells(int Nx, int Ny, int Nz, int T, K kernel) {
    parallel(std::min(T, Nz), [&](int t, int nT) {
        int k0 = (int)((long long)Nz * t / nT), k1 = (int)((long long)Nz * (t + 1) / nT);
        for (int k = k0; k < k1; ++k)
            for (int i = 0; i < Nx; ++i)
                for (int j = 0; j < Ny; ++j) {
                    set_thread(i, j, k, Nx, Ny, Nz);
                    kernel();
                }
    });
}

// Grid-stride launch: gridDim=(1,1,T), each host thread is one emulated CUDA thread.
template <class K>
void launch_stride(int T, K kernel) {
    parallel(T, [&](int t, int nT) {
        set_thread(0, 0, t, 1, 1, nT);
        kernel();
    });
}

}  // namespace

extern "C" {

// ---- grids and initial conditions (initialize_od.cu) -----------------------
void ref_init_grids(float* x, float* y, float* z, float x_min, float x_max, float y_min,
                    float y_max, float z_min, float z_max, int Nx, int Ny, int Nz) {
    // dx as in main.cu:98-100 / no_diffusion.cu:106-108 (fp32)
    float dx = (x_max - x_min) / (Nx - 1);
    float dy = (y_max - y_min) / (Ny - 1);
    float dz = (z_max - z_min) / (Nz - 1);
    for (int i = 0; i < Nx; ++i) { set_thread(i, 0, 0, Nx, 1, 1); InitializeX(x, x_min, dx, Nx); }
    for (int j = 0; j < Ny; ++j) { set_thread(0, j, 0, 1, Ny, 1); InitializeY(y, y_min, dy, Ny); }
    for (int k = 0; k < Nz; ++k) { set_thread(0, 0, k, 1, 1, Nz); InitializeZ(z, z_min, dz, Nz); }
}

void ref_screwpinch_stride(float* Q, float J0, const float* x, const float* y, const float* z,
                           int Nx, int Ny, int Nz, int T) {
    launch_stride(T, [&] { ScrewPinchStride(Q, J0, x, y, z, Nx, Ny, Nz); });
}

void ref_cubic_bennett_vortex_m0(float* Q, float kwave, float A, const float* x, const float* y,
                                 const float* z, int Nx, int Ny, int Nz, int T) {
    launch_stride(T, [&] { CubicBennettVortex_m0(Q, kwave, A, x, y, z, Nx, Ny, Nz); });
}

// the three initial conditions the shipped drivers keep commented out (no_diffusion.cu:169-171)
void ref_cubic_bennett_vortex(float* Q, const float* x, const float* y, const float* z, int Nx, int Ny, int Nz, int T) {
    launch_stride(T, [&] { CubicBennettVortex(Q, x, y, z, Nx, Ny, Nz); });
}

void ref_zpinch(float* Q, float r_max_coeff, const float* x, const float* y, const float* z, int Nx, int Ny, int Nz,
                int T) {
    launch_stride(T, [&] { ZPinch(Q, r_max_coeff, x, y, z, Nx, Ny, Nz); });
}

void ref_screwpinch(float* Q, float J0, float r_max_coeff, const float* x, const float* y, const float* z, int Nx,
                    int Ny, int Nz, int T) {
    launch_cells(Nx, Ny, Nz, T, [&] { ScrewPinch(Q, J0, r_max_coeff, x, y, z, Nx, Ny, Nz); });
}

// ---- Path A granular kernels ------------------------------------------------
void ref_wall_bcs_leftright(float* Q, int Nx, int Ny, int Nz) {
    // <<<(S,1,S),(bx,1,bz)>>> : (x,z) threads (no_diffusion.cu:174)
    for (int k = 0; k < Nz; ++k)
        for (int i = 0; i < Nx; ++i) { set_thread(i, 0, k, Nx, 1, Nz); rigidConductingWallBCsLeftRight(Q, Nx, Ny, Nz); }
}

void ref_wall_bcs_topbottom(float* Q, int Nx, int Ny, int Nz) {
    // shipped launch <<<(1,S,S),(1,by,bz)>>> (no_diffusion.cu:135,144,175): x-extent 1, so the
    // kernel's j = threadIdx.x + blockDim.x*blockIdx.x is always 0 -> no-op
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j) { set_thread(0, j, k, 1, Ny, Nz); rigidConductingWallBCsTopBottom(Q, Nx, Ny, Nz); }
}

void ref_pbcs(float* Q, int Nx, int Ny, int Nz) {
    for (int i = 0; i < Nx; ++i)
        for (int j = 0; j < Ny; ++j) { set_thread(i, j, 0, Nx, Ny, 1); PBCs(Q, Nx, Ny, Nz); }
}

void ref_predictor_nodiff(const float* Q, float* Qint, float dt, float dx, float dy, float dz,
                          int Nx, int Ny, int Nz, int T) {
    launch_cells(Nx, Ny, Nz, T, [&] { ComputeIntermediateVariablesNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); });
}

void ref_qint_bdry_nodiff(const float* Q, float* Qint, float dt, float dx, float dy, float dz,
                          int Nx, int Ny, int Nz) {
    // order of no_diffusion.cu:302-311
    for (int i = 0; i < Nx; ++i)
        for (int j = 0; j < Ny; ++j) { set_thread(i, j, 0, Nx, Ny, 1); QintBdryFrontNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); }
    for (int k = 0; k < Nz; ++k)
        for (int i = 0; i < Nx; ++i) { set_thread(i, 0, k, Nx, 1, Nz); QintBdryLeftRightNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); }
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j) { set_thread(0, j, k, 1, Ny, Nz); QintBdryTopBottomNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); }
    for (int j = 0; j < Ny; ++j) { set_thread(0, j, 0, 1, Ny, 1); QintBdryFrontBottomNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); }
    for (int i = 0; i < Nx; ++i) { set_thread(i, 0, 0, Nx, 1, 1); QintBdryFrontRightNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); }
    for (int k = 0; k < Nz; ++k) { set_thread(0, 0, k, 1, 1, Nz); QintBdryBottomRightNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); }
    for (int i = 0; i < Nx; ++i)
        for (int j = 0; j < Ny; ++j) { set_thread(i, j, 0, Nx, Ny, 1); QintBdryPBCs(Q, Qint, Nx, Ny, Nz); }
}

void ref_corrector_nodiff(float* Q, const float* Qint, float dt, float dx, float dy, float dz,
                          int Nx, int Ny, int Nz, int T) {
    launch_cells(Nx, Ny, Nz, T, [&] { FluidAdvanceLocalNoDiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz); });
}

// ---- Path B granular kernels ------------------------------------------------
void ref_predictor_stride(const float* Q, float* Qint, float D, float dt, float dx, float dy,
                          float dz, int Nx, int Ny, int Nz, int T) {
    launch_stride(T, [&] { ComputeIntermediateVariablesStride(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz); });
}

// full != 0: every (i,j,k) thread runs the mega-kernel, exactly as a domain-covering launch
// would (each repeats the face work; slow).  full == 0: only the threads whose writes are not
// duplicates of another thread's -- (i,j,0), (i,0,k), (0,j,k) -- same result (tested).
void ref_qint_boundary(const float* Q, float* Qint, float D, float dt, float dx, float dy,
                       float dz, int Nx, int Ny, int Nz, int full) {
    auto run = [&](int i, int j, int k) {
        set_thread(i, j, k, Nx, Ny, Nz);
        ComputeIntermediateVariablesBoundary(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
    };
    if (full) {
        for (int k = 0; k < Nz; ++k)
            for (int i = 0; i < Nx; ++i)
                for (int j = 0; j < Ny; ++j) run(i, j, k);
        return;
    }
    for (int k = 1; k < Nz; ++k) {
        for (int i = 0; i < Nx; ++i) run(i, 0, k);
        for (int j = 1; j < Ny; ++j) run(0, j, k);
    }
    for (int i = 0; i < Nx; ++i)
        for (int j = 0; j < Ny; ++j) run(i, j, 0);
}

void ref_corrector_diff(float* Q, const float* Qint, float D, float dt, float dx, float dy,
                        float dz, int Nx, int Ny, int Nz, int T) {
    launch_stride(T, [&] { FluidAdvanceLocal(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz); });
}

void ref_boundary_conditions(float* Q, const float* Qint, float D, float dt, float dx, float dy,
                             float dz, int Nx, int Ny, int Nz) {
    // z-extent 1
    for (int i = 0; i < Nx; ++i)
        for (int j = 0; j < Ny; ++j) { set_thread(i, j, 0, Nx, Ny, 1); BoundaryConditions(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz); }
}

// ---- composite drivers ---------------------------------------------------------
void ref_pathA_prime(float* Q, float* Qint, float dt, float dx, float dy, float dz,
                     int Nx, int Ny, int Nz, int T) {
    // no_diffusion.cu:174-199
    ref_wall_bcs_leftright(Q, Nx, Ny, Nz);
    ref_wall_bcs_topbottom(Q, Nx, Ny, Nz);
    ref_pbcs(Q, Nx, Ny, Nz);
    ref_predictor_nodiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz, T);
    ref_qint_bdry_nodiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
}

void ref_pathA_steps(float* Q, float* Qint, int nsteps, float dt, float dx, float dy, float dz,
                     int Nx, int Ny, int Nz, int T) {
    for (int s = 0; s < nsteps; ++s) {  // no_diffusion.cu:288-311
        ref_corrector_nodiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz, T);
        ref_pbcs(Q, Nx, Ny, Nz);
        ref_predictor_nodiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz, T);
        ref_qint_bdry_nodiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
    }
}

void ref_pathB_prime(float* Q, float* Qint, float D, float dt, float dx, float dy, float dz,
                     int Nx, int Ny, int Nz, int T) {
    // main.cu:108-112
    ref_predictor_stride(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz, T);
    ref_qint_boundary(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz, 0);
}

void ref_pathB_steps(float* Q, float* Qint, int nsteps, float D, float dt, float dx, float dy,
                     float dz, int Nx, int Ny, int Nz, int T) {
    for (int s = 0; s < nsteps; ++s) {  // main.cu:200-213
        ref_corrector_diff(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz, T);
        ref_boundary_conditions(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
        ref_predictor_stride(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz, T);
        ref_qint_boundary(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz, 0);
    }
}

}  // extern "C"
