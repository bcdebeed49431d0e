// Shared host/device plumbing for libimhd_b200: error handling, launch counting, the
// per-launch parameter block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/imhd_b200.h"
#include "imhd_math.cuh"

namespace imhd {

// Per-launch constants.  tx = dt/dx etc. are fp32 divisions, as the reference evaluates
// `(dt / dx)` (kernels_od_intvar.cu:1166-1168).
struct Params {
    int Nx, Ny, Nz;
    int path;
    long long plane;  // Nx*Ny
    long long cube;   // Nx*Ny*Nz (variable stride of a full-domain array)
    float D, dt, dx, dy, dz;
    float tx, ty, tz;
    DiffCoef dc;
};

inline Params make_params(int path, float D, float dt, float dx, float dy, float dz, int Nx, int Ny, int Nz) {
    Params p;
    p.Nx = Nx; p.Ny = Ny; p.Nz = Nz; p.path = path;
    p.plane = (long long)Nx * Ny;
    p.cube = p.plane * Nz;
    p.D = D; p.dt = dt; p.dx = dx; p.dy = dy; p.dz = dz;
    p.tx = dt / dx; p.ty = dt / dy; p.tz = dt / dz;
    p.dc.cx = 1.0 / ((double)dx * (double)dx);
    p.dc.cy = 1.0 / ((double)dy * (double)dy);
    p.dc.cz = 1.0 / ((double)dz * (double)dz);
    p.dc.cxf = (float)p.dc.cx; p.dc.cyf = (float)p.dc.cy; p.dc.czf = (float)p.dc.cz;
    p.dc.c0f = (float)(-2.0 * (p.dc.cx + p.dc.cy + p.dc.cz));
    p.dc.dtD = dt * D;
    return p;
}

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(unsigned n = 1);

#define IMHD_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) return imhd::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define IMHD_LAUNCH_CHECK(n)                                                                 \
    do {                                                                                     \
        imhd::count_launch(n);                                                               \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) return imhd::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

inline int bad_dims(int Nx, int Ny, int Nz) {
    if (Nx < 4 || Ny < 4 || Nz < 4) {
        set_error("grid %dx%dx%d too small (every axis needs >= 4 points)", Nx, Ny, Nz);
        return IMHD_E_INVALID;
    }
    return 0;
}

}  // namespace imhd
