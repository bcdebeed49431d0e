// Internal interface of the multi-GPU z-slab engine (imhd_slabs.cu) used by the context functions of imhd_api.cu.
// Not part of the public ABI (include/imhd_b200.h declares what callers see).
#pragma once
#include "../../include/imhd_b200.h"

namespace imhd {

struct Engine;

// `nlocal` slabs of a domain cut into `world` slabs live in this process: slab ranks[q] on device devices[q].
// uid = the 128-byte NCCL unique id shared by all processes (multi-process mode), or nullptr when every slab of the
// domain is local (single process: ncclCommInitAll).
Engine* eng_create(int Nx, int Ny, int Nz, int world, int nlocal, const int* ranks, const int* devices, const void* uid);
void eng_destroy(Engine* e);
int eng_nlocal(const Engine* e);
int eng_local_extent(const Engine* e, int q, int* k0, int* nzl, int* device);
int eng_init_grids(Engine* e, const float bounds[6]);
int eng_init_ic(Engine* e, int ic, float a, float b);
int eng_set_state(Engine* e, const float* host_full);                 // (8,Nz,Nx,Ny): every local slab takes its planes
int eng_get_state(Engine* e, float* host_full);                       // ... and writes them back (other planes untouched)
int eng_set_state_local(Engine* e, int q, const float* host_slab);    // (8,nzl,Nx,Ny) owned planes of local slab q
int eng_get_state_local(Engine* e, int q, float* host_slab);
int eng_refresh_ghosts(Engine* e);                                    // after set_state_local on every slab
int eng_set_spacing(Engine* e, float dx, float dy, float dz);
int eng_prime(Engine* e, int path, float D, float dt);
int eng_step(Engine* e, int nsteps);
int eng_synchronize(Engine* e);
int eng_stability(Engine* e, float dt, imhd_stability* out);
float* eng_device_state(Engine* e, int q);                            // ghosted (8,nzl+2,Nx,Ny) array of local slab q
void* eng_stream(Engine* e, int q);
void eng_set_edge(int planes);                                        // planes launched ahead at each slab end (tuning)

}  // namespace imhd

// internal entry points of imhd_granular.cu used by the engine
int imhd_init_ic_slab(int ic, float* Q, float a, float b, const float* x, const float* y, const float* z, int Nx, int Ny,
                      int nz_array, int kofs, void* stream);
int imhd_init_axis_slab(float* g, float lo, float d, int n, int ofs, void* stream);
int imhd_wall_leftright_planes(float* Q, int Nx, int Ny, int nz_array, int ka, int kb, void* stream);
// internal entry point of imhd_fused.cu: imhd_qint_plane with blocks of block_rows x 32 threads
int imhd_qint_plane_rows(const float* Q, float* out_plane, int k, const imhd_slab* s, int block_rows, void* stream);
// internal entry points of imhd_stability.cu: the scan without its stream synchronisation, and the decoder of its 16 bytes
int imhd_stability_scan_async(const float* Q, const imhd_slab* s, unsigned long long h[2], void* stream);
void imhd_stability_decode(const unsigned long long h[2], const imhd_slab* s, imhd_stability* host_out);
