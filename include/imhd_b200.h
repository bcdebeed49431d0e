/*
 * imhd_b200.h -- C ABI of libimhd_b200.so, the B200-native (sm_100a) replacement for the
 * Lax-Wendroff hot path of russellmatt66/imhd-CUDA.
 *
 * The reference has no C ABI: its "operator API" is the set of C++-mangled __global__ kernels
 * declared in include/on-device/{kernels_od,kernels_fluidbcs,kernels_od_intvar,kernels_intvarbcs,
 * initialize_od}.cuh, launched by src/on-device/main.cu (Path B, with diffusion) and
 * src/on-device/no_diffusion.cu (Path A).  Each entry point below names the reference
 * interface it replaces (file:line, relative to the reference root).  INTEGRATION.md shows the
 * binding a reference maintainer would add.
 *
 * Conventions
 *  - Plain pointers and sizes only.  `stream` is a cudaStream_t passed as void* (NULL = legacy
 *    default stream).  Everything is stream-ordered; nothing synchronises the device unless
 *    stated.  One host thread per context.
 *  - Every function returns 0 on success or a non-zero code (a cudaError_t, or IMHD_E_*);
 *    imhd_last_error() describes the last failure on the calling thread.  Nothing throws or
 *    calls exit() (the reference's checkCuda aborts: include/on-device/utils/utils.cuh:8-16).
 *  - State arrays are fp32 in the reference's IDX3D layout (lib/on-device/kernels_od.cu:11,16):
 *        l = k*Nx*Ny + i*Ny + j,  variable v at l + v*Nx*Ny*Nz,
 *        v = rho, rhovx, rhovy, rhovz, Bx, By, Bz, e          (j is unit stride)
 *    with 64-bit offsets inside (the reference's int offsets overflow above 306 M cells).
 *  - `path`: IMHD_PATH_A = no_diffusion.cu pipeline, IMHD_PATH_B = main.cu pipeline (D used).
 *  - There is NO CPU fallback: every compute entry point launches sm_100a kernels and fails
 *    with the CUDA error if no device is usable.
 */
#ifndef IMHD_B200_H
#define IMHD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMHD_PATH_A 0 /* src/on-device/no_diffusion.cu:284-312 */
#define IMHD_PATH_B 1 /* src/on-device/main.cu:196-213        */

#define IMHD_E_INVALID 10001 /* bad argument                     */
#define IMHD_E_STATE 10002   /* call order / context state error */
#define IMHD_E_IO 10003      /* file output failed               */

/* Bit flags for imhd_set_options(): which arithmetic the fused kernels use. */
#define IMHD_OPT_DEFAULT 0u

typedef struct imhd_ctx imhd_ctx;

/* ---- library ------------------------------------------------------------------------- */
int imhd_abi_version(void);
const char* imhd_last_error(void);
/* Number of kernels this library has launched on the calling process since load (bench.py's
 * gpu_launches claim is read from here, not estimated). */
uint64_t imhd_launch_count(void);

/* ---- parity-granular operators on caller-owned DEVICE buffers (full domain, IDX3D) -------
 * These mirror the reference kernels one group at a time and keep its fp32/fp64 rounding
 * points (compiled without FMA contraction), so they can be compared call by call. */

/* Predictor: every cell of Qint, exactly as the reference's kernel sequence leaves it.
 *  A: ComputeIntermediateVariablesNoDiff (lib/on-device/kernels_od_intvar.cu:51) +
 *     QintBdry{Front,LeftRight,TopBottom,FrontBottom,FrontRight,BottomRight}NoDiff + QintBdryPBCs
 *     (lib/on-device/kernels_intvarbcs.cu:360-558)
 *  B: ComputeIntermediateVariablesStride (kernels_od_intvar.cu:113) +
 *     ComputeIntermediateVariablesBoundary (kernels_intvarbcs.cu:177) */
int imhd_predictor(const float* Q, float* Qint, int path, float D, float dt, float dx, float dy,
                   float dz, int Nx, int Ny, int Nz, void* stream);

/* Corrector over the volume, in place.
 *  A: FluidAdvanceLocalNoDiff (lib/on-device/kernels_od.cu:353)   B: FluidAdvanceLocal (:82) */
int imhd_corrector(float* Q, const float* Qint, int path, float D, float dt, float dx, float dy,
                   float dz, int Nx, int Ny, int Nz, void* stream);

/* Fluid boundary pass that follows the corrector each step.
 *  A: PBCs (lib/on-device/kernels_fluidbcs.cu:498)
 *  B: BoundaryConditions (kernels_fluidbcs.cu:32), single-application semantics */
int imhd_fluid_bcs(float* Q, const float* Qint, int path, float D, float dt, float dx, float dy,
                   float dz, int Nx, int Ny, int Nz, void* stream);

/* Path A initial boundary pass: rigidConductingWallBCsLeftRight + (no-op) ...TopBottom + PBCs
 * (kernels_fluidbcs.cu:436,467,498; call site no_diffusion.cu:174-177). */
int imhd_initial_bcs(float* Q, int Nx, int Ny, int Nz, void* stream);

/* ---- grids and initial conditions (lib/on-device/initialize_od.cu) ------------------------ */
/* InitializeX/Y/Z (:26-57) with dx = (x_max-x_min)/(Nx-1) in fp32 (main.cu:98-100). */
int imhd_init_grids(float* x, float* y, float* z, float x_min, float x_max, float y_min,
                    float y_max, float z_min, float z_max, int Nx, int Ny, int Nz, void* stream);
/* ScrewPinchStride (:269-345) */
int imhd_init_screwpinch_stride(float* Q, float J0, const float* x, const float* y, const float* z,
                                int Nx, int Ny, int Nz, void* stream);
/* CubicBennettVortex_m0 (:132-205); `k` is accepted and, as in the reference, shadowed by the
 * z loop index. */
int imhd_init_cubic_bennett_vortex_m0(float* Q, float k, float A, const float* x, const float* y,
                                      const float* z, int Nx, int Ny, int Nz, void* stream);
/* The three initial conditions the shipped drivers keep commented out (no_diffusion.cu:169-171):
 * CubicBennettVortex (:59-130), ZPinch (:347-424), ScrewPinch (:207-267; outside the pinch it writes only
 * rho = 0.1 and leaves the other seven variables as it found them). */
int imhd_init_cubic_bennett_vortex(float* Q, const float* x, const float* y, const float* z, int Nx, int Ny,
                                   int Nz, void* stream);
int imhd_init_zpinch(float* Q, float r_max_coeff, const float* x, const float* y, const float* z, int Nx,
                     int Ny, int Nz, void* stream);
int imhd_init_screwpinch(float* Q, float J0, float r_max_coeff, const float* x, const float* y,
                         const float* z, int Nx, int Ny, int Nz, void* stream);

/* ---- fused time step: the product hot path ---------------------------------------------------
 * One sweep Q^n -> Q^{n+1}: predictor, corrector, diffusion and every boundary pass of one
 * reference time step (main.cu:200-213 / no_diffusion.cu:288-311), with Qint living only
 * on chip.  Qin and Qout are distinct device buffers (ping-pong); the reference's `intvars`
 * allocation becomes the second buffer.
 *
 * Slab form (z-slab domain decomposition, one slab per GPU): the buffers hold global planes
 * [k0-1, k0+nzl] -- nzl owned planes plus one ghost plane on each side -- as a
 * (8, nzl+2, Nx, Ny) array; pass k0 = 0, nzl = Nz and ghosts = 0 for a plain (8,Nz,Nx,Ny)
 * full-domain array.  The (8,Nx,Ny) predictor planes at the slab ends are inputs, produced by
 * imhd_qint_plane on the neighbouring rank (or on this one where the slab touches a domain end;
 * Qint is periodic in z with period Nz-1):
 *   qint_lo   = Qint(k0-1), or this slab's own Qint(0) when k0 == 0
 *   qint_hi   = Qint(k0+nzl), or Qint(0) (== Qint(Nz-1)) when the slab ends the domain
 *   qint_wrap = Qint(Nz-2) (== Qint(-1)); only read for path B on the slab with k0 == 0 (k=0 face)
 * On one GPU: qint_lo = qint_hi = Qint(0), qint_wrap = Qint(Nz-2). */
typedef struct {
    int Nx, Ny, Nz;     /* global grid                                        */
    int k0, nzl;        /* first owned global plane, number of owned planes   */
    int ghosts;         /* 0: no ghost planes in the arrays, 1: one each side */
    int path;           /* IMHD_PATH_A / IMHD_PATH_B                           */
    float D, dt, dx, dy, dz;
    float corner_e;     /* path B only: imhd_wall_energy_fixed_point(e of Q(Nx-1,Ny-1,0) at t=0);
                           the value BoundaryConditions leaves in the column (Nx-1,Ny-1) at k=0
                           and k=Nz-1 (lib/on-device/kernels_fluidbcs.cu:178-188,227-231) */
} imhd_slab;

int imhd_step_fused(const float* Qin, float* Qout, const float* qint_lo, const float* qint_hi,
                    const float* qint_wrap, const imhd_slab* s, void* stream);

/* The same step restricted to the output planes [kfrom, kto) of the slab (global indices).  Any split of
 * the owned range writes the same bits as one imhd_step_fused call: the multi-GPU loop computes the
 * planes next to the slab ends first, so that their halo exchange overlaps the interior launch. */
int imhd_step_fused_planes(const float* Qin, float* Qout, const float* qint_lo, const float* qint_hi,
                           const float* qint_wrap, const imhd_slab* s, int kfrom, int kto, void* stream);

/* ... restricted to TWO plane ranges, [kfrom, kmid1) and [kmid2, kto): both slab ends at once.  ONE launch of the
 * marching kernel when the two ranges hold the same number (<= 32) of its planes (plane 0, and plane Nz-1 of the
 * diffusion pipeline, have kernels of their own and do not count), otherwise two launches; the same bits either way. */
int imhd_step_fused_ends(const float* Qin, float* Qout, const float* qint_lo, const float* qint_hi,
                         const float* qint_wrap, const imhd_slab* s, int kfrom, int kmid1, int kmid2, int kto, void* stream);

/* ---- CFL / stability scan (replaces the forked host scanner src/on-device/utils/compute_stability.cpp) ----------
 * One device pass over the owned planes of a slab state (same Q / imhd_slab conventions as imhd_step_fused; dt, dx,
 * dy, dz from the slab): LHS = (dt/dx)|l_x| + (dt/dy)|l_y| + (dt/dz)|l_z| per cell (:157-163) with |l_d| the
 * spectral radius of the ideal-MHD flux Jacobian in closed form (|u_d| + fast magnetosonic speed for a physical
 * state) instead of Eigen on three 8x8 matrices (:165-181).  Synchronises the stream and fills host_out.
 * Across slabs: max over max_lhs (ties: smallest k), sum over violations (imhd-cuda_b200/slab.py). */
typedef struct imhd_stability {
    float max_lhs;                 /* largest LHS among the scanned cells (0 if none is positive and finite) */
    int i, j, k;                   /* its cell, global indices; first in the reference's scan order (k, i, j) */
    unsigned long long violations; /* cells with LHS >= 1 (:120-121) */
    float dt_new;                  /* 0.1 * dt / max_lhs: the reference's proposal (:139-141); 0 if max_lhs == 0 */
} imhd_stability;
int imhd_stability_scan(const float* Q, const imhd_slab* s, imhd_stability* host_out, void* stream);
/* Which spectra the scan uses (process-wide; default IMHD_STABILITY_WAVE_SPEEDS):
 *   IMHD_STABILITY_WAVE_SPEEDS       the exact ideal-MHD wave speeds in x, y and z (closed form);
 *   IMHD_STABILITY_REFERENCE_QUIRKS  the reference's own report: closed form for its x matrix (which has exactly the wave
 *                                    spectrum) and the spectral radius of ITS y and z matrices, transcription slips
 *                                    included (computeB / computeC, compute_stability.cpp:296-459), eigenvalues by
 *                                    Hessenberg reduction + shifted QR per cell on the device in place of Eigen (:165-181). */
#define IMHD_STABILITY_WAVE_SPEEDS 0
#define IMHD_STABILITY_REFERENCE_QUIRKS 1
void imhd_stability_mode(int mode);

/* e <- p(e,0,0)/(gamma-1) iterated to its fixed point (lib/on-device/kernels_fluidbcs.cu:173,187; every
 * x-thread of the reference launch re-applies it).  Scalar host helper, identity when e == 0. */
float imhd_wall_energy_fixed_point(float e, int max_iter);

/* Test hooks: force the z-chunk length of the fused kernel (0 = automatic); kernel-variant flags: bit 0 forces the
 * plain-load variant instead of the TMA one, bit 1 disables the remainder-strip kernel, bit 2 uses it even for
 * plane ranges shorter than 64, bit 3 keeps the strip and the two z faces of path B on the caller's stream instead of
 * running them under the marching kernel on the library's side stream, bit 8 selects the block-per-tile strip kernel instead of the warp-autonomous one, bits 4..7
 * select the marching kernel (0 default, 1 one row per thread, 2 two rows per thread behind a block-wide barrier per plane,
 * 4 two rows per thread with the split-phase exchange barrier on both pipelines).  All give the same bits. */
void imhd_set_chunk(int planes);
void imhd_set_kernel_variant(int flags);

/* Measurement hook (bench.py's roofline leg): while enabled, every launch of the marching kernel over >= 64 planes
 * (with its remainder strip) is bracketed by a CUDA event pair on the launching stream; after synchronising, read the
 * summed duration, the number of launches and the cell-updates they covered.  imhd_fused_timing(0|1) also resets. */
void imhd_fused_timing(int enable);
int imhd_fused_timing_read(double* total_ms, int* launches, long long* cell_updates);

/* Predictor plane Qint(.,.,k) for one owned global plane k into an (8,Nx,Ny) device buffer
 * (the data a neighbouring slab needs as qint_lo / qint_hi). */
int imhd_qint_plane(const float* Q, float* out_plane, int k, const imhd_slab* s, void* stream);

/* ---- context API: what the drop-in drivers (imhd-cuda, imhd-cuda_nodiff) call ------------------
 * Owns the ping-pong state buffers, grids and staging for one GPU (one slab). */
imhd_ctx* imhd_create(int Nx, int Ny, int Nz, int device);
void imhd_destroy(imhd_ctx* ctx);
/* x_min..z_max as argv 8-13 (A) / 7-12 (B) of the reference drivers. */
int imhd_ctx_init_grids(imhd_ctx* ctx, float x_min, float x_max, float y_min, float y_max,
                        float z_min, float z_max);
int imhd_ctx_init_screwpinch_stride(imhd_ctx* ctx, float J0);
int imhd_ctx_init_cubic_bennett_vortex_m0(imhd_ctx* ctx, float k, float A);
/* ---- string-keyed registry: the reference's planned plugin surface
 * (include/on-device/utils/configurers.hpp:13-196; src/on-device/README.md:10-30).  Keys of the reference:
 *   initializers   "screwpinch" {J0, r_max_coeff}, "screwpinch-stride" {J0}, "cubic-bennett-vortex" {}
 *   correctors     "fluidadvancelocal-nodiff"
 *   predictors     "corrector_advance-tp_nodiff", "corrector_advance-stride_nodiff"
 *   fluid BCs      "pcrw-xy_pbc-z"          predictor BCs  "pbc-z"
 * plus, in the slots it leaves open: "cubic-bennett-vortex-m0" {k, A}, "zpinch" {r_max_coeff},
 * "fluidadvancelocal" / "corrector_advance-stride" (the with-diffusion loop of main.cu).
 * Unknown keys fail with the reference's messages ("Unknown simulation type: ...") in imhd_last_error(). */
#define IMHD_REG_INITIALIZER 0
#define IMHD_REG_CORRECTOR 1
#define IMHD_REG_PREDICTOR 2
#define IMHD_REG_FLUID_BCS 3
#define IMHD_REG_PREDICTOR_BCS 4
int imhd_registry_count(int kind);
const char* imhd_registry_name(int kind, int index);
int imhd_registry_initializer_nparams(const char* sim_type); /* -1: unknown key */
/* SimulationInitializer::initialize (configurers.hpp:32-39): run the named IC kernel on the context's state. */
int imhd_ctx_initialize(imhd_ctx* ctx, const char* sim_type, const float* params, int nparams);
/* Resolve a (corrector, predictor, fluid BCs, predictor BCs) selection to IMHD_PATH_A / IMHD_PATH_B for
 * imhd_ctx_prime; bundles from different time loops do not mix. */
int imhd_registry_resolve_path(const char* corrector, const char* predictor, const char* fluid_bcs,
                               const char* predictor_bcs, int* path);
/* Upload a host state (8*Nx*Ny*Nz floats, IDX3D) instead of running an IC kernel. */
int imhd_ctx_set_state(imhd_ctx* ctx, const float* host_Q);
/* Grid spacing for a state uploaded with imhd_ctx_set_state (imhd_ctx_init_grids sets it too). */
int imhd_ctx_set_spacing(imhd_ctx* ctx, float dx, float dy, float dz);
/* Everything the reference drivers do between the IC kernel and the time loop
 * (no_diffusion.cu:174-199 / main.cu:108-112). */
int imhd_ctx_prime(imhd_ctx* ctx, int path, float D, float dt);
/* nsteps iterations of the time loop, fused kernels, no host sync. */
int imhd_ctx_step(imhd_ctx* ctx, int nsteps);
/* Same, through the parity-granular operators (4 reference-shaped passes per step). */
int imhd_ctx_step_granular(imhd_ctx* ctx, int nsteps);
/* Copy the current state / intermediate state / grids to host memory (synchronises). */
int imhd_ctx_get_state(imhd_ctx* ctx, float* host_Q);
int imhd_ctx_get_grids(imhd_ctx* ctx, float* x, float* y, float* z);
/* imhd_stability_scan of the context's current state with its spacing and the given dt. */
int imhd_ctx_stability(imhd_ctx* ctx, float dt, imhd_stability* host_out);
/* Adaptive time step on top of the scan (the reference's README keeps "adaptive dt" as a TODO; its scanner
 * src/on-device/utils/compute_stability.cpp runs once, as a forked host program).  nsteps fused steps; before steps 0,
 * every, 2*every, ... the current state is scanned on the context's stream WITHOUT stalling the loop, and the scan taken
 * before step g*every sets the dt of steps [(g+1)*every, (g+2)*every):
 *     dt = min(dt_max, cfl_target * dt_scan / max_lhs)      (LHS is linear in dt; dt_scan = the dt in use at the scan)
 * -- one group of lag, so that the host only waits for results the device produced a whole group of steps earlier.  The
 * first group runs with the context's dt; the context keeps the last dt.  dt_used (may be NULL) receives the dt of every
 * step.  every >= 2.  Single-GPU contexts. */
int imhd_ctx_step_adaptive(imhd_ctx* ctx, int nsteps, int every, float cfl_target, float dt_max, float* dt_used);
/* Replace the time step set by imhd_ctx_prime (nothing else of the priming depends on it).  Single-GPU contexts. */
int imhd_ctx_set_dt(imhd_ctx* ctx, float dt);
/* Device pointer of the current state (valid until the next step call). */
float* imhd_ctx_device_state(imhd_ctx* ctx);
void* imhd_ctx_stream(imhd_ctx* ctx);
int imhd_ctx_synchronize(imhd_ctx* ctx);

/* ---- output: the reference's on-disk contract without libhdf5 -----------------
 * fluidvars_<it>.h5: 8 one-dimensional fp32 datasets rho, rhovx, rhovy, rhovz, Bx, By, Bz, e of Nx*Ny*Nz
 * elements in IDX3D order (src/on-device/utils/phdf5_write_all.cpp:87-169); with_attributes adds
 * cubeDimensions / cubeDimensionsNames / storagePattern (hdf5_write_attributes.cpp:55-98, frame 0 only).
 * grid.h5: x_grid, y_grid, z_grid with scalar attributes spacing, dimension (hdf5_write_grid.cpp:78-137). */
int imhd_h5_write_fluidvars(const char* path, const float* host_Q, int Nx, int Ny, int Nz, int with_attributes);
int imhd_h5_write_grid(const char* path, const float* x, const float* y, const float* z, int Nx, int Ny, int Nz);
/* Queue the current state as <dir>fluidvars_<frame>.h5 (dir ends with '/'); returns at once -- device
 * snapshot, D2H on a copy stream, file write on a writer thread (replaces main.cu:216-226).  At most
 * two frames are in flight. */
int imhd_ctx_write_frame(imhd_ctx* ctx, const char* dir, int frame);
int imhd_ctx_flush_output(imhd_ctx* ctx);
int imhd_ctx_write_grid(imhd_ctx* ctx, const char* dir);

/* ---- multi-GPU: the z-slab time loop behind the same context functions -----------------------------
 * The reference is single-GPU (its host loop: src/on-device/main.cu:196-238 / no_diffusion.cu:284-337); these two
 * constructors give that loop a domain cut into z-slabs, one per GPU, ghost planes and predictor planes exchanged
 * between ring neighbours over NVLink (ncclSend/ncclRecv on a side stream, under the interior launch).  Every
 * imhd_ctx_* function above works on such a context and means the WHOLE domain (init_grids, the initial conditions,
 * set_state / get_state with full (8,Nz,Nx,Ny) host arrays, prime, step, stability, synchronize, imhd_run_host);
 * results are bit-identical to the single-GPU context.  imhd_ctx_write_frame gathers the slabs and writes
 * synchronously there; imhd_ctx_step_granular and imhd_ctx_get_grids are single-GPU only.  NCCL is bound at run time
 * (dlopen libnccl.so.2); single-GPU use does not need it.
 *   imhd_create_multi : all n_gpus slabs in THIS process, slab q on devices[q] (NULL = devices 0..n_gpus-1)
 *   imhd_create_slab  : slab `rank` of `world` in this process (one process per GPU, e.g. under torchrun);
 *                       nccl_unique_id = the 128 bytes one process obtained from imhd_nccl_unique_id. */
imhd_ctx* imhd_create_multi(int Nx, int Ny, int Nz, int n_gpus, const int* devices);
imhd_ctx* imhd_create_slab(int Nx, int Ny, int Nz, int rank, int world, int device, const void* nccl_unique_id);
int imhd_nccl_unique_id(void* id_out, int bytes); /* bytes >= 128 */
/* Slabs held by this context (1 for a single-GPU or imhd_create_slab context) and their owned global planes
 * [k0, k0+nzl) / device. */
int imhd_ctx_num_slabs(imhd_ctx* ctx);
int imhd_ctx_slab_extent(imhd_ctx* ctx, int q, int* k0, int* nzl, int* device);
/* Slab-local transfers: the owned planes (8, nzl, Nx, Ny) of local slab q from / to host memory (a process that holds
 * one slab never needs the full array).  Ghost planes are refreshed by the next imhd_ctx_prime. */
int imhd_ctx_set_state_local(imhd_ctx* ctx, int q, const float* host_slab);
int imhd_ctx_get_state_local(imhd_ctx* ctx, int q, float* host_slab);
/* Tuning hook: planes next to each slab end that are launched ahead of the interior (>= 3; default 4). */
void imhd_set_edge_planes(int planes);

/* Whole job through HOST buffers (the e2e path bench.py times): upload host_Q_in, prime,
 * run nsteps fused steps, download into host_Q_out.  Both buffers 8*Nx*Ny*Nz floats. */
int imhd_run_host(imhd_ctx* ctx, const float* host_Q_in, float* host_Q_out, int path, float D,
                  float dt, float dx, float dy, float dz, int nsteps);

#ifdef __cplusplus
}
#endif
#endif /* IMHD_B200_H */
