// libimhd_b200: library plumbing (errors, launch counter) and the context API the drop-in
// drivers call.  Host time loop replaces src/on-device/main.cu:196-238 and
// src/on-device/no_diffusion.cu:284-337: no per-phase cudaDeviceSynchronize, no per-step
// blocking D2H, one fused kernel per step on a single stream.
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "imhd_common.cuh"

namespace imhd {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    // same information as the reference's checkCuda (include/on-device/utils/utils.cuh:8-16),
    // but returned instead of exit()
    set_error("GPUassert: %s (%s) %s %d", cudaGetErrorString(e), what, file, line);
    return (int)e;
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace imhd

using namespace imhd;

#include "imhd_engine.h"

struct imhd_ctx {
    imhd::Engine* eng;         // multi-GPU contexts (imhd_create_multi / imhd_create_slab): the z-slab engine owns everything
    int Nx, Ny, Nz, device;
    size_t cells;
    float* buf[2];  // ping-pong state buffers; buf[cur] holds Q^n.  buf[1-cur] is the reference's `intvars`
                    // allocation: Qint for the granular path, Q^{n+1} for the fused path.
    int cur;
    float *gx, *gy, *gz;       // device grids
    float* qint_planes;        // 2 x (8,Nx,Ny): predictor planes for the periodic wrap
    cudaStream_t stream;
    bool have_grids, primed, qint_valid;
    int path;
    float D, dt, dx, dy, dz, corner_e;
    float bounds[6];
    // ---- asynchronous output (replaces the per-step blocking cudaMemcpy + fork of main.cu:216-226) ----
    float* snap;             // device snapshot of the frame being written out
    float* pinned[2];        // pinned host staging, double buffered
    cudaStream_t copy_stream;
    cudaEvent_t snap_done, snap_free, copy_done[2];
    bool snap_used;          // a D2H copy out of `snap` has been queued: the next snapshot waits for snap_free
    int next_slot;
    struct Job { int slot, frame, attrs; std::string path; };
    std::deque<Job>* jobs;
    std::mutex* mu;
    std::condition_variable* cv;
    std::thread* writer;
    bool slot_busy[2];
    bool stop;
    int write_errors;
    char write_error_text[256];  // the writer thread's last error (its set_error is thread-local)
    // ---- adaptive time step (imhd_ctx_step_adaptive): two scans in flight at most ----
    unsigned long long* scan_pinned;   // 2 x 2 words of pinned host memory
    cudaEvent_t scan_done[2];
};

extern "C" int imhd_abi_version(void) { return 1; }
extern "C" const char* imhd_last_error(void) { return g_err; }
extern "C" uint64_t imhd_launch_count(void) { return g_launches.load(); }

extern "C" imhd_ctx* imhd_create(int Nx, int Ny, int Nz, int device) {
    if (bad_dims(Nx, Ny, Nz)) return nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("imhd_create: no usable CUDA device (%s); this library has no CPU path",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) { set_error("imhd_create: device %d out of range [0,%d)", device, ndev); return nullptr; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__); return nullptr; }
    imhd_ctx* c = new (std::nothrow) imhd_ctx();
    if (!c) { set_error("imhd_create: out of host memory"); return nullptr; }
    memset(c, 0, sizeof(*c));
    c->Nx = Nx; c->Ny = Ny; c->Nz = Nz; c->device = device;
    c->cells = (size_t)Nx * Ny * Nz;
    const size_t bytes = 8 * c->cells * sizeof(float);
    bool ok = cudaMalloc(&c->buf[0], bytes) == cudaSuccess && cudaMalloc(&c->buf[1], bytes) == cudaSuccess &&
              cudaMalloc(&c->gx, sizeof(float) * Nx) == cudaSuccess && cudaMalloc(&c->gy, sizeof(float) * Ny) == cudaSuccess &&
              cudaMalloc(&c->gz, sizeof(float) * Nz) == cudaSuccess &&
              cudaMalloc(&c->qint_planes, 2 * 8 * sizeof(float) * (size_t)Nx * Ny) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        cuda_fail(cudaGetLastError(), "imhd_create allocation", __FILE__, __LINE__);
        imhd_destroy(c);
        return nullptr;
    }
    return c;
}

static void free_writer_buffers(imhd_ctx* c);

static imhd_ctx* wrap_engine(Engine* e, int Nx, int Ny, int Nz, int device) {
    if (!e) return nullptr;
    imhd_ctx* c = new (std::nothrow) imhd_ctx();
    if (!c) { eng_destroy(e); set_error("out of host memory"); return nullptr; }
    memset(c, 0, sizeof(*c));
    c->eng = e; c->Nx = Nx; c->Ny = Ny; c->Nz = Nz; c->device = device;
    c->cells = (size_t)Nx * Ny * Nz;
    return c;
}

// The whole domain as n_gpus z-slabs driven from THIS process (one slab per device, ring neighbours over NVLink).
// n_gpus == 1 is the plain single-GPU context.
extern "C" imhd_ctx* imhd_create_multi(int Nx, int Ny, int Nz, int n_gpus, const int* devices) {
    if (n_gpus < 1) { set_error("imhd_create_multi: n_gpus = %d", n_gpus); return nullptr; }
    if (n_gpus == 1) return imhd_create(Nx, Ny, Nz, devices ? devices[0] : 0);
    std::vector<int> ranks(n_gpus), devs(n_gpus);
    for (int q = 0; q < n_gpus; ++q) { ranks[q] = q; devs[q] = devices ? devices[q] : q; }
    return wrap_engine(eng_create(Nx, Ny, Nz, n_gpus, n_gpus, ranks.data(), devs.data(), nullptr), Nx, Ny, Nz, devs[0]);
}

// Slab `rank` of `world` in THIS process (one process per GPU); nccl_unique_id = the 128 bytes of imhd_nccl_unique_id
// made by one process and handed to all of them.
extern "C" imhd_ctx* imhd_create_slab(int Nx, int Ny, int Nz, int rank, int world, int device, const void* nccl_unique_id) {
    if (world == 1) return imhd_create(Nx, Ny, Nz, device);
    if (!nccl_unique_id) { set_error("imhd_create_slab: null NCCL unique id"); return nullptr; }
    return wrap_engine(eng_create(Nx, Ny, Nz, world, 1, &rank, &device, nccl_unique_id), Nx, Ny, Nz, device);
}

extern "C" int imhd_ctx_num_slabs(imhd_ctx* c) { return !c ? 0 : (c->eng ? eng_nlocal(c->eng) : 1); }

extern "C" int imhd_ctx_slab_extent(imhd_ctx* c, int q, int* k0, int* nzl, int* device) {
    if (!c) { set_error("null context"); return IMHD_E_INVALID; }
    if (c->eng) return eng_local_extent(c->eng, q, k0, nzl, device);
    if (q != 0) { set_error("no such local slab %d", q); return IMHD_E_INVALID; }
    if (k0) *k0 = 0;
    if (nzl) *nzl = c->Nz;
    if (device) *device = c->device;
    return 0;
}

#define NOT_ON_SLABS(c, what)                                                                          \
    do {                                                                                               \
        if ((c) && (c)->eng) { set_error(what " is not available on a multi-slab context"); return IMHD_E_STATE; } \
    } while (0)

extern "C" void imhd_destroy(imhd_ctx* c) {
    if (!c) return;
    if (c->eng) { eng_destroy(c->eng); delete c; return; }
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    cudaFree(c->buf[0]); cudaFree(c->buf[1]);
    cudaFree(c->gx); cudaFree(c->gy); cudaFree(c->gz); cudaFree(c->qint_planes);
    if (c->scan_pinned) { cudaFreeHost(c->scan_pinned); cudaEventDestroy(c->scan_done[0]); cudaEventDestroy(c->scan_done[1]); }
    if (c->writer) {
        { std::lock_guard<std::mutex> g(*c->mu); c->stop = true; }
        c->cv->notify_all();
        c->writer->join();
        delete c->writer; delete c->jobs; delete c->mu; delete c->cv;
        free_writer_buffers(c);
    }
    delete c;
}

#define CTX_CHECK(c)                                                  \
    do {                                                              \
        if (!(c)) { set_error("null context"); return IMHD_E_INVALID; } \
        IMHD_CUDA(cudaSetDevice((c)->device));                        \
    } while (0)

extern "C" int imhd_ctx_init_grids(imhd_ctx* c, float x_min, float x_max, float y_min, float y_max, float z_min,
                                   float z_max) {
    CTX_CHECK(c);
    const float b[6] = {x_min, x_max, y_min, y_max, z_min, z_max};
    memcpy(c->bounds, b, sizeof(b));
    if (c->eng) {
        c->dx = (x_max - x_min) / (c->Nx - 1); c->dy = (y_max - y_min) / (c->Ny - 1); c->dz = (z_max - z_min) / (c->Nz - 1);
        c->have_grids = true;
        return eng_init_grids(c->eng, b);
    }
    // fp32, as main.cu:98-100 / no_diffusion.cu:106-108
    c->dx = (x_max - x_min) / (c->Nx - 1);
    c->dy = (y_max - y_min) / (c->Ny - 1);
    c->dz = (z_max - z_min) / (c->Nz - 1);
    if (int e = imhd_init_grids(c->gx, c->gy, c->gz, x_min, x_max, y_min, y_max, z_min, z_max, c->Nx, c->Ny, c->Nz, c->stream))
        return e;
    c->have_grids = true;
    return 0;
}

static int need_grids(imhd_ctx* c) {
    if (!c->have_grids) { set_error("initial condition requested before imhd_ctx_init_grids"); return IMHD_E_STATE; }
    return 0;
}

extern "C" int imhd_ctx_init_screwpinch_stride(imhd_ctx* c, float J0) {
    CTX_CHECK(c);
    if (int e = need_grids(c)) return e;
    if (c->eng) return eng_init_ic(c->eng, 0, J0, 0.f);
    c->primed = false;
    return imhd_init_screwpinch_stride(c->buf[c->cur], J0, c->gx, c->gy, c->gz, c->Nx, c->Ny, c->Nz, c->stream);
}

extern "C" int imhd_ctx_init_cubic_bennett_vortex_m0(imhd_ctx* c, float k, float A) {
    CTX_CHECK(c);
    if (int e = need_grids(c)) return e;
    if (c->eng) return eng_init_ic(c->eng, 1, k, A);
    c->primed = false;
    return imhd_init_cubic_bennett_vortex_m0(c->buf[c->cur], k, A, c->gx, c->gy, c->gz, c->Nx, c->Ny, c->Nz, c->stream);
}

// ---- string-keyed registry: the reference's planned plugin surface (include/on-device/utils/configurers.hpp) ----------
// SimulationInitializer (:13-41), FluidKernelConfigurer (:47-72), PredictorKernelConfigurer (:78-110),
// FluidBoundaryConfigurer (:127-148), PredictorBoundaryConfigurer (:180-196).  The reference's keys are kept; keys it
// does not have yet (its "ADD OTHER INITIALIZERS" / "ADD MORE BUNDLES" slots) are marked ext.
namespace {

struct IcEntry {
    const char* name;
    int nparams;
    int (*launch)(imhd_ctx*, const float*);
    int ic;  // IC id of k_init_state (multi-slab contexts)
};

int ic_screwpinch(imhd_ctx* c, const float* p) {
    // ScrewPinch leaves seven variables untouched outside the pinch (initialize_od.cu:237): give them a defined value
    IMHD_CUDA(cudaMemsetAsync(c->buf[c->cur], 0, 8 * c->cells * sizeof(float), c->stream));
    return imhd_init_screwpinch(c->buf[c->cur], p[0], p[1], c->gx, c->gy, c->gz, c->Nx, c->Ny, c->Nz, c->stream);
}
int ic_screwpinch_stride(imhd_ctx* c, const float* p) {
    return imhd_init_screwpinch_stride(c->buf[c->cur], p[0], c->gx, c->gy, c->gz, c->Nx, c->Ny, c->Nz, c->stream);
}
int ic_bennett(imhd_ctx* c, const float*) {
    return imhd_init_cubic_bennett_vortex(c->buf[c->cur], c->gx, c->gy, c->gz, c->Nx, c->Ny, c->Nz, c->stream);
}
int ic_bennett_m0(imhd_ctx* c, const float* p) {
    return imhd_init_cubic_bennett_vortex_m0(c->buf[c->cur], p[0], p[1], c->gx, c->gy, c->gz, c->Nx, c->Ny, c->Nz, c->stream);
}
int ic_zpinch(imhd_ctx* c, const float* p) {
    return imhd_init_zpinch(c->buf[c->cur], p[0], c->gx, c->gy, c->gz, c->Nx, c->Ny, c->Nz, c->stream);
}

const IcEntry kInitializers[] = {
    {"screwpinch", 2, ic_screwpinch, 4},                  // J0, r_max_coeff            configurers.hpp:21
    {"screwpinch-stride", 1, ic_screwpinch_stride, 0},    // J0                         :24
    {"cubic-bennett-vortex", 0, ic_bennett, 2},           //                            :27
    {"cubic-bennett-vortex-m0", 2, ic_bennett_m0, 1},     // k, A                       ext (no_diffusion.cu:168)
    {"zpinch", 1, ic_zpinch, 3},                          // r_max_coeff                ext (no_diffusion.cu:169)
};

struct BundleEntry {
    const char* name;
    int path;  // IMHD_PATH_A, IMHD_PATH_B, or -1 = either
};
const BundleEntry kCorrectors[] = {{"fluidadvancelocal-nodiff", IMHD_PATH_A},   // configurers.hpp:54
                                   {"fluidadvancelocal", IMHD_PATH_B}};         // ext (main.cu:202)
const BundleEntry kPredictors[] = {{"corrector_advance-tp_nodiff", IMHD_PATH_A},      // :87
                                   {"corrector_advance-stride_nodiff", IMHD_PATH_A},  // :90 (same arithmetic, other launch shape)
                                   {"corrector_advance-stride", IMHD_PATH_B}};        // ext (main.cu:207)
const BundleEntry kFluidBcs[] = {{"pcrw-xy_pbc-z", -1}};                        // :130
const BundleEntry kPredictorBcs[] = {{"pbc-z", -1}};                            // :185

template <class T, int N>
constexpr int count_of(const T (&)[N]) { return N; }

const BundleEntry* find_bundle(const BundleEntry* t, int n, const char* key) {
    for (int q = 0; q < n; ++q)
        if (key && !strcmp(t[q].name, key)) return &t[q];
    return nullptr;
}

}  // namespace

extern "C" int imhd_registry_count(int kind) {
    switch (kind) {
        case IMHD_REG_INITIALIZER: return count_of(kInitializers);
        case IMHD_REG_CORRECTOR: return count_of(kCorrectors);
        case IMHD_REG_PREDICTOR: return count_of(kPredictors);
        case IMHD_REG_FLUID_BCS: return count_of(kFluidBcs);
        case IMHD_REG_PREDICTOR_BCS: return count_of(kPredictorBcs);
    }
    return 0;
}

extern "C" const char* imhd_registry_name(int kind, int index) {
    if (index < 0 || index >= imhd_registry_count(kind)) return nullptr;
    switch (kind) {
        case IMHD_REG_INITIALIZER: return kInitializers[index].name;
        case IMHD_REG_CORRECTOR: return kCorrectors[index].name;
        case IMHD_REG_PREDICTOR: return kPredictors[index].name;
        case IMHD_REG_FLUID_BCS: return kFluidBcs[index].name;
        case IMHD_REG_PREDICTOR_BCS: return kPredictorBcs[index].name;
    }
    return nullptr;
}

extern "C" int imhd_registry_initializer_nparams(const char* sim_type) {
    for (const IcEntry& e : kInitializers)
        if (sim_type && !strcmp(e.name, sim_type)) return e.nparams;
    return -1;
}

extern "C" int imhd_ctx_initialize(imhd_ctx* c, const char* sim_type, const float* params, int nparams) {
    CTX_CHECK(c);
    if (int e = need_grids(c)) return e;
    for (const IcEntry& e : kInitializers) {
        if (!sim_type || strcmp(e.name, sim_type)) continue;
        if (nparams != e.nparams || (e.nparams && !params)) {
            set_error("imhd_ctx_initialize: \"%s\" takes %d parameter(s), got %d", e.name, e.nparams, nparams);
            return IMHD_E_INVALID;
        }
        if (c->eng) return eng_init_ic(c->eng, e.ic, e.nparams > 0 ? params[0] : 0.f, e.nparams > 1 ? params[1] : 0.f);
        c->primed = false;
        return e.launch(c, params);
    }
    set_error("Unknown simulation type: %s", sim_type ? sim_type : "(null)");  // configurers.hpp:36
    return IMHD_E_INVALID;
}

extern "C" int imhd_registry_resolve_path(const char* corrector, const char* predictor, const char* fluid_bcs,
                                          const char* predictor_bcs, int* path) {
    const BundleEntry* co = find_bundle(kCorrectors, count_of(kCorrectors), corrector);
    if (!co) { set_error("Unknown kernel bundle selected: %s", corrector ? corrector : "(null)"); return IMHD_E_INVALID; }   // :67
    const BundleEntry* pr = find_bundle(kPredictors, count_of(kPredictors), predictor);
    if (!pr) { set_error("Unknown I.V. kernel bundle selection: %s", predictor ? predictor : "(null)"); return IMHD_E_INVALID; }  // :105
    if (!find_bundle(kFluidBcs, count_of(kFluidBcs), fluid_bcs)) {
        set_error("Unknown bcs selected: %s", fluid_bcs ? fluid_bcs : "(null)");  // :144
        return IMHD_E_INVALID;
    }
    if (!find_bundle(kPredictorBcs, count_of(kPredictorBcs), predictor_bcs)) {
        set_error("Unknown bcs selected: %s", predictor_bcs ? predictor_bcs : "(null)");  // :192
        return IMHD_E_INVALID;
    }
    if (co->path != pr->path) {
        set_error("corrector bundle \"%s\" and predictor bundle \"%s\" belong to different time loops (with / without diffusion)",
                  co->name, pr->name);
        return IMHD_E_INVALID;
    }
    if (path) *path = co->path;
    return 0;
}

extern "C" int imhd_ctx_stability(imhd_ctx* c, float dt, imhd_stability* host_out) {
    CTX_CHECK(c);
    if (c->eng) return eng_stability(c->eng, dt, host_out);
    imhd_slab s;
    memset(&s, 0, sizeof(s));
    s.Nx = c->Nx; s.Ny = c->Ny; s.Nz = c->Nz; s.k0 = 0; s.nzl = c->Nz; s.ghosts = 0;
    s.dt = dt; s.dx = c->dx; s.dy = c->dy; s.dz = c->dz;
    return imhd_stability_scan(c->buf[c->cur], &s, host_out, c->stream);
}

extern "C" int imhd_ctx_set_state(imhd_ctx* c, const float* host_Q) {
    CTX_CHECK(c);
    if (!host_Q) { set_error("imhd_ctx_set_state: null host buffer"); return IMHD_E_INVALID; }
    if (c->eng) return eng_set_state(c->eng, host_Q);
    c->primed = false;
    IMHD_CUDA(cudaMemcpyAsync(c->buf[c->cur], host_Q, 8 * c->cells * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

extern "C" int imhd_ctx_set_spacing(imhd_ctx* c, float dx, float dy, float dz) {
    CTX_CHECK(c);
    c->dx = dx; c->dy = dy; c->dz = dz;
    if (c->eng) return eng_set_spacing(c->eng, dx, dy, dz);
    return 0;
}

extern "C" int imhd_ctx_prime(imhd_ctx* c, int path, float D, float dt) {
    CTX_CHECK(c);
    if (path != IMHD_PATH_A && path != IMHD_PATH_B) { set_error("imhd_ctx_prime: bad path %d", path); return IMHD_E_INVALID; }
    if (c->eng) {
        if (int e = eng_prime(c->eng, path, D, dt)) return e;
        c->path = path; c->D = D; c->dt = dt; c->primed = true;
        return 0;
    }
    if (!(c->dx > 0.f) || !(c->dy > 0.f) || !(c->dz > 0.f)) {
        set_error("imhd_ctx_prime: grid spacing unset (call imhd_ctx_init_grids or imhd_ctx_set_spacing)");
        return IMHD_E_STATE;
    }
    c->path = path; c->D = D; c->dt = dt;
    // no_diffusion.cu:174-177: wall BCs + PBCs on the initial state (path A only; main.cu applies none)
    if (path == IMHD_PATH_A)
        if (int e = imhd_initial_bcs(c->buf[c->cur], c->Nx, c->Ny, c->Nz, c->stream)) return e;
    // The reference also computes Qint^0 here (no_diffusion.cu:183-199 / main.cu:108-111).  Qint is a pure
    // function of Q, so the fused path recomputes it on chip every step and the granular
    // path computes it lazily at the start of a step.
    c->corner_e = 0.0f;
    if (path == IMHD_PATH_B) {  // the column (Nx-1,Ny-1) holds this wall value at k=0 and k=Nz-1 from step 1 on (B-8)
        float e0 = 0.0f;
        const size_t l = (size_t)(c->Nx - 1) * c->Ny + (c->Ny - 1);
        IMHD_CUDA(cudaMemcpyAsync(&e0, c->buf[c->cur] + l + 7 * c->cells, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        IMHD_CUDA(cudaStreamSynchronize(c->stream));
        c->corner_e = imhd_wall_energy_fixed_point(e0, c->Nx);
    }
    c->qint_valid = false;
    c->primed = true;
    return 0;
}

extern "C" int imhd_ctx_step_granular(imhd_ctx* c, int nsteps) {
    CTX_CHECK(c);
    NOT_ON_SLABS(c, "imhd_ctx_step_granular");
    if (!c->primed) { set_error("imhd_ctx_step_granular before imhd_ctx_prime"); return IMHD_E_STATE; }
    float* Q = c->buf[c->cur];
    float* Qi = c->buf[1 - c->cur];
    for (int s = 0; s < nsteps; ++s) {
        if (!c->qint_valid)
            if (int e = imhd_predictor(Q, Qi, c->path, c->D, c->dt, c->dx, c->dy, c->dz, c->Nx, c->Ny, c->Nz, c->stream)) return e;
        if (int e = imhd_corrector(Q, Qi, c->path, c->D, c->dt, c->dx, c->dy, c->dz, c->Nx, c->Ny, c->Nz, c->stream)) return e;
        if (int e = imhd_fluid_bcs(Q, Qi, c->path, c->D, c->dt, c->dx, c->dy, c->dz, c->Nx, c->Ny, c->Nz, c->stream)) return e;
        c->qint_valid = false;
    }
    return 0;
}

extern "C" int imhd_ctx_step(imhd_ctx* c, int nsteps) {
    CTX_CHECK(c);
    if (c->eng) return eng_step(c->eng, nsteps);
    if (!c->primed) { set_error("imhd_ctx_step before imhd_ctx_prime"); return IMHD_E_STATE; }
    imhd_slab s;
    s.Nx = c->Nx; s.Ny = c->Ny; s.Nz = c->Nz; s.k0 = 0; s.nzl = c->Nz; s.ghosts = 0;
    s.path = c->path; s.D = c->D; s.dt = c->dt; s.dx = c->dx; s.dy = c->dy; s.dz = c->dz; s.corner_e = c->corner_e;
    const size_t pl8 = 8 * (size_t)c->Nx * c->Ny;
    for (int st = 0; st < nsteps; ++st) {
        const float* Qin = c->buf[c->cur];
        float* Qout = c->buf[1 - c->cur];
        // periodic wrap: Qint(Nz-1) == Qint(0) and Qint(-1) == Qint(Nz-2)
        float* q0 = c->qint_planes;
        float* qw = c->qint_planes + pl8;
        if (int e = imhd_qint_plane(Qin, q0, 0, &s, c->stream)) return e;
        if (c->path == IMHD_PATH_B)
            if (int e = imhd_qint_plane(Qin, qw, c->Nz - 2, &s, c->stream)) return e;
        if (int e = imhd_step_fused(Qin, Qout, q0, q0, qw, &s, c->stream)) return e;
        c->cur = 1 - c->cur;
    }
    c->qint_valid = false;
    return 0;
}

extern "C" int imhd_ctx_set_dt(imhd_ctx* c, float dt) {
    CTX_CHECK(c);
    if (!(dt > 0.f)) { set_error("imhd_ctx_set_dt: dt = %g", (double)dt); return IMHD_E_INVALID; }
    NOT_ON_SLABS(c, "imhd_ctx_set_dt");
    c->dt = dt;
    return 0;
}

// Adaptive time step without stalling the time loop.  Every `every` steps the CFL scan of the current state is queued behind
// the step that produced it (kernel + 16-byte D2H into pinned memory + an event); its result decides the dt of the group of
// steps that starts `every` steps LATER: the host only ever waits for a scan the device passed a whole group of steps ago, so
// the device always has `every` steps queued (the reference's README lists the CFL-driven dt as a TODO; its scanner is a
// forked host program run once, src/on-device/utils/compute_stability.cpp).
extern "C" int imhd_ctx_step_adaptive(imhd_ctx* c, int nsteps, int every, float cfl_target, float dt_max, float* dt_used) {
    CTX_CHECK(c);
    NOT_ON_SLABS(c, "imhd_ctx_step_adaptive");
    if (!c->primed) { set_error("imhd_ctx_step_adaptive before imhd_ctx_prime"); return IMHD_E_STATE; }
    if (every < 2 || !(cfl_target > 0.f) || !(dt_max > 0.f) || nsteps < 0) {
        set_error("imhd_ctx_step_adaptive: every = %d (>= 2), cfl_target = %g (> 0), dt_max = %g (> 0)", every, (double)cfl_target, (double)dt_max);
        return IMHD_E_INVALID;
    }
    if (!c->scan_pinned) {
        IMHD_CUDA(cudaMallocHost(&c->scan_pinned, 4 * sizeof(unsigned long long)));
        IMHD_CUDA(cudaEventCreateWithFlags(&c->scan_done[0], cudaEventDisableTiming));
        IMHD_CUDA(cudaEventCreateWithFlags(&c->scan_done[1], cudaEventDisableTiming));
    }
    imhd_slab s;
    memset(&s, 0, sizeof(s));
    s.Nx = c->Nx; s.Ny = c->Ny; s.Nz = c->Nz; s.k0 = 0; s.nzl = c->Nz; s.ghosts = 0;
    s.dx = c->dx; s.dy = c->dy; s.dz = c->dz;
    bool pending[2] = {false, false};
    float dt_scan[2] = {0.f, 0.f};
    int rc = 0;
    for (int it = 0; it < nsteps && rc == 0; ++it) {
        if (it % every == 0) {
            const int slot = (it / every) & 1, prev = 1 - slot;
            if (pending[prev]) {   // the scan taken `every` steps ago: the device passed it a group of steps ago
                IMHD_CUDA(cudaEventSynchronize(c->scan_done[prev]));
                pending[prev] = false;
                imhd_stability r;
                s.dt = dt_scan[prev];
                imhd_stability_decode(c->scan_pinned + 2 * prev, &s, &r);
                if (r.max_lhs > 0.f) {   // LHS is linear in dt: this dt puts the largest LHS of that state at cfl_target
                    const float dt_new = cfl_target * dt_scan[prev] / r.max_lhs;
                    c->dt = dt_new < dt_max ? dt_new : dt_max;
                }
            }
            s.dt = c->dt;
            if ((rc = imhd_stability_scan_async(c->buf[c->cur], &s, c->scan_pinned + 2 * slot, c->stream)) != 0) break;
            IMHD_CUDA(cudaEventRecord(c->scan_done[slot], c->stream));
            pending[slot] = true;
            dt_scan[slot] = c->dt;
        }
        if (dt_used) dt_used[it] = c->dt;
        rc = imhd_ctx_step(c, 1);
    }
    for (int q = 0; q < 2; ++q)   // nothing of this call writes into the pinned words after it returns
        if (pending[q]) cudaEventSynchronize(c->scan_done[q]);
    return rc;
}

extern "C" int imhd_ctx_get_state(imhd_ctx* c, float* host_Q) {
    CTX_CHECK(c);
    if (!host_Q) { set_error("imhd_ctx_get_state: null host buffer"); return IMHD_E_INVALID; }
    if (c->eng) return eng_get_state(c->eng, host_Q);
    IMHD_CUDA(cudaMemcpyAsync(host_Q, c->buf[c->cur], 8 * c->cells * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    IMHD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int imhd_ctx_get_grids(imhd_ctx* c, float* x, float* y, float* z) {
    CTX_CHECK(c);
    NOT_ON_SLABS(c, "imhd_ctx_get_grids");
    if (int e = need_grids(c)) return e;
    IMHD_CUDA(cudaMemcpyAsync(x, c->gx, sizeof(float) * c->Nx, cudaMemcpyDeviceToHost, c->stream));
    IMHD_CUDA(cudaMemcpyAsync(y, c->gy, sizeof(float) * c->Ny, cudaMemcpyDeviceToHost, c->stream));
    IMHD_CUDA(cudaMemcpyAsync(z, c->gz, sizeof(float) * c->Nz, cudaMemcpyDeviceToHost, c->stream));
    IMHD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" float* imhd_ctx_device_state(imhd_ctx* c) { return !c ? nullptr : (c->eng ? eng_device_state(c->eng, 0) : c->buf[c->cur]); }
extern "C" void* imhd_ctx_stream(imhd_ctx* c) { return !c ? nullptr : (c->eng ? eng_stream(c->eng, 0) : (void*)c->stream); }

// slab-local state transfer of a multi-slab context: the owned planes (8, nzl, Nx, Ny) of local slab q, host memory.
// After imhd_ctx_set_state_local on every slab (of every process) the next imhd_ctx_prime refreshes the ghost planes.
extern "C" int imhd_ctx_set_state_local(imhd_ctx* c, int q, const float* host_slab) {
    CTX_CHECK(c);
    if (!c->eng) return q == 0 ? imhd_ctx_set_state(c, host_slab) : IMHD_E_INVALID;
    return eng_set_state_local(c->eng, q, host_slab);
}
extern "C" int imhd_ctx_get_state_local(imhd_ctx* c, int q, float* host_slab) {
    CTX_CHECK(c);
    if (!c->eng) return q == 0 ? imhd_ctx_get_state(c, host_slab) : IMHD_E_INVALID;
    return eng_get_state_local(c->eng, q, host_slab);
}

extern "C" int imhd_ctx_synchronize(imhd_ctx* c) {
    CTX_CHECK(c);
    if (c->eng) return eng_synchronize(c->eng);
    IMHD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int imhd_run_host(imhd_ctx* c, const float* host_Q_in, float* host_Q_out, int path, float D, float dt,
                             float dx, float dy, float dz, int nsteps) {
    CTX_CHECK(c);
    if (int e = imhd_ctx_set_state(c, host_Q_in)) return e;
    if (int e = imhd_ctx_set_spacing(c, dx, dy, dz)) return e;
    if (int e = imhd_ctx_prime(c, path, D, dt)) return e;
    if (int e = imhd_ctx_step(c, nsteps)) return e;
    return imhd_ctx_get_state(c, host_Q_out);
}

// ---- output: fluidvars_<frame>.h5 / grid.h5 -----------------------------------------------
extern "C" int imhd_h5_write_fluidvars(const char*, const float*, int, int, int, int);
extern "C" int imhd_h5_write_grid(const char*, const float*, const float*, const float*, int, int, int);

static void writer_main(imhd_ctx* c) {
    cudaSetDevice(c->device);
    for (;;) {
        imhd_ctx::Job job;
        {
            std::unique_lock<std::mutex> lk(*c->mu);
            c->cv->wait(lk, [&] { return c->stop || !c->jobs->empty(); });
            if (c->jobs->empty()) return;
            job = c->jobs->front();
            c->jobs->pop_front();
        }
        cudaEventSynchronize(c->copy_done[job.slot]);
        const int rc = imhd_h5_write_fluidvars(job.path.c_str(), c->pinned[job.slot], c->Nx, c->Ny, c->Nz, job.attrs);
        {
            std::lock_guard<std::mutex> g(*c->mu);
            if (rc) {
                ++c->write_errors;
                snprintf(c->write_error_text, sizeof(c->write_error_text), "%s", imhd_last_error());
            }
            c->slot_busy[job.slot] = false;
        }
        c->cv->notify_all();
    }
}

static void free_writer_buffers(imhd_ctx* c) {
    cudaFree(c->snap); c->snap = nullptr;
    for (int s = 0; s < 2; ++s) {
        if (c->pinned[s]) cudaFreeHost(c->pinned[s]);
        if (c->copy_done[s]) cudaEventDestroy(c->copy_done[s]);
        c->pinned[s] = nullptr; c->copy_done[s] = nullptr;
    }
    if (c->snap_done) cudaEventDestroy(c->snap_done);
    if (c->snap_free) cudaEventDestroy(c->snap_free);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    c->snap_done = c->snap_free = nullptr; c->copy_stream = nullptr;
}

static int start_writer(imhd_ctx* c) {
    if (c->writer) return 0;
    const size_t bytes = 8 * c->cells * sizeof(float);
    bool ok = cudaMalloc(&c->snap, bytes) == cudaSuccess;
    for (int s = 0; s < 2 && ok; ++s)
        ok = cudaMallocHost(&c->pinned[s], bytes) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->copy_done[s], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->snap_done, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->snap_free, cudaEventDisableTiming) == cudaSuccess &&
         cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {  // roll back: nothing half-allocated survives, the next call starts from scratch
        const int e = cuda_fail(cudaGetLastError(), "output staging allocation", __FILE__, __LINE__);
        free_writer_buffers(c);
        return e ? e : IMHD_E_STATE;
    }
    c->snap_used = false;
    c->jobs = new std::deque<imhd_ctx::Job>();
    c->mu = new std::mutex();
    c->cv = new std::condition_variable();
    c->writer = new std::thread(writer_main, c);
    return 0;
}

// Queue frame `frame` of the current state for output as <dir>fluidvars_<frame>.h5 (attributes on frame 0 only, as
// main.cu:141-145).  Returns at once: a device-side snapshot (1 copy at HBM speed) decouples the time loop from the
// D2H copy and the file write, which run on a copy stream and a writer thread.  The compute stream only ever waits
// for the PREVIOUS frame's D2H copy (before it overwrites the snapshot buffer), never for this frame's.
extern "C" int imhd_ctx_write_frame(imhd_ctx* c, const char* dir, int frame) {
    CTX_CHECK(c);
    if (c->eng) {  // multi-slab context: gather the slabs into one host array and write it (synchronous)
        if (!dir) { set_error("imhd_ctx_write_frame: null directory"); return IMHD_E_INVALID; }
        std::vector<float> host(8 * c->cells);
        if (int e = eng_get_state(c->eng, host.data())) return e;
        return imhd_h5_write_fluidvars((std::string(dir) + "fluidvars_" + std::to_string(frame) + ".h5").c_str(), host.data(), c->Nx,
                                       c->Ny, c->Nz, frame == 0 ? 1 : 0);
    }
    if (!dir) { set_error("imhd_ctx_write_frame: null directory"); return IMHD_E_INVALID; }
    if (int e = start_writer(c)) return e;
    const int slot = c->next_slot;
    {   // wait until the staging slot's previous frame is on disk (back-pressure: at most 2 frames in flight)
        std::unique_lock<std::mutex> lk(*c->mu);
        c->cv->wait(lk, [&] { return !c->slot_busy[slot]; });
        c->slot_busy[slot] = true;
    }
    const size_t bytes = 8 * c->cells * sizeof(float);
    cudaError_t e = cudaSuccess;
    if (c->snap_used) e = cudaStreamWaitEvent(c->stream, c->snap_free, 0);  // previous frame drained out of `snap`?
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->snap, c->buf[c->cur], bytes, cudaMemcpyDeviceToDevice, c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(c->snap_done, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, c->snap_done, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->pinned[slot], c->snap, bytes, cudaMemcpyDeviceToHost, c->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(c->copy_done[slot], c->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(c->snap_free, c->copy_stream);
    if (e != cudaSuccess) {  // give the slot back, or the next frame on it (and flush) would wait for ever
        { std::lock_guard<std::mutex> g(*c->mu); c->slot_busy[slot] = false; }
        c->cv->notify_all();
        return cuda_fail(e, "imhd_ctx_write_frame", __FILE__, __LINE__);
    }
    c->snap_used = true;
    {
        std::lock_guard<std::mutex> g(*c->mu);
        c->jobs->push_back({slot, frame, frame == 0 ? 1 : 0, std::string(dir) + "fluidvars_" + std::to_string(frame) + ".h5"});
    }
    c->cv->notify_all();
    c->next_slot = 1 - slot;
    return 0;
}

// Block until every queued frame is on disk; returns IMHD_E_IO if any write failed.
extern "C" int imhd_ctx_flush_output(imhd_ctx* c) {
    CTX_CHECK(c);
    if (!c->writer) return 0;
    std::unique_lock<std::mutex> lk(*c->mu);
    c->cv->wait(lk, [&] { return c->jobs->empty() && !c->slot_busy[0] && !c->slot_busy[1]; });
    if (c->write_errors) {
        set_error("%d frame(s) could not be written; last error: %s", c->write_errors, c->write_error_text);
        return IMHD_E_IO;
    }
    return 0;
}

extern "C" int imhd_ctx_write_grid(imhd_ctx* c, const char* dir) {
    CTX_CHECK(c);
    if (!dir) { set_error("imhd_ctx_write_grid: null directory"); return IMHD_E_INVALID; }
    if (int e = need_grids(c)) return e;
    if (c->eng) {  // the same fp32 expression as the device kernel (x[i] = x_min + i * dx, initialize_od.cu:26-57)
        std::vector<float> x(c->Nx), y(c->Ny), z(c->Nz);
        for (int i = 0; i < c->Nx; ++i) { volatile float t = (float)(unsigned)i * c->dx; x[i] = c->bounds[0] + t; }
        for (int i = 0; i < c->Ny; ++i) { volatile float t = (float)(unsigned)i * c->dy; y[i] = c->bounds[2] + t; }
        for (int i = 0; i < c->Nz; ++i) { volatile float t = (float)(unsigned)i * c->dz; z[i] = c->bounds[4] + t; }
        return imhd_h5_write_grid((std::string(dir) + "grid.h5").c_str(), x.data(), y.data(), z.data(), c->Nx, c->Ny, c->Nz);
    }
    std::vector<float> x(c->Nx), y(c->Ny), z(c->Nz);
    if (int e = imhd_ctx_get_grids(c, x.data(), y.data(), z.data())) return e;
    return imhd_h5_write_grid((std::string(dir) + "grid.h5").c_str(), x.data(), y.data(), z.data(), c->Nx, c->Ny, c->Nz);
}
