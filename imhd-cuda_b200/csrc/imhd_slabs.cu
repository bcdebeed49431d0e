// Multi-GPU engine: the z-slab time loop in C++ behind the C ABI (the reference is single-GPU:
// src/on-device/main.cu:196-238 / no_diffusion.cu:284-337 is the loop this replaces).
//
// The domain is cut into `world` z-slabs (k is the slowest index of the reference layout, so a slab is one contiguous
// plane range per variable).  A slab lives on one GPU as a ghosted (8, nzl+2, Nx, Ny) array, ping-pong buffered.  The
// slabs this process owns are driven from one host thread:
//     imhd_create_multi   every slab in this process (one per device; NCCL communicators from ncclCommInitAll)
//     imhd_create_slab    one slab per process (torchrun: one process per GPU; communicator from a shared unique id)
// Both are the same code: per step and slab
//   1. the fused kernel on the planes next to the two slab ends (imhd_step_fused_ends: ONE launch, EDGE planes each), an event,
//      then the interior planes -- all on the slab's main stream;
//   2. on the slab's side stream, under the interior launch: the new end planes go to the neighbours' ghost planes
//      (for path A the plane the last slab sends up is the periodic copy Q[.,.,0] <- Q[.,.,Nz-1],
//      lib/on-device/kernels_fluidbcs.cu:498-510), then the two predictor planes of the NEW state the neighbours need
//      (imhd_qint_plane) are exchanged into the other half of a double-buffered plane set;
//   3. the main stream waits for the side stream before the next step.
// The exchanges are ncclSend/ncclRecv ring neighbours over NVLink (one ncclGroup for all local slabs); there is no
// reduction on the path (the reference has no global dt control), so there is no collective to fuse into a kernel.
// Every value is computed by the same device function from the same inputs as in the single-GPU context, so results
// are bit-identical for any number of slabs (tests).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a single-GPU user needs no NCCL, and inside a Python process
// that has imported torch the already-loaded copy is reused instead of a second one.
#include <cuda.h>
#include <dlfcn.h>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <nccl.h>

#include <vector>

#include "imhd_common.cuh"
#include "imhd_engine.h"

namespace imhd {

namespace {

struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId;
    decltype(&ncclCommInitRank) CommInitRank;
    decltype(&ncclCommInitAll) CommInitAll;
    decltype(&ncclCommDestroy) CommDestroy;
    decltype(&ncclSend) Send;
    decltype(&ncclRecv) Recv;
    decltype(&ncclGroupStart) GroupStart;
    decltype(&ncclGroupEnd) GroupEnd;
    decltype(&ncclBroadcast) Broadcast;
    decltype(&ncclAllGather) AllGather;
    decltype(&ncclGetErrorString) GetErrorString;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static int state = 0;  // 0 untried, 1 ok, -1 failed
    if (state == 0) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        state = -1;
        if (h) {
            bool ok = true;
#define IMHD_SYM(field, name) ok = ok && (api.field = (decltype(api.field))dlsym(h, name)) != nullptr
            IMHD_SYM(GetUniqueId, "ncclGetUniqueId");
            IMHD_SYM(CommInitRank, "ncclCommInitRank");
            IMHD_SYM(CommInitAll, "ncclCommInitAll");
            IMHD_SYM(CommDestroy, "ncclCommDestroy");
            IMHD_SYM(Send, "ncclSend");
            IMHD_SYM(Recv, "ncclRecv");
            IMHD_SYM(GroupStart, "ncclGroupStart");
            IMHD_SYM(GroupEnd, "ncclGroupEnd");
            IMHD_SYM(Broadcast, "ncclBroadcast");
            IMHD_SYM(AllGather, "ncclAllGather");
            IMHD_SYM(GetErrorString, "ncclGetErrorString");
#undef IMHD_SYM
            if (ok) state = 1;
        }
    }
    if (state != 1) {
        set_error("multi-GPU needs NCCL: libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbol");
        return nullptr;
    }
    return &api;
}

int nccl_fail(ncclResult_t r, const char* what) {
    NcclApi* n = nccl_api();
    set_error("NCCL error in %s: %s", what, n ? n->GetErrorString(r) : "?");
    return IMHD_E_STATE;
}
#define IMHD_NCCL(call)                                        \
    do {                                                       \
        ncclResult_t r__ = (call);                             \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);  \
    } while (0)

// dst/src: 8 chunks of `n` floats, `dvs` / `svs` floats apart (one plane of a state array <-> a packed (8,Nx,Ny) plane)
__global__ void __launch_bounds__(256) k_copy8(float* __restrict__ dst, long long dvs, const float* __restrict__ src, long long svs,
                                               long long n) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
#pragma unroll
    for (int v = 0; v < 8; ++v) dst[c + v * dvs] = src[c + v * svs];
}

int g_edge = 4;

}  // namespace

void eng_set_edge(int planes) { g_edge = planes < 3 ? 3 : planes; }

struct SlabDev {
    int dev, rank, k0, nzl;
    float* Q[2];              // (8, nzl+2, Nx, Ny): array plane 0 is global plane k0-1
    float *gx, *gy, *gz;      // gz: the array's own planes (nzl+2 entries)
    float* planes;            // [2 sets][lo, hi, up, down] + send_up, send_down, 2 x (recv_lo, recv_hi), each (8,Nx,Ny)
    float* scratch;           // 16 floats (corner energy broadcast, stability rows)
    cudaStream_t main, side, ends;   // ends: highest priority -- the slab-end launch runs beside the interior launch, its blocks first
    cudaEvent_t ev_edges, ev_comm, ev_fork;
    ncclComm_t comm;
    // direct exchange: the neighbours' `planes` allocations as this process sees them ([0] the slab above, [1] the slab below)
    float* peer[2];
    bool peer_ipc[2];          // mapped with cudaIpcOpenMemHandle (closed on destroy)
    unsigned seq;              // exchanges posted so far: the value the arrival counters reach
    imhd_slab desc;
    float* set(int which, int idx, size_t pl8) const { return planes + ((size_t)which * 4 + idx) * pl8; }
    float* stage(int idx, size_t pl8) const { return planes + (8 + (size_t)idx) * pl8; }
    // four 32-bit words behind the fourteen planes: [0] / [1] arrival counters of the up-going / down-going message (written by
    // the slab below / above), [2] the sequence number this slab publishes
    static size_t words_at(size_t pl8) { return 14 * pl8; }
};
constexpr size_t kPlanesTail = 64;   // floats
enum { P_LO = 0, P_HI = 1, P_UP = 2, P_DOWN = 3 };
// the receive staging planes are double buffered by the parity of the exchange count: a neighbour that is one exchange ahead
// (it may post exchange n + 1 as soon as this slab has POSTED n, before this slab has unpacked n) writes the other pair
enum { S_SEND_UP = 0, S_SEND_DOWN = 1, S_RECV_LO = 2, S_RECV_HI = 3 };
static int recv_stage(int idx, unsigned seq) { return idx + 2 * (int)(seq & 1u); }

struct Engine {
    int Nx, Ny, Nz, world;
    size_t plane, pl8;
    std::vector<SlabDev> s;
    NcclApi* nccl;
    int cur, qcur, path;
    bool q_ready, primed, have_grids, overlap, ends_concurrent;
    bool direct;               // plane exchange by the copy engines into the neighbours' buffers (else ncclSend / ncclRecv)
    float D, dt, dx, dy, dz, corner_e;
    float bounds[6];
};

static int slab_k0(int Nz, int world, int r) { return (int)(((long long)Nz * r) / world); }

#define ENG_DEV(sl) IMHD_CUDA(cudaSetDevice((sl).dev))

void eng_destroy(Engine* e) {
    if (!e) return;
    for (SlabDev& sl : e->s) {
        cudaSetDevice(sl.dev);
        if (sl.main) cudaStreamSynchronize(sl.main);
        if (sl.side) cudaStreamSynchronize(sl.side);
        if (sl.ends) cudaStreamSynchronize(sl.ends);
        // teardown order: this slab's streams are idle (every plane a neighbour sent it has landed: its side stream waited
        // for them), so the neighbours' buffers mapped here are no longer written -> unmap them, THEN leave the communicator,
        // THEN free what this slab exported (an exporter should not free a region an importer still maps; every slab of the
        // domain is destroyed collectively, like the communicator)
        for (int d = 0; d < 2; ++d)
            if (sl.peer_ipc[d] && sl.peer[d]) cudaIpcCloseMemHandle(sl.peer[d]);
        if (sl.comm && e->nccl) e->nccl->CommDestroy(sl.comm);
        cudaFree(sl.Q[0]); cudaFree(sl.Q[1]); cudaFree(sl.gx); cudaFree(sl.gy); cudaFree(sl.gz);
        cudaFree(sl.planes); cudaFree(sl.scratch);
        if (sl.ev_edges) cudaEventDestroy(sl.ev_edges);
        if (sl.ev_comm) cudaEventDestroy(sl.ev_comm);
        if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
        if (sl.ends) cudaStreamDestroy(sl.ends);
        if (sl.main) cudaStreamDestroy(sl.main);
        if (sl.side) cudaStreamDestroy(sl.side);
    }
    delete e;
}

// ---- direct plane exchange ------------------------------------------------------------------------------------------
// ncclSend / ncclRecv run as kernels of their own (one 512+-thread block per channel): they need whole SMs, and the interior
// launch under which the exchange is meant to hide holds every SM with one 8-warp block for a third of the step -- the NCCL
// kernels wait for the first blocks to retire and then delay the blocks behind them (measured on 2 B200: 1.608 ms per step
// against 1.480 ms with the exchange switched off, and 1.81 ms with NCCL held to one or two channels).  So the planes travel
// by the COPY ENGINES instead: a slab writes its packed plane straight into the neighbour's receive buffer over NVLink
// (cudaMemcpyAsync to the peer's allocation -- mapped with CUDA IPC when the neighbour is another process, used as it is
// when it is a slab of this process), then a 32-bit sequence number into the neighbour's arrival counter (a second copy on
// the same stream: ordered behind the data), and the receiving stream waits for its counter with cuStreamWaitValue32.  No
// SM is involved; pack / unpack / predictor-plane kernels are small enough to be co-resident with a marching block.
// The buffer discipline is the one the NCCL path has: every slab waits for BOTH neighbours' planes of exchange n before it
// posts exchange n + 1, and a receive buffer is only rewritten two exchanges of the same kind later.
// NCCL stays for the set-up (handle all-gather), the corner-energy broadcast and the CFL all-gather, and as the fallback
// (another node, no peer access, IMHD_SLAB_EXCHANGE=nccl).
typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct StreamOps { StreamValue32Fn wait = nullptr, write = nullptr; };
static StreamOps* stream_ops() {
    static StreamOps ops;
    static int state = 0;
    if (state == 0) {
        state = -1;
        void *w = nullptr, *r = nullptr;
        cudaDriverEntryPointQueryResult q1, q2;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &w, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuStreamWriteValue32", &r, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess && w && r) {
            ops.wait = (StreamValue32Fn)w;
            ops.write = (StreamValue32Fn)r;
            state = 1;
        }
        cudaGetLastError();
    }
    return state == 1 ? &ops : nullptr;
}

static SlabDev* local_slab(Engine* e, int rank) {
    for (SlabDev& sl : e->s)
        if (sl.rank == rank) return &sl;
    return nullptr;
}

// Collective over all slabs of the domain: maps the neighbours' buffers; e->direct ends up true on every slab or on none.
static int setup_direct(Engine* e) {
    e->direct = false;
    NcclApi* n = e->nccl;
    const int nl = (int)e->s.size(), W = e->world;
    const char* mode = getenv("IMHD_SLAB_EXCHANGE");
    int want = !(mode && strcmp(mode, "nccl") == 0) && stream_ops() != nullptr;
    // one row per slab: [0] ok so far, [1..16] the IPC handle of its `planes` allocation (64 bytes), [17] process id
    constexpr int ROW = 20;
    std::vector<int> mine((size_t)nl * ROW, 0), all((size_t)nl * W * ROW, 0);
    for (int q = 0; q < nl; ++q) {
        SlabDev& sl = e->s[q];
        ENG_DEV(sl);
        int* row = &mine[(size_t)q * ROW];
        cudaIpcMemHandle_t h;
        static_assert(sizeof(h) == 64, "CUDA IPC handle size");
        row[0] = want && cudaIpcGetMemHandle(&h, sl.planes) == cudaSuccess;
        if (row[0]) memcpy(row + 1, &h, sizeof(h));
        row[17] = (int)getpid();
        cudaGetLastError();
    }
    auto gather = [&]() -> int {   // row q of every local slab -> all rows, on every slab
        std::vector<int*> dsend(nl, nullptr), drecv(nl, nullptr);
        for (int q = 0; q < nl; ++q) {
            SlabDev& sl = e->s[q];
            ENG_DEV(sl);
            IMHD_CUDA(cudaMalloc(&dsend[q], sizeof(int) * ROW));
            IMHD_CUDA(cudaMalloc(&drecv[q], sizeof(int) * ROW * W));
            IMHD_CUDA(cudaMemcpyAsync(dsend[q], &mine[(size_t)q * ROW], sizeof(int) * ROW, cudaMemcpyHostToDevice, sl.main));
        }
        IMHD_NCCL(n->GroupStart());
        for (int q = 0; q < nl; ++q) {
            cudaSetDevice(e->s[q].dev);
            IMHD_NCCL(n->AllGather(dsend[q], drecv[q], ROW, ncclInt, e->s[q].comm, e->s[q].main));
        }
        IMHD_NCCL(n->GroupEnd());
        for (int q = 0; q < nl; ++q) {
            SlabDev& sl = e->s[q];
            ENG_DEV(sl);
            IMHD_CUDA(cudaMemcpyAsync(&all[(size_t)q * W * ROW], drecv[q], sizeof(int) * ROW * W, cudaMemcpyDeviceToHost, sl.main));
            IMHD_CUDA(cudaStreamSynchronize(sl.main));
            cudaFree(dsend[q]); cudaFree(drecv[q]);
        }
        return 0;
    };
    if (int rc = gather()) return rc;
    // map the two neighbours of every local slab
    for (int q = 0; q < nl; ++q) {
        SlabDev& sl = e->s[q];
        ENG_DEV(sl);
        const int* rows = &all[(size_t)q * W * ROW];
        int ok = 1;
        for (int r = 0; r < W; ++r) ok = ok && rows[(size_t)r * ROW];
        const int nb[2] = {(sl.rank + 1) % W, (sl.rank + W - 1) % W};
        for (int d = 0; d < 2 && ok; ++d) {
            if (SlabDev* other = local_slab(e, nb[d])) {   // a slab of this process: its pointer as it is, peer access if the devices offer it
                sl.peer[d] = other->planes;
                int can = 0;
                if (other->dev != sl.dev && cudaDeviceCanAccessPeer(&can, sl.dev, other->dev) == cudaSuccess && can) {
                    const cudaError_t pe = cudaDeviceEnablePeerAccess(other->dev, 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                }
                cudaGetLastError();
            } else if (d == 1 && nb[1] == nb[0] && sl.peer[0]) {   // two slabs in two processes: one neighbour, one mapping
                sl.peer[1] = sl.peer[0];
            } else {
                cudaIpcMemHandle_t h;
                memcpy(&h, rows + (size_t)nb[d] * ROW + 1, sizeof(h));
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess && ptr) {
                    sl.peer[d] = (float*)ptr;
                    sl.peer_ipc[d] = true;
                } else {
                    cudaGetLastError();
                    ok = 0;
                }
            }
        }
        mine[(size_t)q * ROW] = ok;
    }
    // second round: everybody mapped everything?
    if (int rc = gather()) return rc;
    int ok = 1;
    for (int r = 0; r < W; ++r) ok = ok && all[(size_t)r * ROW];
    e->direct = ok != 0;
    if (!e->direct)
        for (SlabDev& sl : e->s)
            for (int d = 0; d < 2; ++d) {
                if (sl.peer_ipc[d] && sl.peer[d]) { cudaSetDevice(sl.dev); cudaIpcCloseMemHandle(sl.peer[d]); }
                sl.peer[d] = nullptr; sl.peer_ipc[d] = false;
            }
    return 0;
}

Engine* eng_create(int Nx, int Ny, int Nz, int world, int nlocal, const int* ranks, const int* devices, const void* uid) {
    if (bad_dims(Nx, Ny, Nz)) return nullptr;
    if (world < 2 || nlocal < 1 || nlocal > world || !ranks || !devices) {
        set_error("slab engine: bad decomposition (world=%d, local slabs=%d)", world, nlocal);
        return nullptr;
    }
    if (Nz / world < 3) { set_error("Nz=%d is too thin for %d slabs (need >= 3 planes per slab)", Nz, world); return nullptr; }
    if (!uid && nlocal != world) { set_error("slab engine: a process that holds only some slabs needs a shared NCCL unique id"); return nullptr; }
    NcclApi* n = nccl_api();
    if (!n) return nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("slab engine: no usable CUDA device; this library has no CPU path");
        return nullptr;
    }
    Engine* e = new Engine();
    e->Nx = Nx; e->Ny = Ny; e->Nz = Nz; e->world = world;
    e->plane = (size_t)Nx * Ny; e->pl8 = 8 * e->plane;
    e->nccl = n;
    e->cur = e->qcur = 0; e->path = IMHD_PATH_B;
    e->q_ready = e->primed = e->have_grids = false;
    e->D = e->dt = e->dx = e->dy = e->dz = e->corner_e = 0.f;
    e->s.resize(nlocal);
    bool ok = true;
    int min_nzl = Nz;
    for (int q = 0; q < nlocal && ok; ++q) {
        SlabDev& sl = e->s[q];
        memset(&sl, 0, sizeof(sl));
        sl.dev = devices[q]; sl.rank = ranks[q];
        if (sl.dev < 0 || sl.dev >= ndev || sl.rank < 0 || sl.rank >= world) { set_error("slab engine: bad device %d / slab %d", sl.dev, sl.rank); ok = false; break; }
        sl.k0 = slab_k0(Nz, world, sl.rank);
        sl.nzl = slab_k0(Nz, world, sl.rank + 1) - sl.k0;
        min_nzl = sl.nzl < min_nzl ? sl.nzl : min_nzl;
        const size_t qbytes = e->pl8 * (size_t)(sl.nzl + 2) * sizeof(float);
        int prio_lo = 0, prio_hi = 0;
        if (cudaSetDevice(sl.dev) == cudaSuccess) cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        ok = cudaSetDevice(sl.dev) == cudaSuccess && cudaMalloc(&sl.Q[0], qbytes) == cudaSuccess && cudaMalloc(&sl.Q[1], qbytes) == cudaSuccess &&
             cudaMalloc(&sl.gx, sizeof(float) * Nx) == cudaSuccess && cudaMalloc(&sl.gy, sizeof(float) * Ny) == cudaSuccess &&
             cudaMalloc(&sl.gz, sizeof(float) * (sl.nzl + 2)) == cudaSuccess &&
             cudaMalloc(&sl.planes, (14 * e->pl8 + kPlanesTail) * sizeof(float)) == cudaSuccess &&
             cudaMemset(sl.planes + 14 * e->pl8, 0, kPlanesTail * sizeof(float)) == cudaSuccess && cudaMalloc(&sl.scratch, 64 * sizeof(double)) == cudaSuccess &&
             cudaMemset(sl.Q[0], 0, qbytes) == cudaSuccess && cudaMemset(sl.Q[1], 0, qbytes) == cudaSuccess &&
             cudaStreamCreateWithFlags(&sl.main, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&sl.side, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithPriority(&sl.ends, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_edges, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_comm, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) cuda_fail(cudaGetLastError(), "slab engine allocation", __FILE__, __LINE__);
    }
    if (ok) {
        ncclResult_t r = ncclSuccess;
        if (uid) {
            ncclUniqueId id;
            memcpy(&id, uid, sizeof(id));
            r = n->GroupStart();
            for (int q = 0; q < nlocal && r == ncclSuccess; ++q) {
                cudaSetDevice(e->s[q].dev);
                r = n->CommInitRank(&e->s[q].comm, world, id, e->s[q].rank);
            }
            if (r == ncclSuccess) r = n->GroupEnd();
        } else {
            std::vector<ncclComm_t> comms(world);
            std::vector<int> devs(world);
            for (int q = 0; q < nlocal; ++q) devs[e->s[q].rank] = e->s[q].dev;
            r = n->CommInitAll(comms.data(), world, devs.data());
            if (r == ncclSuccess)
                for (int q = 0; q < nlocal; ++q) e->s[q].comm = comms[e->s[q].rank];
        }
        if (r != ncclSuccess) { nccl_fail(r, "communicator set-up"); ok = false; }
    }
    if (!ok) { eng_destroy(e); return nullptr; }
    if (setup_direct(e) != 0) { eng_destroy(e); return nullptr; }
    // every slab of the domain has to take the same decision (min over all slabs: floor(Nz/world))
    e->overlap = Nz / world >= 2 * g_edge + 2;
    {   // test / measurement hook: IMHD_SLAB_ENDS_CONCURRENT=0 puts the end launch in front of the interior launch on one stream
        const char* v = getenv("IMHD_SLAB_ENDS_CONCURRENT");
        e->ends_concurrent = !(v && v[0] == '0');
    }
    (void)min_nzl;
    return e;
}

int eng_nlocal(const Engine* e) { return e ? (int)e->s.size() : 0; }

int eng_local_extent(const Engine* e, int q, int* k0, int* nzl, int* device) {
    if (!e || q < 0 || q >= (int)e->s.size()) { set_error("no such local slab %d", q); return IMHD_E_INVALID; }
    if (k0) *k0 = e->s[q].k0;
    if (nzl) *nzl = e->s[q].nzl;
    if (device) *device = e->s[q].dev;
    return 0;
}

float* eng_device_state(Engine* e, int q) { return e && q >= 0 && q < (int)e->s.size() ? e->s[q].Q[e->cur] : nullptr; }
void* eng_stream(Engine* e, int q) { return e && q >= 0 && q < (int)e->s.size() ? (void*)e->s[q].main : nullptr; }

static void fill_desc(Engine* e) {
    for (SlabDev& sl : e->s) {
        imhd_slab& d = sl.desc;
        memset(&d, 0, sizeof(d));
        d.Nx = e->Nx; d.Ny = e->Ny; d.Nz = e->Nz; d.k0 = sl.k0; d.nzl = sl.nzl; d.ghosts = 1; d.path = e->path;
        d.D = e->D; d.dt = e->dt; d.dx = e->dx; d.dy = e->dy; d.dz = e->dz; d.corner_e = e->corner_e;
    }
}

int eng_init_grids(Engine* e, const float b[6]) {
    memcpy(e->bounds, b, sizeof(e->bounds));
    e->dx = (b[1] - b[0]) / (e->Nx - 1);  // fp32, as main.cu:98-100
    e->dy = (b[3] - b[2]) / (e->Ny - 1);
    e->dz = (b[5] - b[4]) / (e->Nz - 1);
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        if (int rc = imhd_init_axis_slab(sl.gx, b[0], e->dx, e->Nx, 0, sl.main)) return rc;
        if (int rc = imhd_init_axis_slab(sl.gy, b[2], e->dy, e->Ny, 0, sl.main)) return rc;
        if (int rc = imhd_init_axis_slab(sl.gz, b[4], e->dz, sl.nzl + 2, sl.k0 - 1, sl.main)) return rc;
    }
    e->have_grids = true;
    return 0;
}

int eng_init_ic(Engine* e, int ic, float a, float b) {
    if (!e->have_grids) { set_error("initial condition requested before imhd_ctx_init_grids"); return IMHD_E_STATE; }
    e->primed = false; e->q_ready = false;
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        if (ic == 4)  // ScrewPinch leaves seven variables untouched outside the pinch: give them a defined value
            IMHD_CUDA(cudaMemsetAsync(sl.Q[e->cur], 0, e->pl8 * (size_t)(sl.nzl + 2) * sizeof(float), sl.main));
        if (int rc = imhd_init_ic_slab(ic, sl.Q[e->cur], a, b, sl.gx, sl.gy, sl.gz, e->Nx, e->Ny, sl.nzl + 2, sl.k0 - 1, sl.main)) return rc;
    }
    return 0;
}

static int copy_planes(Engine* e, SlabDev& sl, float* dev_Q, const float* host_src, float* host_dst, int glo, int ghi, size_t host_vs,
                       int host_k0) {
    // global planes [glo, ghi) between the device array (plane 0 = k0-1) and a host array whose plane 0 is host_k0
    const size_t n = (size_t)(ghi - glo) * e->plane;
    const size_t dvs = (size_t)(sl.nzl + 2) * e->plane;
    for (int v = 0; v < 8; ++v) {
        float* d = dev_Q + v * dvs + (size_t)(glo - sl.k0 + 1) * e->plane;
        const size_t ho = v * host_vs + (size_t)(glo - host_k0) * e->plane;
        if (host_src) IMHD_CUDA(cudaMemcpyAsync(d, host_src + ho, n * sizeof(float), cudaMemcpyHostToDevice, sl.main));
        else          IMHD_CUDA(cudaMemcpyAsync(host_dst + ho, d, n * sizeof(float), cudaMemcpyDeviceToHost, sl.main));
    }
    return 0;
}

int eng_set_state(Engine* e, const float* host_full) {
    if (!host_full) { set_error("imhd_ctx_set_state: null host buffer"); return IMHD_E_INVALID; }
    e->primed = false; e->q_ready = false;
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        const int lo = sl.k0 - 1 < 0 ? 0 : sl.k0 - 1, hi = sl.k0 + sl.nzl + 1 > e->Nz ? e->Nz : sl.k0 + sl.nzl + 1;
        if (int rc = copy_planes(e, sl, sl.Q[e->cur], host_full, nullptr, lo, hi, (size_t)e->Nz * e->plane, 0)) return rc;
    }
    return 0;
}

int eng_synchronize(Engine* e) {
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        IMHD_CUDA(cudaStreamSynchronize(sl.main));
        IMHD_CUDA(cudaStreamSynchronize(sl.side));
    }
    return 0;
}

int eng_get_state(Engine* e, float* host_full) {
    if (!host_full) { set_error("imhd_ctx_get_state: null host buffer"); return IMHD_E_INVALID; }
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        if (int rc = copy_planes(e, sl, sl.Q[e->cur], nullptr, host_full, sl.k0, sl.k0 + sl.nzl, (size_t)e->Nz * e->plane, 0)) return rc;
    }
    return eng_synchronize(e);
}

int eng_set_state_local(Engine* e, int q, const float* host_slab) {
    if (q < 0 || q >= (int)e->s.size() || !host_slab) { set_error("imhd_ctx_set_state_local: bad slab / null buffer"); return IMHD_E_INVALID; }
    SlabDev& sl = e->s[q];
    ENG_DEV(sl);
    e->primed = false; e->q_ready = false;
    return copy_planes(e, sl, sl.Q[e->cur], host_slab, nullptr, sl.k0, sl.k0 + sl.nzl, (size_t)sl.nzl * e->plane, sl.k0);
}

int eng_get_state_local(Engine* e, int q, float* host_slab) {
    if (q < 0 || q >= (int)e->s.size() || !host_slab) { set_error("imhd_ctx_get_state_local: bad slab / null buffer"); return IMHD_E_INVALID; }
    SlabDev& sl = e->s[q];
    ENG_DEV(sl);
    if (int rc = copy_planes(e, sl, sl.Q[e->cur], nullptr, host_slab, sl.k0, sl.k0 + sl.nzl, (size_t)sl.nzl * e->plane, sl.k0)) return rc;
    IMHD_CUDA(cudaStreamSynchronize(sl.main));
    return 0;
}

int eng_set_spacing(Engine* e, float dx, float dy, float dz) {
    e->dx = dx; e->dy = dy; e->dz = dz;
    return 0;
}

// ---- exchanges ---------------------------------------------------------------------------------------------------
static cudaStream_t stream_of(SlabDev& sl, bool side) { return side ? sl.side : sl.main; }

// Kernels that run on the side stream under the interior launch use blocks small enough to fit into what a resident block of
// the marching kernel AND its co-resident strip block leave of an SM (4608 registers), so they never wait for an SM of their
// own (measured on 4 B200: the same step time as with 256-thread blocks -- kept because it cannot queue behind the strip).
static int copy8(Engine* e, float* dst, long long dvs, const float* src, long long svs, cudaStream_t st, bool small = false) {
    const unsigned threads = small ? 64 : 256;
    k_copy8<<<(unsigned)((e->plane + threads - 1) / threads), threads, 0, st>>>(dst, dvs, src, svs, (long long)e->plane);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

// one ring exchange of a packed plane per direction, all local slabs in one NCCL group:
//   send_up -> slab above (its recv_from_down), send_down -> slab below (its recv_from_up)
static int ring_exchange(Engine* e, int i_send_up, int i_send_down, int i_recv_down, int i_recv_up, bool is_set, int which, bool side) {
    for (SlabDev& sl : e->s) ++sl.seq;   // the same count on every slab of the domain: every exchange is collective
    if (!is_set) {                       // staging planes: this exchange's receive pair
        i_recv_down = recv_stage(i_recv_down, e->s[0].seq);
        i_recv_up = recv_stage(i_recv_up, e->s[0].seq);
    }
    if (e->direct) {
        StreamOps* ops = stream_ops();
        const size_t bytes = e->pl8 * sizeof(float), words = SlabDev::words_at(e->pl8);
        auto off = [&](int idx) { return is_set ? ((size_t)which * 4 + idx) * e->pl8 : (8 + (size_t)idx) * e->pl8; };   // as set() / stage()
        for (SlabDev& sl : e->s) {   // post: plane, then sequence number, to either neighbour
            ENG_DEV(sl);
            cudaStream_t st = stream_of(sl, side);
            const unsigned seq = sl.seq;
            if (ops->write((CUstream)st, (CUdeviceptr)(sl.planes + words + 2), seq, 0) != CUDA_SUCCESS) { set_error("cuStreamWriteValue32 failed"); return IMHD_E_STATE; }
            IMHD_CUDA(cudaMemcpyAsync(sl.peer[0] + off(i_recv_down), sl.planes + off(i_send_up), bytes, cudaMemcpyDefault, st));   // up-going
            IMHD_CUDA(cudaMemcpyAsync(sl.peer[0] + words + 0, sl.planes + words + 2, sizeof(unsigned), cudaMemcpyDefault, st));
            IMHD_CUDA(cudaMemcpyAsync(sl.peer[1] + off(i_recv_up), sl.planes + off(i_send_down), bytes, cudaMemcpyDefault, st));   // down-going
            IMHD_CUDA(cudaMemcpyAsync(sl.peer[1] + words + 1, sl.planes + words + 2, sizeof(unsigned), cudaMemcpyDefault, st));
        }
        for (SlabDev& sl : e->s) {   // both neighbours' planes have landed
            ENG_DEV(sl);
            cudaStream_t st = stream_of(sl, side);
            for (int w = 0; w < 2; ++w)
                if (ops->wait((CUstream)st, (CUdeviceptr)(sl.planes + words + w), sl.seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) { set_error("cuStreamWaitValue32 failed"); return IMHD_E_STATE; }
        }
        return 0;
    }
    NcclApi* n = e->nccl;
    IMHD_NCCL(n->GroupStart());
    for (SlabDev& sl : e->s) {
        const int up = (sl.rank + 1) % e->world, down = (sl.rank + e->world - 1) % e->world;
        auto buf = [&](int idx) { return is_set ? sl.set(which, idx, e->pl8) : sl.stage(idx, e->pl8); };
        cudaStream_t st = stream_of(sl, side);
        cudaSetDevice(sl.dev);
        // between the same pair (two slabs: up == down) messages match in posting order: up-going first on both sides
        IMHD_NCCL(n->Send(buf(i_send_up), e->pl8, ncclFloat, up, sl.comm, st));
        IMHD_NCCL(n->Recv(buf(i_recv_down), e->pl8, ncclFloat, down, sl.comm, st));
        IMHD_NCCL(n->Send(buf(i_send_down), e->pl8, ncclFloat, down, sl.comm, st));
        IMHD_NCCL(n->Recv(buf(i_recv_up), e->pl8, ncclFloat, up, sl.comm, st));
    }
    IMHD_NCCL(n->GroupEnd());
    return 0;
}

// new boundary planes of state array `qi` -> the neighbours' ghost planes.  pbc: apply the path A periodic copy
// Q[.,.,0] <- Q[.,.,Nz-1] (the plane the last slab sends up lands in slab 0's OWNED plane 0).
static int exchange_ghosts(Engine* e, int qi, bool pbc, bool side) {
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        const long long vs = (long long)(sl.nzl + 2) * e->plane;
        cudaStream_t st = stream_of(sl, side);
        if (int rc = copy8(e, sl.stage(S_SEND_UP, e->pl8), e->plane, sl.Q[qi] + (size_t)sl.nzl * e->plane, vs, st, side)) return rc;   // plane k1-1
        if (int rc = copy8(e, sl.stage(S_SEND_DOWN, e->pl8), e->plane, sl.Q[qi] + e->plane, vs, st, side)) return rc;                  // plane k0
    }
    if (int rc = ring_exchange(e, S_SEND_UP, S_SEND_DOWN, S_RECV_LO, S_RECV_HI, false, 0, side)) return rc;
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        const long long vs = (long long)(sl.nzl + 2) * e->plane;
        cudaStream_t st = stream_of(sl, side);
        const float *rlo = sl.stage(recv_stage(S_RECV_LO, sl.seq), e->pl8), *rhi = sl.stage(recv_stage(S_RECV_HI, sl.seq), e->pl8);
        if (sl.rank > 0) {
            if (int rc = copy8(e, sl.Q[qi], vs, rlo, e->plane, st, side)) return rc;
        } else if (pbc) {
            if (int rc = copy8(e, sl.Q[qi] + e->plane, vs, rlo, e->plane, st, side)) return rc;
        }
        if (sl.rank < e->world - 1)
            if (int rc = copy8(e, sl.Q[qi] + (size_t)(sl.nzl + 1) * e->plane, vs, rhi, e->plane, st, side)) return rc;
    }
    return 0;
}

// predictor planes of state array `qi` into plane set `which`:
//   up-going   Qint(k1-1) -> the slab above (its lo); the last slab sends Qint(Nz-2) == Qint(-1), slab 0's wrap plane
//   down-going Qint(k0)   -> the slab below (its hi); slab 0 sends Qint(0) == Qint(Nz-1)
static int exchange_qint(Engine* e, int qi, int which, bool side) {
    for (SlabDev& sl : e->s) {
        ENG_DEV(sl);
        cudaStream_t st = stream_of(sl, side);
        const int up_plane = sl.rank < e->world - 1 ? sl.k0 + sl.nzl - 1 : e->Nz - 2;
        const int rows = side ? 1 : 8;   // under the interior launch: one-warp blocks (80 registers) fit beside a marching block
        if (int rc = imhd_qint_plane_rows(sl.Q[qi], sl.set(which, P_UP, e->pl8), up_plane, &sl.desc, rows, st)) return rc;
        if (int rc = imhd_qint_plane_rows(sl.Q[qi], sl.set(which, P_DOWN, e->pl8), sl.k0, &sl.desc, rows, st)) return rc;
    }
    return ring_exchange(e, P_UP, P_DOWN, P_LO, P_HI, true, which, side);
}

int eng_refresh_ghosts(Engine* e) { return exchange_ghosts(e, e->cur, false, false); }

int eng_prime(Engine* e, int path, float D, float dt) {
    if (path != IMHD_PATH_A && path != IMHD_PATH_B) { set_error("imhd_ctx_prime: bad path %d", path); return IMHD_E_INVALID; }
    if (!(e->dx > 0.f) || !(e->dy > 0.f) || !(e->dz > 0.f)) {
        set_error("imhd_ctx_prime: grid spacing unset (call imhd_ctx_init_grids or imhd_ctx_set_spacing)");
        return IMHD_E_STATE;
    }
    e->path = path; e->D = D; e->dt = dt; e->corner_e = 0.f;
    NcclApi* n = e->nccl;
    if (path == IMHD_PATH_A) {
        // no_diffusion.cu:174-177 on the global state: wall values on j = 0, Ny-1 for k in [1, Nz-2], then PBCs
        for (SlabDev& sl : e->s) {
            ENG_DEV(sl);
            const int ka = sl.rank == 0 ? 2 : 1, kb = sl.rank == e->world - 1 ? sl.nzl : sl.nzl + 1;  // array planes
            if (int rc = imhd_wall_leftright_planes(sl.Q[e->cur], e->Nx, e->Ny, sl.nzl + 2, ka, kb, sl.main)) return rc;
        }
        if (int rc = exchange_ghosts(e, e->cur, true, false)) return rc;
    } else {
        // the column (Nx-1,Ny-1) holds the wall value of (Nx-1,Ny-1,0) at k = 0 and k = Nz-1 from step 1 on (B-8): the
        // slab that owns plane 0 derives it, every slab receives it
        for (SlabDev& sl : e->s) {
            if (sl.rank != 0) continue;
            ENG_DEV(sl);
            float e0 = 0.f;
            const size_t l = (size_t)(e->Nx - 1) * e->Ny + (e->Ny - 1);
            const size_t vs = (size_t)(sl.nzl + 2) * e->plane;
            IMHD_CUDA(cudaMemcpyAsync(&e0, sl.Q[e->cur] + 7 * vs + e->plane + l, sizeof(float), cudaMemcpyDeviceToHost, sl.main));
            IMHD_CUDA(cudaStreamSynchronize(sl.main));
            const float ce = imhd_wall_energy_fixed_point(e0, e->Nx);
            IMHD_CUDA(cudaMemcpyAsync(sl.scratch, &ce, sizeof(float), cudaMemcpyHostToDevice, sl.main));
        }
        IMHD_NCCL(n->GroupStart());
        for (SlabDev& sl : e->s) {
            cudaSetDevice(sl.dev);
            IMHD_NCCL(n->Broadcast(sl.scratch, sl.scratch, 1, ncclFloat, 0, sl.comm, sl.main));
        }
        IMHD_NCCL(n->GroupEnd());
        for (SlabDev& sl : e->s) {
            ENG_DEV(sl);
            IMHD_CUDA(cudaMemcpyAsync(&e->corner_e, sl.scratch, sizeof(float), cudaMemcpyDeviceToHost, sl.main));
            IMHD_CUDA(cudaStreamSynchronize(sl.main));
        }
        if (int rc = exchange_ghosts(e, e->cur, false, false)) return rc;
    }
    fill_desc(e);
    e->q_ready = false;
    e->primed = true;
    return 0;
}

static void planes_of(Engine* e, SlabDev& sl, const float** lo, const float** hi, const float** wrap) {
    if (sl.rank == 0) {  // received Qint(Nz-2) from the last slab; the plane below plane 1 is this slab's own Qint(0)
        *lo = sl.set(e->qcur, P_DOWN, e->pl8); *hi = sl.set(e->qcur, P_HI, e->pl8); *wrap = sl.set(e->qcur, P_LO, e->pl8);
    } else {
        *lo = sl.set(e->qcur, P_LO, e->pl8); *hi = sl.set(e->qcur, P_HI, e->pl8); *wrap = nullptr;
    }
}

int eng_step(Engine* e, int nsteps) {
    if (!e->primed) { set_error("imhd_ctx_step before imhd_ctx_prime"); return IMHD_E_STATE; }
    const int E = g_edge;
    for (int it = 0; it < nsteps; ++it) {
        const int in = e->cur, out = 1 - e->cur;
        if (!e->q_ready) {
            if (int rc = exchange_qint(e, in, e->qcur, false)) return rc;
            e->q_ready = true;
        }
        for (SlabDev& sl : e->s) {
            ENG_DEV(sl);
            const float *lo, *hi, *wrap;
            planes_of(e, sl, &lo, &hi, &wrap);
            const int k0 = sl.k0, k1 = sl.k0 + sl.nzl;
            if (!e->overlap) {
                if (int rc = imhd_step_fused_planes(sl.Q[in], sl.Q[out], lo, hi, wrap, &sl.desc, k0, k1, sl.main)) return rc;
                continue;
            }
            // both ends hold E planes of the marching kernel (plane 0, and plane Nz-1 of path B, have kernels of their own and
            // do not count): equal ranges go out as ONE launch
            const int e0 = k0 + E + (k0 == 0 ? 1 : 0), e1 = k1 - E - (k1 == e->Nz && e->path == IMHD_PATH_B ? 1 : 0);
            // The end launch and the interior launch are independent (both read the old state and write disjoint planes):
            // the ends go to a stream of the highest priority forked off the main stream, so the two launches are runnable
            // together, the block scheduler places the end blocks first and fills every SM they leave with interior blocks
            // -- no drain / fill between the launches, no partial last wave of end blocks.
            cudaStream_t es = e->ends_concurrent ? sl.ends : sl.main;
            if (e->ends_concurrent) {
                IMHD_CUDA(cudaEventRecord(sl.ev_fork, sl.main));
                IMHD_CUDA(cudaStreamWaitEvent(sl.ends, sl.ev_fork, 0));
            }
            if (int rc = imhd_step_fused_ends(sl.Q[in], sl.Q[out], lo, hi, wrap, &sl.desc, k0, e0, e1, k1, es)) return rc;
            IMHD_CUDA(cudaEventRecord(sl.ev_edges, es));
            if (int rc = imhd_step_fused_planes(sl.Q[in], sl.Q[out], lo, hi, wrap, &sl.desc, e0, e1, sl.main)) return rc;
            IMHD_CUDA(cudaStreamWaitEvent(sl.side, sl.ev_edges, 0));   // (the main stream follows the side stream below)
        }
        if (!e->overlap) {
            if (int rc = exchange_ghosts(e, out, e->path == IMHD_PATH_A, false)) return rc;
            e->cur = out;
            e->q_ready = false;
            continue;
        }
        // under the interior launches: halo of the new state, then the predictor planes of the next step
        if (int rc = exchange_ghosts(e, out, e->path == IMHD_PATH_A, true)) return rc;
        e->qcur = 1 - e->qcur;
        if (int rc = exchange_qint(e, out, e->qcur, true)) return rc;
        for (SlabDev& sl : e->s) {
            ENG_DEV(sl);
            IMHD_CUDA(cudaEventRecord(sl.ev_comm, sl.side));
            IMHD_CUDA(cudaStreamWaitEvent(sl.main, sl.ev_comm, 0));
        }
        e->cur = out;
    }
    return 0;
}

// CFL scan over the whole domain: per-slab scans combined as the reference's raster scan would see them (the first
// cell in k, i, j order wins ties; src/on-device/utils/compute_stability.cpp:139-141)
int eng_stability(Engine* e, float dt, imhd_stability* out) {
    if (!out) { set_error("imhd_ctx_stability: null result"); return IMHD_E_INVALID; }
    NcclApi* n = e->nccl;
    const int nl = (int)e->s.size();
    std::vector<double> rows((size_t)e->world * 8, 0.0);
    for (int q = 0; q < nl; ++q) {
        SlabDev& sl = e->s[q];
        ENG_DEV(sl);
        imhd_slab d = sl.desc;
        d.Nx = e->Nx; d.Ny = e->Ny; d.Nz = e->Nz; d.k0 = sl.k0; d.nzl = sl.nzl; d.ghosts = 1;
        d.dt = dt; d.dx = e->dx; d.dy = e->dy; d.dz = e->dz;
        imhd_stability r;
        if (int rc = imhd_stability_scan(sl.Q[e->cur], &d, &r, sl.main)) return rc;
        double row[8] = {(double)r.max_lhs, (double)r.i, (double)r.j, (double)r.k, (double)r.violations, 0, 0, 0};
        IMHD_CUDA(cudaMemcpyAsync(sl.scratch, row, sizeof(row), cudaMemcpyHostToDevice, sl.main));
        IMHD_CUDA(cudaStreamSynchronize(sl.main));
    }
    std::vector<double*> gathered(nl, nullptr);
    for (int q = 0; q < nl; ++q) {
        ENG_DEV(e->s[q]);
        IMHD_CUDA(cudaMalloc(&gathered[q], sizeof(double) * 8 * e->world));
    }
    IMHD_NCCL(n->GroupStart());
    for (int q = 0; q < nl; ++q) {
        cudaSetDevice(e->s[q].dev);
        IMHD_NCCL(n->AllGather(e->s[q].scratch, gathered[q], 8, ncclFloat64, e->s[q].comm, e->s[q].main));
    }
    IMHD_NCCL(n->GroupEnd());
    for (int q = 0; q < nl; ++q) {
        ENG_DEV(e->s[q]);
        if (q == 0) IMHD_CUDA(cudaMemcpyAsync(rows.data(), gathered[q], sizeof(double) * 8 * e->world, cudaMemcpyDeviceToHost, e->s[q].main));
        IMHD_CUDA(cudaStreamSynchronize(e->s[q].main));
        cudaFree(gathered[q]);
    }
    int best = 0;
    unsigned long long viol = 0;
    for (int r = 0; r < e->world; ++r) {
        if (rows[(size_t)r * 8] > rows[(size_t)best * 8]) best = r;
        viol += (unsigned long long)rows[(size_t)r * 8 + 4];
    }
    out->max_lhs = (float)rows[(size_t)best * 8];
    out->i = (int)rows[(size_t)best * 8 + 1]; out->j = (int)rows[(size_t)best * 8 + 2]; out->k = (int)rows[(size_t)best * 8 + 3];
    out->violations = viol;
    out->dt_new = out->max_lhs > 0.f ? 0.1f * dt / out->max_lhs : 0.f;
    return 0;
}

}  // namespace imhd

// ---- public entry points that do not need a context ---------------------------------------------------------------
extern "C" int imhd_nccl_unique_id(void* id_out, int bytes) {
    using namespace imhd;
    if (!id_out || bytes < (int)sizeof(ncclUniqueId)) { set_error("imhd_nccl_unique_id: need a %d-byte buffer", (int)sizeof(ncclUniqueId)); return IMHD_E_INVALID; }
    NcclApi* n = nccl_api();
    if (!n) return IMHD_E_STATE;
    ncclUniqueId id;
    IMHD_NCCL(n->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

extern "C" void imhd_set_edge_planes(int planes) { imhd::eng_set_edge(planes); }
