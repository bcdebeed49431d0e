/*
 * TEST INFRASTRUCTURE ONLY -- not product code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only
 * as the checker.  The product (imhd-cuda_b200/) never links or calls it.
 *
 * CPU restatement of the reference's Lax-Wendroff hot path (russellmatt66/imhd-CUDA,
 * lib/on-device/).  Every function cites the reference file:line it follows.  The
 * restatement keeps the reference's fp32/fp64 rounding points (C float/double promotion
 * rules are the ones the CUDA sources compile under) and its live quirks (SURVEY.md
 * Appendix B), but evaluates each flux once per cell instead of once per use.
 *
 * PARITY PINNING: the reference ships no time-stepped golden vector (its only one,
 * debug/data/rhovz/var_0.csv, pins the initial condition).  This restatement is pinned
 * against the reference ITSELF run here: oracle/_ref/libimhd_ref_cpu.so is the
 * reference's unmodified kernel sources compiled for the host (oracle/Makefile `ref`),
 * and tests/test_oracle_golden.py requires BIT-EXACT agreement of every entry point
 * below with it, plus agreement with the committed fixtures under tests/golden/ that
 * were generated from it (tests/golden/make_golden.py).
 *
 * Layout (lib/on-device/kernels_od.cu:11,16): l = k*Nx*Ny + i*Ny + j, variable v at
 * l + v*Nx*Ny*Nz, v = rho, rhovx, rhovy, rhovz, Bx, By, Bz, e.  64-bit offsets here
 * (the reference's int arithmetic overflows above 306 M cells).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- tiny pthread parallel-for over a z range (no OpenMP dependency) ---- */
typedef void (*range_fn)(int k0, int k1, void* ctx);
typedef struct { range_fn fn; void* ctx; int k0, k1; } job_t;
static void* job_main(void* p) { job_t* j = (job_t*)p; j->fn(j->k0, j->k1, j->ctx); return NULL; }
static int n_threads(void) {
    const char* e = getenv("IMHD_ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}
int oracle_num_threads(void) { return n_threads(); }
static void parallel_range(int k0, int k1, range_fn fn, void* ctx) {
    int T = n_threads();
    if (T > k1 - k0) T = k1 - k0;
    if (T <= 1) { if (k1 > k0) fn(k0, k1, ctx); return; }
    pthread_t th[256]; job_t jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].k0 = k0 + (int)((long long)(k1 - k0) * t / T);
        jobs[t].k1 = k0 + (int)((long long)(k1 - k0) * (t + 1) / T);
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
}

#define GAMMA (5.0 / 3.0) /* include/on-device/kernels_od.cuh:16 (a double) */

enum { RHO = 0, MX = 1, MY = 2, MZ = 3, BX = 4, BY = 5, BZ = 6, EN = 7 };
enum { PATH_A = 0, PATH_B = 1 };

typedef struct {
    int Nx, Ny, Nz;
    size_t cube;
} dims_t;

static inline size_t IDX(const dims_t* d, int i, int j, int k) {
    return (size_t)k * d->Nx * d->Ny + (size_t)i * d->Ny + j;
}

static inline double sq(float x) { return (double)x * (double)x; } /* pow(float,2): exact in double */

static inline void load8(const float* A, const dims_t* d, size_t l, float U[8]) {
    for (int v = 0; v < 8; ++v) U[v] = A[l + (size_t)v * d->cube];
}

/* ---- thermo helpers: helper_functions.cu:7-26 (value args) == :29-62 (memory-reading) ---- */
static inline float h_Bsq(float bx, float by, float bz) { return (float)(sq(bx) + sq(by) + sq(bz)); }
static inline float h_KE(float rho, float mx, float my, float mz) { /* no 1/2: SURVEY.md B-1 */
    return (float)((1.0 / rho) * (sq(mx) + sq(my) + sq(mz)));
}
static inline float h_p(float e, float Bsq, float KE) { /* (e - KE) is a float subtraction */
    return (float)((GAMMA - 1.0) * ((e - KE) - Bsq / 2.0));
}
static inline float h_Bdotu(float rho, float mx, float my, float mz, float bx, float by, float bz) {
    return (float)((1.0 / rho) * (mx * bx + my * by + mz * bz)); /* inner sum is fp32 */
}

/* ---- flux tensor, INDEXED family: kernels_od_fluxes.cu:112-275 (predictor, BoundaryConditions) ---- */
typedef struct { float F[8], G[8], H[8]; } flux_t;

static void flux_indexed(const float U[8], flux_t* f) {
    const float rho = U[RHO], mx = U[MX], my = U[MY], mz = U[MZ], bx = U[BX], by = U[BY], bz = U[BZ], e = U[EN];
    const double inv = 1.0 / rho;
    const float Bsq = h_Bsq(bx, by, bz);
    const float ke = h_KE(rho, mx, my, mz);
    const float p = h_p(e, Bsq, ke);
    const float Bdotu = h_Bdotu(rho, mx, my, mz, bx, by, bz);
    f->F[RHO] = mx; f->G[RHO] = my; f->H[RHO] = mz;                    /* :112-126 */
    f->F[MX] = (float)(inv * sq(mx) - sq(bx) + p + Bsq / 2.0);        /* :129-137 */
    f->G[MX] = (float)(inv * mx * my - bx * by);                       /* :138-143 */
    f->H[MX] = (float)(inv * mx * mz - bx * bz);                       /* :144-149 */
    f->F[MY] = f->G[MX];                                                /* :152-158 */
    f->G[MY] = (float)(inv * sq(my) - sq(by) + p + Bsq / 2.0);        /* :159-167 */
    f->H[MY] = (float)(inv * my * mz - by * bz);                       /* :168-173 */
    f->F[MZ] = f->H[MX]; f->G[MZ] = f->H[MY];                          /* :176-183 */
    f->H[MZ] = (float)(inv * sq(mz) - sq(bz) + p + Bsq / 2.0);        /* :184-192 */
    f->F[BX] = 0.0f;                                                    /* :195-198 */
    f->G[BX] = (float)(inv * mx * by - bx * my);                       /* :199-204 (B-2) */
    f->H[BX] = (float)(inv * mx * bz - bx * mz);                       /* :205-210 */
    f->F[BY] = (float)(-1.0 * f->G[BX]);                               /* :213-216 */
    f->G[BY] = 0.0f;
    f->H[BY] = (float)(inv * my * bz - by * mz);                       /* :221-226 */
    f->F[BZ] = (float)(-1.0 * f->H[BX]);                               /* :229-236 */
    f->G[BZ] = (float)(-1.0 * f->H[BY]);
    f->H[BZ] = 0.0f;
    f->F[EN] = e + p + Bsq * (mx / rho) - Bdotu * bx;                  /* :243-275 all fp32 (B-3) */
    f->G[EN] = e + p + Bsq * (my / rho) - Bdotu * by;
    f->H[EN] = e + p + Bsq * (mz / rho) - Bdotu * bz;
}

/* ---- flux tensor, LOCAL family: kernels_od_fluxes.cu:8-104 (corrector) ---- */
static inline float lf_mom_diag(float rho, float m, float b, float p, float Bsq) { /* :12-14,49-51,82-84 */
    return (float)(sq(m) / rho - sq(b) + p + 0.5 * Bsq);
}
static inline float lf_mom_off(float rho, float ma, float mb, float ba, float bb) { /* :16-22,45-47,... fp32 */
    return (ma * mb) / rho - ba * bb;
}
/* d-direction flux of B_c: (m_c/rho)*B_d - (m_d/rho)*B_c, fp32 (:28-34,57-67,90-96) */
static inline float lf_ind(float rho, float m_c, float m_d, float b_c, float b_d) {
    return (m_c / rho) * b_d - (m_d / rho) * b_c;
}
static inline float lf_en(float rho, float m, float b, float e, float p, float Bsq, float Bdotu) { /* :36-38,69-71,102-104 */
    return (float)((e + p + 0.5 * Bsq) * (m / rho) - Bdotu * b);
}

/* ---- diffusion: diffusion.cu:8-19 (local) == :42-73 (indexed) ---- */
static inline float num_diff(float q, float qip1, float qjp1, float qkp1, float qim1, float qjm1, float qkm1,
                             float D, float dx, float dy, float dz) {
    return (float)(D * ((1.0 / sq(dx)) * (qip1 - 2.0 * q + qim1) + (1.0 / sq(dy)) * (qjp1 - 2.0 * q + qjm1) +
                        (1.0 / sq(dz)) * (qkp1 - 2.0 * q + qkm1)));
}

/* =====================================================================================
 * Predictor.  Writes EVERY cell of Qint exactly as the reference's kernel sequence
 * leaves it:
 *   path A: ComputeIntermediateVariablesNoDiff (kernels_od_intvar.cu:51-74) then
 *           QintBdry{Front,LeftRight,TopBottom,FrontBottom,FrontRight,BottomRight}NoDiff
 *           and QintBdryPBCs (kernels_intvarbcs.cu:360-558)
 *   path B: ComputeIntermediateVariablesStride (kernels_od_intvar.cu:113-153) then
 *           ComputeIntermediateVariablesBoundary (kernels_intvarbcs.cu:177-356)
 * Cell classes (identical for A and B): generic int* (kernels_od_intvar.cu:1160-1253)
 * for i<=Nx-2, j<=Ny-2; "Right" (j=Ny-1) / "Bottom" (i=Nx-1) / "BottomRight" variants
 * with the outward flux zeroed (kernels_intvarbcs.cu:560-738,1024-1110); the k=0 edge
 * lines use FrontRight / FrontBottom with their typos (:840-1018);
 * plane Nz-1 is a copy of plane 0 (:383-398).  Path B adds dt*D*lap(Q) on [1,N-2]^3 only.
 * ===================================================================================== */
static void flux_plane(const float* Q, const dims_t* d, int k, flux_t* out) {
    for (int i = 0; i < d->Nx; ++i)
        for (int j = 0; j < d->Ny; ++j) {
            float U[8];
            load8(Q, d, IDX(d, i, j, k), U);
            flux_indexed(U, &out[(size_t)i * d->Ny + j]);
        }
}

typedef struct {
    const float* Q; float* Qint; int path; float D, dt, dx, dy, dz; dims_t d;
} pred_ctx;

static void predictor_range(int ka, int kb, void* vctx) {
    const pred_ctx* a = (const pred_ctx*)vctx;
    const dims_t* d = &a->d;
    const int Nx = d->Nx, Ny = d->Ny, Nz = d->Nz, path = a->path;
    const float *Q = a->Q, D = a->D, dt = a->dt, dx = a->dx, dy = a->dy, dz = a->dz;
    float* Qint = a->Qint;
    const size_t plane = (size_t)Nx * Ny;
    flux_t* f0 = (flux_t*)malloc(plane * sizeof(flux_t)); /* plane k   */
    flux_t* f1 = (flux_t*)malloc(plane * sizeof(flux_t)); /* plane k+1 */
    for (int k = ka; k < kb; ++k) {
        if (k == ka) flux_plane(Q, d, k, f0);
        else { flux_t* t = f0; f0 = f1; f1 = t; }
        flux_plane(Q, d, k + 1, f1);
        for (int i = 0; i < Nx; ++i)
            for (int j = 0; j < Ny; ++j) {
                const size_t l = IDX(d, i, j, k), c = (size_t)i * Ny + j;
                const int bottom = (i == Nx - 1), right = (j == Ny - 1);
                const int frontright = (k == 0 && right && !bottom);
                const int frontbottom = (k == 0 && bottom && !right);
                for (int v = 0; v < 8; ++v) {
                    float dF, dG, dH, base = Q[l + (size_t)v * d->cube];
                    dF = bottom ? -f0[c].F[v] : f0[c + Ny].F[v] - f0[c].F[v];
                    dG = right ? -f0[c].G[v] : f0[c + 1].G[v] - f0[c].G[v];
                    dH = f1[c].H[v] - f0[c].H[v];
                    /* live typos, SURVEY.md B-15 */
                    if (frontright && v == EN) dF = f0[c].F[v] - f0[c].F[v];             /* kernels_intvarbcs.cu:923 */
                    if (frontbottom && (v == MZ || v == EN)) dH = f1[c].H[v] - f0[c].G[v]; /* :973, :1017 */
                    if (frontbottom && v == BZ) dG = f0[c + 1].G[v] - f0[c + 1].G[v];     /* :1005 */
                    if (bottom && right && v == MZ) base = Q[l + (size_t)MX * d->cube];   /* :1062 */
                    float r = base - (dt / dx) * dF - (dt / dy) * dG - (dt / dz) * dH;
                    if (path == PATH_B && i > 0 && i < Nx - 1 && j > 0 && j < Ny - 1 && k > 0) {
                        const float* q = Q + (size_t)v * d->cube; /* kernels_od_intvar.cu:130-145 */
                        r = r + dt * num_diff(q[l], q[l + Ny], q[l + 1], q[l + plane], q[l - Ny], q[l - 1],
                                              q[l - plane], D, dx, dy, dz);
                    }
                    Qint[l + (size_t)v * d->cube] = r;
                }
            }
    }
    (void)Nz;
    free(f0);
    free(f1);
}

void oracle_predictor(const float* Q, float* Qint, int path, float D, float dt, float dx, float dy, float dz,
                      int Nx, int Ny, int Nz) {
    pred_ctx a = {Q, Qint, path, D, dt, dx, dy, dz, {Nx, Ny, Nz, (size_t)Nx * Ny * Nz}};
    const size_t plane = (size_t)Nx * Ny;
    parallel_range(0, Nz - 1, predictor_range, &a);
    for (int v = 0; v < 8; ++v) /* QintBdryPBCs, kernels_intvarbcs.cu:392-396 / :348-352 */
        memcpy(Qint + (size_t)v * a.d.cube + (size_t)(Nz - 1) * plane, Qint + (size_t)v * a.d.cube, plane * sizeof(float));
}

/* =====================================================================================
 * Corrector, one cell (kernels_od.cu:378-522 / :120-345 and LaxWendroffAdv*Local :1206-1332).
 * c = Qint(i,j,k), xi = Qint(i-1,j,k), yj = Qint(i,j-1,k), zk = Qint(i,j,k-1).
 * ===================================================================================== */
static void corrector_cell(const float q[8], const float c[8], const float xi[8], const float yj[8],
                           const float zk[8], int path, float dt, float dx, float dy, float dz, float out[8]) {
    const float KEc = h_KE(c[RHO], c[MX], c[MY], c[MZ]), KEi = h_KE(xi[RHO], xi[MX], xi[MY], xi[MZ]);
    const float KEj = h_KE(yj[RHO], yj[MX], yj[MY], yj[MZ]), KEk = h_KE(zk[RHO], zk[MX], zk[MY], zk[MZ]);
    const float Bc = h_Bsq(c[BX], c[BY], c[BZ]), Bi = h_Bsq(xi[BX], xi[BY], xi[BZ]), Bj = h_Bsq(yj[BX], yj[BY], yj[BZ]);
    const float Bk = h_Bsq(xi[BX], yj[BY], zk[BZ]); /* mixed neighbours: kernels_od.cu:212,431 (B-4) */
    const float pc = h_p(c[EN], Bc, KEc), pi = h_p(xi[EN], Bi, KEi), pj = h_p(yj[EN], Bj, KEj), pk = h_p(zk[EN], Bk, KEk);
    const float Dc = h_Bdotu(c[RHO], c[MX], c[MY], c[MZ], c[BX], c[BY], c[BZ]);
    const float Di = h_Bdotu(xi[RHO], xi[MX], xi[MY], xi[MZ], xi[BX], xi[BY], xi[BZ]);
    const float Dj = h_Bdotu(yj[RHO], yj[MX], yj[MY], yj[MZ], yj[BX], yj[BY], yj[BZ]);
    const float Dk = h_Bdotu(zk[RHO], zk[MX], yj[MY], zk[MZ], zk[BX], zk[BY], zk[BZ]); /* :232,441 (B-5) */
    const float tx = dt / dx, ty = dt / dy, tz = dt / dz;
    float dF[8], dG[8], dH[8];
    /* rho: kernels_od.cu:1206-1217; path B passes rho_int_im1 for rhovx_int_im1 (:236-240, B-6) */
    dF[RHO] = c[MX] - (path == PATH_B ? xi[RHO] : xi[MX]);
    dG[RHO] = c[MY] - yj[MY];
    dH[RHO] = c[MZ] - zk[MZ];
    /* rhovx :1219-1234 */
    dF[MX] = lf_mom_diag(c[RHO], c[MX], c[BX], pc, Bc) - lf_mom_diag(xi[RHO], xi[MX], xi[BX], pi, Bi);
    dG[MX] = lf_mom_off(c[RHO], c[MX], c[MY], c[BX], c[BY]) - lf_mom_off(yj[RHO], yj[MX], yj[MY], yj[BX], yj[BY]);
    dH[MX] = lf_mom_off(c[RHO], c[MX], c[MZ], c[BX], c[BZ]) - lf_mom_off(zk[RHO], zk[MX], zk[MZ], zk[BX], zk[BZ]);
    /* rhovy :1236-1251 */
    dF[MY] = lf_mom_off(c[RHO], c[MX], c[MY], c[BX], c[BY]) - lf_mom_off(xi[RHO], xi[MX], xi[MY], xi[BX], xi[BY]);
    dG[MY] = lf_mom_diag(c[RHO], c[MY], c[BY], pc, Bc) - lf_mom_diag(yj[RHO], yj[MY], yj[BY], pj, Bj);
    dH[MY] = lf_mom_off(c[RHO], c[MY], c[MZ], c[BY], c[BZ]) - lf_mom_off(zk[RHO], zk[MY], zk[MZ], zk[BY], zk[BZ]);
    /* rhovz :1253-1268 */
    dF[MZ] = lf_mom_off(c[RHO], c[MX], c[MZ], c[BX], c[BZ]) - lf_mom_off(xi[RHO], xi[MX], xi[MZ], xi[BX], xi[BZ]);
    dG[MZ] = lf_mom_off(c[RHO], c[MY], c[MZ], c[BY], c[BZ]) - lf_mom_off(yj[RHO], yj[MY], yj[MZ], yj[BY], yj[BZ]);
    dH[MZ] = lf_mom_diag(c[RHO], c[MZ], c[BZ], pc, Bc) - lf_mom_diag(zk[RHO], zk[MZ], zk[BZ], pk, Bk);
    /* Bx :1270-1283 */
    dF[BX] = 0.0f - 0.0f;
    dG[BX] = lf_ind(c[RHO], c[MX], c[MY], c[BX], c[BY]) - lf_ind(yj[RHO], yj[MX], yj[MY], yj[BX], yj[BY]);
    dH[BX] = lf_ind(c[RHO], c[MX], c[MZ], c[BX], c[BZ]) - lf_ind(zk[RHO], zk[MX], zk[MZ], zk[BX], zk[BZ]);
    /* By :1285-1298 */
    dF[BY] = lf_ind(c[RHO], c[MY], c[MX], c[BY], c[BX]) - lf_ind(xi[RHO], xi[MY], xi[MX], xi[BY], xi[BX]);
    dG[BY] = 0.0f - 0.0f;
    dH[BY] = lf_ind(c[RHO], c[MY], c[MZ], c[BY], c[BZ]) - lf_ind(zk[RHO], zk[MY], zk[MZ], zk[BY], zk[BZ]);
    /* Bz :1300-1313 */
    dF[BZ] = lf_ind(c[RHO], c[MZ], c[MX], c[BZ], c[BX]) - lf_ind(xi[RHO], xi[MZ], xi[MX], xi[BZ], xi[BX]);
    dG[BZ] = lf_ind(c[RHO], c[MZ], c[MY], c[BZ], c[BY]) - lf_ind(yj[RHO], yj[MZ], yj[MY], yj[BZ], yj[BY]);
    dH[BZ] = 0.0f - 0.0f;
    /* e :1315-1332 */
    dF[EN] = lf_en(c[RHO], c[MX], c[BX], c[EN], pc, Bc, Dc) - lf_en(xi[RHO], xi[MX], xi[BX], xi[EN], pi, Bi, Di);
    dG[EN] = lf_en(c[RHO], c[MY], c[BY], c[EN], pc, Bc, Dc) - lf_en(yj[RHO], yj[MY], yj[BY], yj[EN], pj, Bj, Dj);
    dH[EN] = lf_en(c[RHO], c[MZ], c[BZ], c[EN], pc, Bc, Dc) - lf_en(zk[RHO], zk[MZ], zk[BZ], zk[EN], pk, Bk, Dk);
    for (int v = 0; v < 8; ++v) /* whole expression in fp64, rounded once (A.6) */
        out[v] = (float)(0.5 * (q[v] + c[v]) - 0.5 * tx * dF[v] - 0.5 * ty * dG[v] - 0.5 * tz * dH[v]);
}

typedef struct {
    float* Q; const float* Qint; int path; float D, dt, dx, dy, dz; dims_t d;
} corr_ctx;

static void corrector_range(int ka, int kb, void* vctx) {
    const corr_ctx* a = (const corr_ctx*)vctx;
    const dims_t* d = &a->d;
    const int Ny = d->Ny, path = a->path;
    const int iend = path == PATH_A ? d->Nx : d->Nx - 1, jend = path == PATH_A ? d->Ny : d->Ny - 1;
    const size_t plane = (size_t)d->Nx * d->Ny;
    float* Q = a->Q;
    const float* Qint = a->Qint;
    for (int k = ka; k < kb; ++k)
        for (int i = 1; i < iend; ++i)
            for (int j = 1; j < jend; ++j) {
                const size_t l = IDX(d, i, j, k);
                float q[8], c[8], xi[8], yj[8], zk[8], out[8];
                load8(Q, d, l, q); load8(Qint, d, l, c);
                load8(Qint, d, l - Ny, xi); load8(Qint, d, l - 1, yj); load8(Qint, d, l - plane, zk);
                corrector_cell(q, c, xi, yj, zk, path, a->dt, a->dx, a->dy, a->dz, out);
                for (int v = 0; v < 8; ++v) {
                    if (path == PATH_B) { /* + dt * numericalDiffusionLocal(Qint), kernels_od.cu:241-345 */
                        const float* qi = Qint + (size_t)v * d->cube;
                        out[v] = out[v] + a->dt * num_diff(c[v], qi[l + Ny], qi[l + 1], qi[l + plane], xi[v], yj[v],
                                                           zk[v], a->D, a->dx, a->dy, a->dz);
                    }
                    Q[l + (size_t)v * d->cube] = out[v];
                }
            }
}

/* Path A: FluidAdvanceLocalNoDiff (kernels_od.cu:353-525), cells i,j,k >= 1 incl. far faces (B-14),
 * in place (reads only its own Q cell). */
void oracle_corrector_nodiff(float* Q, const float* Qint, float dt, float dx, float dy, float dz, int Nx, int Ny, int Nz) {
    corr_ctx a = {Q, Qint, PATH_A, 0.0f, dt, dx, dy, dz, {Nx, Ny, Nz, (size_t)Nx * Ny * Nz}};
    parallel_range(1, Nz, corrector_range, &a);
}

void oracle_pbcs(float* Q, int Nx, int Ny, int Nz) { /* kernels_fluidbcs.cu:498-510 */
    const size_t plane = (size_t)Nx * Ny, cube = plane * Nz;
    for (int v = 0; v < 8; ++v) memcpy(Q + v * cube, Q + v * cube + (size_t)(Nz - 1) * plane, plane * sizeof(float));
}

/* wall energy e <- p(e,0,0)/(gamma-1): kernels_fluidbcs.cu:173,451,461 (B-12) */
static inline float wall_e(float e) {
    float p = (float)((GAMMA - 1.0) * ((e - 0.0f) - 0.0f / 2.0));
    return (float)(p / (GAMMA - 1.0));
}
static void wall_cell(float* Q, const dims_t* d, size_t l) {
    Q[l] = 1.0f;
    for (int v = 1; v < 7; ++v) Q[l + (size_t)v * d->cube] = 0.0f;
    Q[l + (size_t)EN * d->cube] = wall_e(Q[l + (size_t)EN * d->cube]);
}

/* Path A init only: rigidConductingWallBCsLeftRight (kernels_fluidbcs.cu:436-464): j=0 and j=Ny-1 for
 * all i, 0<k<Nz-1.  rigidConductingWallBCsTopBottom is a no-op under the shipped launch (B-11). */
void oracle_wall_bcs_leftright(float* Q, int Nx, int Ny, int Nz) {
    dims_t dd = {Nx, Ny, Nz, (size_t)Nx * Ny * Nz};
    for (int k = 1; k < Nz - 1; ++k)
        for (int i = 0; i < Nx; ++i) {
            wall_cell(Q, &dd, IDX(&dd, i, 0, k));
            wall_cell(Q, &dd, IDX(&dd, i, Ny - 1, k));
        }
}

/* Path B: FluidAdvanceLocal (kernels_od.cu:82-350): [1,N-2]^3, corrector + dt*D*lap(Qint), in place. */
void oracle_corrector_diff(float* Q, const float* Qint, float D, float dt, float dx, float dy, float dz,
                           int Nx, int Ny, int Nz) {
    corr_ctx a = {Q, Qint, PATH_B, D, dt, dx, dy, dz, {Nx, Ny, Nz, (size_t)Nx * Ny * Nz}};
    parallel_range(1, Nz - 1, corrector_range, &a);
}

/* Path B: BoundaryConditions (kernels_fluidbcs.cu:32-235) with its deterministic single-application
 * meaning (launch z-extent 1):
 *  (1) k=0 face, i in [1,Nx-2], j in [1,Ny-2]: corrector with INDEXED fluxes of Qint, k-1 -> Nz-2 (B-10),
 *      + dt*numericalDiffusionFront(Qint) (diffusion.cu:79-106); whole expression fp64 (:53-116)
 *  (2) wall values at (0,j,0) and (Nx-1,j,0) for all j (:164-188; the j-wall blocks are dead, B-8);
 *      every x-thread of the launch re-applies it, so e goes through wall_e() Nx times
 *  (3) the "PBC" copies only the column (Nx-1,Ny-1): Q[..,Nz-1] <- Q[..,0] (:227-231, B-8) */
void oracle_boundary_conditions(float* Q, const float* Qint, float D, float dt, float dx, float dy, float dz,
                                int Nx, int Ny, int Nz) {
    dims_t dd = {Nx, Ny, Nz, (size_t)Nx * Ny * Nz};
    const dims_t* d = &dd;
    const size_t plane = (size_t)Nx * Ny;
    flux_t* f0 = (flux_t*)malloc(plane * sizeof(flux_t));
    flux_t* fb = (flux_t*)malloc(plane * sizeof(flux_t));
    flux_plane(Qint, d, 0, f0);
    flux_plane(Qint, d, Nz - 2, fb);
    const float tx = dt / dx, ty = dt / dy, tz = dt / dz;
    for (int i = 1; i < Nx - 1; ++i)
        for (int j = 1; j < Ny - 1; ++j) {
            const size_t l = IDX(d, i, j, 0), c = (size_t)i * Ny + j;
            for (int v = 0; v < 8; ++v) {
                const float* qi = Qint + (size_t)v * d->cube;
                float* q = Q + (size_t)v * d->cube;
                float nd = num_diff(qi[l], qi[l + Ny], qi[l + 1], qi[IDX(d, i, j, 1)], qi[l - Ny], qi[l - 1],
                                    qi[IDX(d, i, j, Nz - 2)], D, dx, dy, dz);
                q[l] = (float)(0.5 * (q[l] + qi[l]) - 0.5 * tx * (f0[c].F[v] - f0[c - Ny].F[v]) -
                               0.5 * ty * (f0[c].G[v] - f0[c - 1].G[v]) - 0.5 * tz * (f0[c].H[v] - fb[c].H[v]) +
                               dt * nd);
            }
        }
    free(f0);
    free(fb);
    for (int rep = 0; rep < Nx; ++rep)
        for (int j = 0; j < Ny; ++j) {
            wall_cell(Q, d, IDX(d, 0, j, 0));
            wall_cell(Q, d, IDX(d, Nx - 1, j, 0));
        }
    for (int v = 0; v < 8; ++v)
        Q[IDX(d, Nx - 1, Ny - 1, Nz - 1) + (size_t)v * d->cube] = Q[IDX(d, Nx - 1, Ny - 1, 0) + (size_t)v * d->cube];
}

/* ---- composite drivers: launch order of no_diffusion.cu:174-199,288-311 and main.cu:108-112,200-213 ---- */
void oracle_prime(float* Q, float* Qint, int path, float D, float dt, float dx, float dy, float dz, int Nx, int Ny, int Nz) {
    if (path == PATH_A) {
        oracle_wall_bcs_leftright(Q, Nx, Ny, Nz);
        oracle_pbcs(Q, Nx, Ny, Nz);
    }
    oracle_predictor(Q, Qint, path, D, dt, dx, dy, dz, Nx, Ny, Nz);
}

void oracle_steps(float* Q, float* Qint, int path, int nsteps, float D, float dt, float dx, float dy, float dz,
                  int Nx, int Ny, int Nz) {
    for (int s = 0; s < nsteps; ++s) {
        if (path == PATH_A) {
            oracle_corrector_nodiff(Q, Qint, dt, dx, dy, dz, Nx, Ny, Nz);
            oracle_pbcs(Q, Nx, Ny, Nz);
        } else {
            oracle_corrector_diff(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
            oracle_boundary_conditions(Q, Qint, D, dt, dx, dy, dz, Nx, Ny, Nz);
        }
        oracle_predictor(Q, Qint, path, D, dt, dx, dy, dz, Nx, Ny, Nz);
    }
}

/* ---- grids and initial conditions: initialize_od.cu:26-57, :269-345, :132-205 ---- */
void oracle_init_grids(float* x, float* y, float* z, float x_min, float x_max, float y_min, float y_max,
                       float z_min, float z_max, int Nx, int Ny, int Nz) {
    float dx = (x_max - x_min) / (Nx - 1), dy = (y_max - y_min) / (Ny - 1), dz = (z_max - z_min) / (Nz - 1);
    for (int i = 0; i < Nx; ++i) x[i] = x_min + i * dx;
    for (int j = 0; j < Ny; ++j) y[j] = y_min + j * dy;
    for (int k = 0; k < Nz; ++k) z[k] = z_min + k * dz;
}

static void vacuum_cell(float* Q, size_t l, size_t cube) {
    Q[l] = 0.01f;
    for (int v = 1; v < 8; ++v) Q[l + (size_t)v * cube] = 0.0f;
}

static float total_energy(const float* Q, size_t l, size_t cube, float p) { /* initialize_od.cu:327-334 */
    return (float)((p / (GAMMA - 1.0)) + (sq(Q[l + MX * cube]) + sq(Q[l + MY * cube]) + sq(Q[l + MZ * cube])) / (2.0 * Q[l]) +
                   0.5 * (sq(Q[l + BX * cube]) + sq(Q[l + BY * cube]) + sq(Q[l + BZ * cube])));
}

void oracle_screwpinch_stride(float* Q, float J0, const float* gx, const float* gy, const float* gz,
                              int Nx, int Ny, int Nz) {
    (void)gz;
    dims_t dd = {Nx, Ny, Nz, (size_t)Nx * Ny * Nz};
    const size_t cube = dd.cube;
    const float r_pinch = (float)(0.25 * sqrtf((float)(sq(gx[Nx - 1]) + sq(gy[Ny - 1])))); /* :281 */
    const float Jr = 0.0f, Jphi = 0.0f, Br = 0.0f, B0 = 1.0f;
    for (int k = 0; k < Nz; ++k)
        for (int i = 0; i < Nx; ++i)
            for (int j = 0; j < Ny; ++j) {
                const size_t l = IDX(&dd, i, j, k);
                const float x = gx[i], y = gy[j];
                const float r = sqrtf((float)(sq(x) + sq(y)));
                vacuum_cell(Q, l, cube);
                if (r < r_pinch) {
                    const float Btheta = (float)(0.5 * J0 * r * (1.0 - 0.5 * sq(r) / sq(r_pinch))); /* :313 */
                    const float p = (float)(-0.25 * (sq(J0) / pow(r_pinch, 4)) *
                                            (pow(r, 6) / 6.0 - 0.75 * sq(r_pinch) * pow(r, 4) + pow(r_pinch, 4) * sq(r))); /* :316 */
                    Q[l] = 1.0f;
                    Q[l + MX * cube] = Jr * x - Jphi * y / r;
                    Q[l + MY * cube] = Jr * y + Jphi * x / r;
                    Q[l + MZ * cube] = (float)(J0 * (1 - sq(r) / sq(r_pinch)));
                    Q[l + BX * cube] = Br * x - Btheta * y / r;
                    Q[l + BY * cube] = Br * y + Btheta * x / r;
                    Q[l + BZ * cube] = B0;
                    Q[l + EN * cube] = total_energy(Q, l, cube, p);
                }
            }
}

void oracle_cubic_bennett_vortex_m0(float* Q, float kw, float A, const float* gx, const float* gy, const float* gz,
                                    int Nx, int Ny, int Nz) {
    dims_t dd = {Nx, Ny, Nz, (size_t)Nx * Ny * Nz};
    const size_t cube = dd.cube;
    const float r_pinch = (float)(0.25 * sqrtf((float)(sq(gx[Nx - 1]) + sq(gy[Ny - 1])))); /* :145 */
    const float Br = 0.0f;
    (void)kw;
    for (int k = 0; k < Nz; ++k)
        for (int i = 0; i < Nx; ++i)
            for (int j = 0; j < Ny; ++j) {
                const size_t l = IDX(&dd, i, j, k);
                const float xt = gx[i], yt = gy[j], z = gz[k];
                const float phi = sqrtf((float)(sq(xt) + sq(yt)));
                vacuum_cell(Q, l, cube);
                if (phi < r_pinch) {
                    const float Btheta = (float)(-(1) * (pow(phi, 3) - 3 * sq(phi) - 6 * phi + 6 * (phi + 1) * logf(phi + 1)) /
                                                 (2 * phi * (phi + 1))); /* :177 */
                    const float p = (float)(1 - (pow(phi, 3)) / sq(phi + 1) * (phi - 10)); /* :180 */
                    Q[l] = (float)(1.0 + A * cosf(k * z)); /* loop index k shadows the wavenumber argument (:158,183) */
                    Q[l + MZ * cube] = (float)((1) * sq(phi) / sq(phi + 1));
                    Q[l + BX * cube] = Br * xt - Btheta * yt / phi;
                    Q[l + BY * cube] = Br * yt + Btheta * xt / phi;
                    Q[l + EN * cube] = total_energy(Q, l, cube, p);
                }
            }
}

/* initialize_od.cu:59-130 -- the z-invariant cubic Bennett vortex (commented out in no_diffusion.cu:170) */
void oracle_cubic_bennett_vortex(float* Q, const float* gx, const float* gy, const float* gz, int Nx, int Ny, int Nz) {
    (void)gz;
    dims_t dd = {Nx, Ny, Nz, (size_t)Nx * Ny * Nz};
    const size_t cube = dd.cube;
    const float r_pinch = (float)(0.25 * sqrtf((float)(sq(gx[Nx - 1]) + sq(gy[Ny - 1])))); /* :71 */
    const float Br = 0.0f;
    for (int k = 0; k < Nz; ++k)
        for (int i = 0; i < Nx; ++i)
            for (int j = 0; j < Ny; ++j) {
                const size_t l = IDX(&dd, i, j, k);
                const float xt = gx[i], yt = gy[j];
                const float phi = sqrtf((float)(sq(xt) + sq(yt)));
                vacuum_cell(Q, l, cube);
                if (phi < r_pinch) {
                    const float Btheta = (float)(-(pow(phi, 3) - 3 * sq(phi) - 6 * phi + 6 * (phi + 1) * logf(phi + 1)) /
                                                 (2 * phi * (phi + 1))); /* :101 */
                    const float p = (float)pow(phi, 3);                   /* :103 */
                    Q[l] = 1.0f;
                    Q[l + MZ * cube] = (float)(sq(phi) / sq(phi + 1));
                    Q[l + BX * cube] = Br * xt - Btheta * yt / phi;
                    Q[l + BY * cube] = Br * yt + Btheta * xt / phi;
                    Q[l + EN * cube] = total_energy(Q, l, cube, p);
                }
            }
}

/* initialize_od.cu:347-424 */
void oracle_zpinch(float* Q, float r_max_coeff, const float* gx, const float* gy, const float* gz, int Nx, int Ny, int Nz) {
    (void)gz;
    dims_t dd = {Nx, Ny, Nz, (size_t)Nx * Ny * Nz};
    const size_t cube = dd.cube;
    const float r_pinch = r_max_coeff * sqrtf((float)(sq(gx[Nx - 1]) + sq(gy[Ny - 1]))); /* :359, fp32 product */
    const float Br = 0.0f;
    for (int k = 0; k < Nz; ++k)
        for (int i = 0; i < Nx; ++i)
            for (int j = 0; j < Ny; ++j) {
                const size_t l = IDX(&dd, i, j, k);
                const float x = gx[i], y = gy[j];
                const float r = sqrtf((float)(sq(x) + sq(y)));
                vacuum_cell(Q, l, cube);
                if (r < r_pinch) {
                    const float Btheta = (float)(0.5 * (r + 0.5 * pow(r, 3) / sq(r_pinch))); /* :394 */
                    const float p = (float)(1 + 0.5 * (0.5 * sq(r) / sq(r_pinch) + 0.375 * pow(r, 4) / pow(r_pinch, 4) -
                                                       (1.0 / 12.0) * pow(r, 6) / pow(r_pinch, 6))); /* :398 */
                    Q[l] = 1.0f;
                    Q[l + MZ * cube] = (float)(1 + sq(r) / sq(r_pinch));
                    Q[l + BX * cube] = Br * x - Btheta * y / r;
                    Q[l + BY * cube] = Br * y + Btheta * x / r;
                    Q[l + EN * cube] = total_energy(Q, l, cube, p);
                }
            }
}

/* initialize_od.cu:207-267 -- one thread per cell; outside the pinch ONLY rho (= 0.1) is written, the other seven
 * variables keep whatever the buffer held (the reference never clears its cudaMalloc) */
void oracle_screwpinch(float* Q, float J0, float r_max_coeff, const float* gx, const float* gy, const float* gz,
                       int Nx, int Ny, int Nz) {
    (void)gz;
    dims_t dd = {Nx, Ny, Nz, (size_t)Nx * Ny * Nz};
    const size_t cube = dd.cube;
    const float r_pinch = r_max_coeff * sqrtf((float)(sq(gx[Nx - 1]) + sq(gy[Ny - 1]))); /* :219 */
    const float Jr = 0.0f, Jphi = 0.0f, Br = 0.0f, B0 = 1.0f;
    for (int k = 0; k < Nz; ++k)
        for (int i = 0; i < Nx; ++i)
            for (int j = 0; j < Ny; ++j) {
                const size_t l = IDX(&dd, i, j, k);
                const float x = gx[i], y = gy[j];
                const float r = sqrtf((float)(sq(x) + sq(y)));
                Q[l] = 0.1f; /* :237 */
                if (r < r_pinch) {
                    const float Btheta = (float)(0.5 * J0 * r * (1.0 - 0.5 * sq(r) / sq(r_pinch))); /* :243 */
                    const float p = (float)(-0.25 * (sq(J0) / pow(r_pinch, 4)) *
                                            (pow(r, 6) / 6.0 - 0.75 * sq(r_pinch) * pow(r, 4) + pow(r_pinch, 4) * sq(r))); /* :246 */
                    Q[l] = 1.0f;
                    Q[l + MX * cube] = Jr * x - Jphi * y / r;
                    Q[l + MY * cube] = Jr * y + Jphi * x / r;
                    Q[l + MZ * cube] = (float)(J0 * (1 - sq(r) / sq(r_pinch)));
                    Q[l + BX * cube] = Br * x - Btheta * y / r;
                    Q[l + BY * cube] = Br * y + Btheta * x / r;
                    Q[l + BZ * cube] = B0;
                    Q[l + EN * cube] = total_energy(Q, l, cube, p);
                }
            }
}
