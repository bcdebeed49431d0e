// Fused time step: ONE sweep Q^n -> Q^{n+1} per step.  Predictor (all cell classes), corrector,
// diffusion and every boundary pass of one reference time step (main.cu:200-213 /
// no_diffusion.cu:288-311) are evaluated on chip; the intermediate state Qint never touches HBM
// (it is a pure function of Q: SURVEY.md A.3).  Algorithmic traffic: read 8 + write 8 fp32 per
// cell-update = 64 B.
//
// Decomposition
//   - thread block = TI x 32 threads; threadIdx.x (a warp) runs along j, the unit-stride axis of
//     the reference layout, threadIdx.y along i.  Each thread owns ONE (i,j) column and marches
//     along k over a z-chunk, keeping a register queue of the column: Q(k), Q(k+1), H(Q(k+1)),
//     Qint(k-1), Qint(k).
//   - per plane, every thread publishes Q(k+1) and Qint(k) of its column once; i+-1 neighbours
//     are read back from shared memory (double buffered -> one __syncthreads per plane), j+-1
//     neighbours come from warp shuffles.
//   - the outer ring of the tile (1 cell for path A, 2 for path B) only feeds its neighbours;
//     out-of-domain threads idle.  Tiles overlap by that ring.
//   - z is cut into chunks for load balance; a chunk re-derives Qint(ka-1), Qint(ka) in two
//     warm-up planes.  At slab ends the predictor plane just outside the slab comes from the
//     qint_lo / qint_hi buffers (periodic wrap on one GPU, neighbour rank on several).
//   - the two planes with their own rules stay OUT of the marching loop (they are 2 of Nz planes and
//     would drag divergent, register-hungry code through every iteration): plane 0 (path A: periodic
//     copy of plane Nz-1; path B: the BoundaryConditions face) and, for path B, plane Nz-1 (carried
//     over except one column) are written by the small plane kernels below.
//
// Arithmetic: the "fast" recipe of imhd_math.cuh (fp32 fluxes, one reciprocal per state), with the
// O(1) sums arranged to round where the reference rounds (DESIGN.md "Precision").  The k=0 face of
// path B (1 plane in Nz) uses the exact recipe.  The translation unit is compiled with
// -fmad=false and every fused multiply-add is written out, so a value depends only on its inputs,
// never on which kernel/block computed it: results are bit-identical for any chunking or slab
// decomposition.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <vector>

#include "imhd_common.cuh"
#include "imhd_engine.h"

namespace imhd {

struct FusedArgs {
    const float* Qin;   // variable 0, array plane 0
    float* Qout;
    long long vs;       // variable stride of Qin/Qout in floats
    unsigned vs32;      // the same, unsigned 32-bit: 8 * vs * 4 B < 180 GB keeps vs below 2^32, so an output address is ONE
                        // 32x32->64-bit multiply-add on the 64-bit plane pointer
    int kbase;          // global plane index held by array plane 0
    int kmin, kmax;     // global planes that may be READ from Qin: [kmin, kmax]
    int k0, k1;         // owned (written) global planes [k0, k1)
    int ka0, kb0;       // planes the marching kernel owns: [max(k0,1), min(k1, A: Nz / B: Nz-1))
    int kfrom, kto;     // planes THIS launch writes, a sub-range of [ka0, kb0) (comm/compute overlap launches the slab ends first)
    int clen, pad_[3];  // planes per z-chunk.  clen == chunk: the chunks tile [kfrom, kto).  clen < chunk: a launch over TWO plane
                        // ranges of clen planes each, [kfrom, kfrom + clen) and [kfrom + chunk, kto) -- both slab ends of the
                        // multi-GPU loop in one launch (one full set of blocks instead of two half-empty ones).
                        // pad_: the fields behind keep their offsets modulo 16 -- ptxas pairs constant-bank loads by alignment,
                        // and another pairing re-colours the registers of the whole plane loop (measured: 3 % slower)
    const float* qlo;   // (8,Nx,Ny) Qint at global plane ka0-1  (own Qint(0) when k0 == 0)
    const float* qhi;   // (8,Nx,Ny) Qint at global plane hi_plane = min(k1, Nz-1)
    const float* qwrap; // (8,Nx,Ny) Qint(Nz-2) == Qint(-1): path B, slab with k0 == 0 only (k=0 face)
    int hi_plane;
    int chunk;          // z-chunk c of a launch starts at plane kfrom + c * chunk and holds clen planes
    int ntile_i, ntile_j;
    int jstrip;         // remainder-strip kernel only: j of its first thread row
    float corner_e;     // path B: fixed-point wall energy of column (Nx-1,Ny-1) (B-8)
    Params P;
};

template <int PATH>
struct Ring {  // halo ring width of a tile
    static constexpr int O = PATH == IMHD_PATH_A ? 1 : 2;
};

// planes [ka, kb) of z-chunk zc
// (kept to exactly this arithmetic: the register allocation of the marching kernels' plane loop follows the shape of the
// prologue, and a select between two ranges here cost the path B kernel 3 % -- tools/experiments/README.md)
__device__ __forceinline__ void chunk_range(const FusedArgs& A, int zc, int& ka, int& kb) {
    ka = A.kfrom + zc * A.chunk;
    kb = min(ka + A.clen, A.kto);
}

__device__ __forceinline__ void ldg8(const float* __restrict__ A, long long off, long long vs, float U[8]) {
#pragma unroll
    for (int v = 0; v < 8; ++v) U[v] = __ldg(A + off + v * vs);
}

__device__ __forceinline__ void hflux(const float U[8], float h[8]) { flux_idx<DIR_Z>(U, make_prim(U), h); }

// -----------------------------------------------------------------------------------------------
// Predictor (fast recipe) from evaluated fluxes, generic over the value type (imhd_math.cuh: float = one cell,
// float2 = the two rows of a register-tiled thread):
//   c = Q(i,j,k);  dF = F(Q(i+1)) - F(Q(i)) (the caller forms it: inside a register-tiled thread the i+1 flux is the
//   next row's own flux; `bottom` cells pass -F);  gc, gyp = G(Q) at the cell and at j+1;  hc, hp = H(Q) at the cell
//   and at k+1;  xsum = Q(i+1) + Q(i-1), ym, yp, zm, zp = the other neighbours of c (only where `lap`).
// Cell classes: kernels_od_intvar.cu:1160-1253, kernels_intvarbcs.cu:560-1110.
// -----------------------------------------------------------------------------------------------
template <int PATH, class V>
__device__ __forceinline__ void qint_combine(const V c[8], const V dF[8], const V gc[8], const V gyp[8], const V hc[8],
                                             const V hp[8], const V xsum[8], const V ym[8], const V yp[8], const V zm[8],
                                             const V zp[8], FlagT<V> bottom, FlagT<V> right, FlagT<V> lap, const Params& P,
                                             V out[8]) {
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const V dG = vsel(right, vneg(gc[v]), vsub(gyp[v], gc[v]));
        const V dH = vsub(hp[v], hc[v]);
        V base = c[v];
        if (v == MZ) base = vsel(fand(bottom, right), c[MX], c[MZ]);  // intRhoVZBottomRight starts from the rhovx slot (B-15)
        V r = vfmac(-P.tz, dH, vfmac(-P.ty, dG, vfmac(-P.tx, dF[v], base)));
        if (PATH == IMHD_PATH_B) r = vsel(lap, add_diffusion(r, c[v], xsum[v], vadd(yp[v], ym[v]), vadd(zp[v], zm[v]), P.dc), r);
        out[v] = r;
    }
}

// Predictor at one cell from raw states: c = Q(i,j,k), xp = Q(i+1,j,k), yp = Q(i,j+1,k), hc = H(Q(i,j,k)),
// hp = H(Q(i,j,k+1)); xm, ym, zm, zp only when `lap`.  `front` = the k = 0 edge lines with their typos (B-15).
template <int PATH>
__device__ __forceinline__ void qint_cell(const float c[8], const float xp[8], const float yp[8], const float hc[8],
                                          const float hp[8], const float xm[8], const float ym[8], const float zm[8],
                                          const float zp[8], bool bottom, bool right, bool front, bool lap,
                                          const Params& P, float out[8]) {
    const Prim s = make_prim(c);
    float f[8], g[8], fx[8], gy[8], dF[8], xsum[8], hcc[8];
    flux_idx<DIR_X>(c, s, f);
    flux_idx<DIR_Y>(c, s, g);
    flux_idx<DIR_X>(xp, make_prim(xp), fx);
    flux_idx<DIR_Y>(yp, make_prim(yp), gy);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        dF[v] = bottom ? -f[v] : fx[v] - f[v];
        xsum[v] = PATH == IMHD_PATH_B ? xp[v] + xm[v] : 0.0f;
        hcc[v] = hc[v];
    }
    if (front && right && !bottom) dF[EN] = f[EN] - f[EN];   // intEFrontRight: x difference == 0
    if (front && bottom && !right) {                         // int{RhoVZ,E}FrontBottom subtract a Y flux in the z difference,
        hcc[MZ] = g[MZ];                                     // intBZFrontBottom has a null y difference
        hcc[EN] = g[EN];
        g[BZ] = gy[BZ];
    }
    qint_combine<PATH, float>(c, dF, g, gy, hcc, hp, xsum, ym, yp, zm, zp, {bottom}, {right}, {lap}, P, out);
}

// -----------------------------------------------------------------------------------------------
// Corrector (fast recipe).  q = Q(i,j,k); c, xm, ym, zm = Qint at the cell and at i-1, j-1, k-1; xp, yp, zp = Qint at
// i+1, j+1, k+1 (path B diffusion only).
// kernels_od.cu:378-522 / :120-345, LaxWendroffAdv*Local :1206-1332, quirks B-4, B-5, B-6.
// -----------------------------------------------------------------------------------------------
// H'(Qint(i,j,k-1)) with the reference's mixed neighbours: Bsq from (Bx(i-1), By(j-1), Bz(k-1)) (B-4), Bdotu with
// rhovy(j-1) (B-5); KE and the velocities from the k-1 state itself
template <class V>
__device__ __forceinline__ void hflux_km1(const V zm[8], V xmBX, V ymBY, V ymMY, V hk[8]) {
    PrimT<V> sk;
    const V inv = vrcp(zm[RHO]);
    sk.ux = vmul(zm[MX], inv); sk.uy = vmul(zm[MY], inv); sk.uz = vmul(zm[MZ], inv);
    sk.Bsq = vfma(zm[BZ], zm[BZ], vfma(ymBY, ymBY, vmul(xmBX, xmBX)));
    const V ke = vfma3(sk.uz, zm[MZ], vfma3(sk.uy, zm[MY], vmul(sk.ux, zm[MX])));
    sk.p = vmulc(kGm1f, vfmac(-0.5f, sk.Bsq, vsub(zm[EN], ke)));
    sk.ptot = vfmac(0.5f, sk.Bsq, sk.p);
    sk.Bdotu = vfma3(sk.uz, zm[BZ], vfma3(vmul(ymMY, inv), zm[BY], vmul(sk.ux, zm[BX])));
    flux_loc<DIR_Z>(zm, sk, hk);
}

// From evaluated fluxes: dF = F'(Qint(i)) - F'(Qint(i-1)) (formed by the caller, with B-6), gc, hc = G', H' of Qint at
// the cell; gj = G'(Qint(j-1)), hk = hflux_km1; xsum = Qint(i+1) + Qint(i-1).
template <int PATH, class V>
__device__ __forceinline__ void corr_combine(const V q[8], const V c[8], const V dF[8], const V gc[8], const V gj[8],
                                             const V hc[8], const V hk[8], const V xsum[8], const V ym[8], const V yp[8],
                                             const V zm[8], const V zp[8], const Params& P, V out[8]) {
    const float hx = 0.5f * P.tx, hy = 0.5f * P.ty, hz = 0.5f * P.tz;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const V dG = vsub(gc[v], gj[v]), dH = vsub(hc[v], hk[v]);
        // reference: float(0.5*(q+c) - 0.5tx dF - 0.5ty dG - 0.5tz dH) with q+c an fp32 sum and the rest
        // fp64.  0.5*s is exact and T is ~1e-2 of it, so one fp32 rounding of (0.5 s - T) reproduces it.
        const V s = vadd(q[v], c[v]);
        const V T = vfmac(hx, dF[v], vfmac(hy, dG, vmulc(hz, dH)));
        V r = vfmac(0.5f, s, vneg(T));
        if (PATH == IMHD_PATH_B) r = add_diffusion(r, c[v], xsum[v], vadd(yp[v], ym[v]), vadd(zp[v], zm[v]), P.dc);
        out[v] = r;
    }
}

// From raw states.  Must be called by all 32 lanes of a warp whose lanes hold consecutive j (LANES_I: consecutive
// i); it shuffles.
template <int PATH, bool LANES_I = false>
__device__ __forceinline__ void corr_cell(const float q[8], const float c[8], const float xm[8], const float ym[8],
                                          const float zm[8], const float xp[8], const float yp[8], const float zp[8],
                                          const Params& P, float out[8]) {
    const Prim sc = make_prim(c);
    float fc[8], gc[8], hc[8], fi[8], gj[8], hk[8], dF[8], xsum[8];
    flux_loc<DIR_X>(c, sc, fc);
    flux_loc<DIR_Y>(c, sc, gc);
    flux_loc<DIR_Z>(c, sc, hc);
    // G'(Qint(i,j-1,k)) is what lane-1 has just computed as its own gc (same inputs, same function, same bits): one
    // shuffle per component instead of re-deriving the neighbour's primitives and flux.  Lane 0 never writes output.
    // With lanes along i (remainder-strip kernel) the same holds for F'(Qint(i-1,j,k)) instead.
    if (LANES_I) {
        flux_loc<DIR_Y>(ym, make_prim(ym), gj);
#pragma unroll
        for (int v = 0; v < 8; ++v) fi[v] = v == BX ? 0.0f : __shfl_up_sync(0xffffffffu, fc[v], 1);
    } else {
        flux_loc<DIR_X>(xm, make_prim(xm), fi);
#pragma unroll
        for (int v = 0; v < 8; ++v) gj[v] = v == BY ? 0.0f : __shfl_up_sync(0xffffffffu, gc[v], 1);
    }
    if (PATH == IMHD_PATH_B) fi[RHO] = xm[RHO];  // B-6
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        dF[v] = fc[v] - fi[v];
        xsum[v] = PATH == IMHD_PATH_B ? xp[v] + xm[v] : 0.0f;
    }
    hflux_km1<float>(zm, xm[BX], ym[BY], ym[MY], hk);
    corr_combine<PATH, float>(q, c, dF, gc, gj, hc, hk, xsum, ym, yp, zm, zp, P, out);
}

// Path B, k = 0 face (BoundaryConditions, kernels_fluidbcs.cu:52-116): corrector with INDEXED fluxes of
// Qint, k-1 -> Nz-2, + dt*numericalDiffusionFront; exact recipe (1 plane in Nz, not worth a fast one).
__device__ __forceinline__ void front_cell_exact(const float q[8], const float c[8], const float xm[8], const float ym[8],
                                              const float zm[8], const float xp[8], const float yp[8],
                                              const float zp[8], const Params& P, float out[8]) {
    float f[8], g[8], h[8], t[8], dF[8], dG[8], dH[8];
    const Aux<true> a = make_aux<true>(c);
    flux_indexed<true, DIR_X>(c, a, f);
    flux_indexed<true, DIR_Y>(c, a, g);
    flux_indexed<true, DIR_Z>(c, a, h);
    flux_indexed<true, DIR_X>(xm, make_aux<true>(xm), t);
#pragma unroll
    for (int v = 0; v < 8; ++v) dF[v] = f[v] - t[v];
    flux_indexed<true, DIR_Y>(ym, make_aux<true>(ym), t);
#pragma unroll
    for (int v = 0; v < 8; ++v) dG[v] = g[v] - t[v];
    flux_indexed<true, DIR_Z>(zm, make_aux<true>(zm), t);
#pragma unroll
    for (int v = 0; v < 8; ++v) dH[v] = h[v] - t[v];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float nd = num_diff<true>(c[v], xp[v], yp[v], zp[v], xm[v], ym[v], zm[v], P.D, P.dc);
        const float pn = P.dt * nd;
        out[v] = (float)(0.5 * (q[v] + c[v]) - 0.5 * P.tx * dF[v] - 0.5 * P.ty * dG[v] - 0.5 * P.tz * dH[v] + pn);
    }
}

// -----------------------------------------------------------------------------------------------
// The fused z-marching kernel, plain-load variant: planes [ka0, kb0) of the slab.  Used when the
// grid does not meet TMA's 16-byte row-stride rule (Ny % 4 != 0); same device functions, same bits
// as the TMA variant below.
// -----------------------------------------------------------------------------------------------
template <int PATH, int TI>
__global__ void __launch_bounds__(TI * 32, 1) k_fused_step_ldg(const FusedArgs A) {
    constexpr int O = Ring<PATH>::O;
    constexpr int WI = TI - 2 * O, WJ = 32 - 2 * O;
    extern __shared__ float smem[];
    // sQ[buf][v][ti][lane], then sQi[buf][v][ti][lane]
    float* sQ = smem;
    float* sQi = smem + 2 * 8 * TI * 32;

    const Params& P = A.P;
    const int lane = threadIdx.x, ti = threadIdx.y;
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int i = bi * WI + 1 - O + ti, j = bj * WJ + 1 - O + lane;
    const bool in_dom = i >= 0 && i < P.Nx && j >= 0 && j < P.Ny;
    const int ic = min(max(i, 0), P.Nx - 1), jc = min(max(j, 0), P.Ny - 1);
    const long long lcol = (long long)ic * P.Ny + jc;
    const bool top = (i == 0), bottom = (i == P.Nx - 1), left = (j == 0), right = (j == P.Ny - 1);
    const bool interior_ij = in_dom && !top && !bottom && !left && !right;
    // cells the corrector updates in this path (k range is handled by the loop bounds)
    const bool upd = PATH == IMHD_PATH_A ? (in_dom && !top && !left) : interior_ij;
    // cells this thread writes: the tile's inner window, widened to the domain edge on edge tiles
    const int oi_lo = bi == 0 ? 0 : 1 + bi * WI, oi_hi = bi == A.ntile_i - 1 ? P.Nx : 1 + (bi + 1) * WI;
    const int oj_lo = bj == 0 ? 0 : 1 + bj * WJ, oj_hi = bj == A.ntile_j - 1 ? P.Ny : 1 + (bj + 1) * WJ;
    const bool owner = in_dom && i >= oi_lo && i < oi_hi && j >= oj_lo && j < oj_hi;
    const int tim = max(ti - 1, 0), tip = min(ti + 1, TI - 1);
    const int so = ti * 32 + lane, som = tim * 32 + lane, sop = tip * 32 + lane;  // smem slots: own, i-1, i+1

    int ka, kb;
    chunk_range(A, blockIdx.z, ka, kb);
    const bool first = ka == A.ka0;  // the plane below is the slab's qint_lo; otherwise re-derive it in a warm-up plane
    const int ks = first ? ka - 1 : ka - 2;

    auto plane_ptr = [&](int k) -> const float* {
        const int kc = min(max(k, A.kmin), A.kmax);
        return A.Qin + (long long)(kc - A.kbase) * P.plane + lcol;
    };

    float q0[8], q1[8], h1[8], qim[8], qic[8];
    ldg8(plane_ptr(ks), 0, A.vs, q0);
    ldg8(plane_ptr(ks + 1), 0, A.vs, q1);
    hflux(q1, h1);
#pragma unroll
    for (int v = 0; v < 8; ++v) { qim[v] = 1.0f; qic[v] = 1.0f; }
    if (first) ldg8(A.qlo, lcol, P.plane, qic);  // Qint(ka-1)

    float* outp = A.Qout + (long long)(ka - A.kbase) * P.plane + lcol;
    for (int k = ks; k < kb; ++k) {
        const int buf = (k - ks) & 1;
        const int kp = k + 1;
        float qn[8], hn[8], qip[8];
        ldg8(plane_ptr(k + 2), 0, A.vs, qn);
        float* bQ = sQ + buf * 8 * TI * 32;
        float* bQi = sQi + buf * 8 * TI * 32;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            bQ[v * TI * 32 + so] = q1[v];
            bQi[v * TI * 32 + so] = qic[v];
        }
        __syncthreads();

        // ---- predictor plane kp = k+1 (kp >= 1 here: plane 0 is never re-derived in the loop) --------------
        hflux(qn, hn);
        if (kp == A.hi_plane) {
            ldg8(A.qhi, lcol, P.plane, qip);
        } else {
            float xp[8], yp[8], xm[8], ym[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xp[v] = bQ[v * TI * 32 + sop];
                yp[v] = __shfl_down_sync(0xffffffffu, q1[v], 1);
            }
            if (PATH == IMHD_PATH_B) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    xm[v] = bQ[v * TI * 32 + som];
                    ym[v] = __shfl_up_sync(0xffffffffu, q1[v], 1);
                }
            }
            qint_cell<PATH>(q1, xp, yp, h1, hn, xm, ym, q0, qn, bottom, right, false, interior_ij, P, qip);
        }

        // ---- corrector plane k -----------------------------------------------------------------------
        if (k >= ka) {
            float out[8], xm[8], ym[8], xp[8], yp[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xm[v] = bQi[v * TI * 32 + som];
                ym[v] = __shfl_up_sync(0xffffffffu, qic[v], 1);
            }
            if (PATH == IMHD_PATH_B) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    xp[v] = bQi[v * TI * 32 + sop];
                    yp[v] = __shfl_down_sync(0xffffffffu, qic[v], 1);
                }
            }
            corr_cell<PATH>(q0, qic, xm, ym, qim, xp, yp, qip, P, out);
            if (owner) {
#pragma unroll
                for (int v = 0; v < 8; ++v) outp[v * A.vs] = upd ? out[v] : q0[v];  // untouched cells are carried over
            }
            outp += P.plane;
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            q0[v] = q1[v]; q1[v] = qn[v]; h1[v] = hn[v];
            qim[v] = qic[v]; qic[v] = qip[v];
        }
    }
}

// -----------------------------------------------------------------------------------------------
// The fused z-marching kernel, TMA variant (the hot path).
//
//   - thread tile = the region where the predictor is evaluated: TI x 32 cells, one column per thread
//   - per plane ONE cp.async.bulk.tensor (4-D box: 40 x (TI+2) x 1 plane x 8 variables, 23 KB) stages the Q
//     tile INCLUDING a one-cell ring around the thread tile into shared memory; out-of-domain elements are
//     zero-filled by the hardware and only ever feed lanes whose results are discarded.  Three stages
//     (planes k+1, k+2 resident, k+3 in flight), one mbarrier each; the stage of plane k is recycled right
//     after the per-plane __syncthreads.
//   - own-column queue in registers: Q(k), Q(k+1), H(Q(k+1)), Qint(k-1), Qint(k); Q(k+2) is read from the
//     tile; the four lateral neighbours of Q(k+1) come from the tile, those of Qint(k) from a
//     double-buffered exchange array (i+-1) and warp shuffles (j+-1)
//   - the corrector is written for ti in [1, TI-2], lane in [1, 30] (path B; one more row/lane for path A):
//     14x30 of 16x32 threads produce output (82 %), tiles overlap by the one-cell Qint ring
//   - the k loop is unrolled by 3 so that stage indices are static and the register queue rotates by renaming
// -----------------------------------------------------------------------------------------------
constexpr int kTC = 40;  // tile columns: 32 + 2 ring + up to 3 of alignment slack (a TMA box must START on a 16-byte boundary)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(0)
        : "memory");
}

template <int PATH, int TI>
struct TmaGeo {
    static constexpr int TR = TI + 2;                                   // tile rows
    static constexpr int WI = PATH == IMHD_PATH_A ? TI - 1 : TI - 2;    // output rows per tile
    static constexpr int WJ = PATH == IMHD_PATH_A ? 31 : 30;            // output lanes per tile
    static constexpr int STAGE_FLOATS = 8 * TR * kTC;
    static constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
    static constexpr int XCH_FLOATS = 2 * 8 * TI * 32;
    static constexpr size_t SMEM = 3 * STAGE_BYTES + XCH_FLOATS * 4 + 64;
};

template <int PATH, int TI>
struct TmaThread {  // per-thread constants of the TMA kernel
    bool bottom, right, interior_ij, upd, owner, corr_row;
    int own, so, som, sop;  // tile offset of the own cell; exchange slots own / i-1 / i+1
    long long lcol;
};

// One plane of the march: predictor plane k+1, corrector plane k.  tq1 / tq2 = tiles of planes k+1 / k+2.
template <int PATH, int TI>
__device__ __forceinline__ void tma_plane(const FusedArgs& A, const TmaThread<PATH, TI>& T, int k, int ka, const float* tq1,
                                          const float* tq2, float* xq, const float (&q0)[8], const float (&q1)[8],
                                          float (&qn)[8], const float (&h1)[8], float (&hn)[8], const float (&qim)[8],
                                          const float (&qic)[8], float (&qip)[8], float*& outp) {
    using G = TmaGeo<PATH, TI>;
    const Params& P = A.P;
    constexpr int VS = G::TR * kTC;  // variable stride inside a tile
#pragma unroll
    for (int v = 0; v < 8; ++v) qn[v] = tq2[v * VS + T.own];
    hflux(qn, hn);
    if (k + 1 == A.hi_plane) {
        ldg8(A.qhi, T.lcol, P.plane, qip);
    } else {
        float xp[8], yp[8], xm[8], ym[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            xp[v] = tq1[v * VS + T.own + kTC];
            yp[v] = tq1[v * VS + T.own + 1];
        }
        if (PATH == IMHD_PATH_B) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xm[v] = tq1[v * VS + T.own - kTC];
                ym[v] = tq1[v * VS + T.own - 1];
            }
        }
        qint_cell<PATH>(q1, xp, yp, h1, hn, xm, ym, q0, qn, T.bottom, T.right, false, T.interior_ij, P, qip);
    }
    if (k >= ka) {
        if (T.corr_row) {  // warp-uniform: the first (and for path B the last) row of the tile only feeds its neighbours
            float out[8], xm[8], ym[8], xp[8], yp[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xm[v] = xq[v * TI * 32 + T.som];
                ym[v] = __shfl_up_sync(0xffffffffu, qic[v], 1);
            }
            if (PATH == IMHD_PATH_B) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    xp[v] = xq[v * TI * 32 + T.sop];
                    yp[v] = __shfl_down_sync(0xffffffffu, qic[v], 1);
                }
            }
            corr_cell<PATH>(q0, qic, xm, ym, qim, xp, yp, qip, P, out);
            if (T.owner) {
#pragma unroll
                for (int v = 0; v < 8; ++v) outp[v * A.vs] = T.upd ? out[v] : q0[v];  // untouched cells are carried over
            }
        } else if (T.owner) {
#pragma unroll
            for (int v = 0; v < 8; ++v) outp[v * A.vs] = q0[v];
        }
        outp += P.plane;
    }
}

template <int PATH, int TI>
__global__ void __launch_bounds__(TI * 32, 1) k_fused_step_tma(const FusedArgs A, const __grid_constant__ CUtensorMap tmap) {
    using G = TmaGeo<PATH, TI>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                       // [3][8][TR][kTC]
    float* xch = tiles + 3 * G::STAGE_FLOATS;                                 // [2][8][TI][32]  Qint exchange
    uint64_t* full = reinterpret_cast<uint64_t*>(xch + G::XCH_FLOATS);        // [3]

    const Params& P = A.P;
    const int lane = threadIdx.x, ti = threadIdx.y;
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int ib = bi * G::WI, jb = bj * G::WJ;
    const int i = ib + ti, j = jb + lane;
    const bool in_dom = i < P.Nx && j < P.Ny;
    TmaThread<PATH, TI> T;
    T.bottom = (i == P.Nx - 1);
    T.right = (j == P.Ny - 1);
    T.interior_ij = in_dom && i > 0 && !T.bottom && j > 0 && !T.right;
    T.upd = PATH == IMHD_PATH_A ? (in_dom && i > 0 && j > 0) : T.interior_ij;
    const int oi_lo = bi == 0 ? 0 : ib + 1, oi_hi = bi == A.ntile_i - 1 ? P.Nx : ib + G::WI + 1;
    const int oj_lo = bj == 0 ? 0 : jb + 1, oj_hi = bj == A.ntile_j - 1 ? P.Ny : jb + G::WJ + 1;
    T.owner = in_dom && i >= oi_lo && i < oi_hi && j >= oj_lo && j < oj_hi;
    T.corr_row = ti >= 1 && (PATH == IMHD_PATH_A || ti <= TI - 2);
    // the box starts at the 4-column boundary at or below jb-1 (measured: a misaligned inner coordinate traps)
    const int c0 = ((jb - 1 + 4) / 4) * 4 - 4;
    T.own = (ti + 1) * kTC + (jb - 1 - c0) + lane + 1;
    T.so = ti * 32 + lane;
    T.som = max(ti - 1, 0) * 32 + lane;
    T.sop = min(ti + 1, TI - 1) * 32 + lane;
    T.lcol = (long long)min(i, P.Nx - 1) * P.Ny + min(j, P.Ny - 1);

    int ka, kb;
    chunk_range(A, blockIdx.z, ka, kb);
    const bool first = ka == A.ka0;  // the plane below is the slab's qint_lo; otherwise re-derive it in a warm-up plane
    const int ks = first ? ka - 1 : ka - 2;
    const bool producer = (threadIdx.x == 0 && threadIdx.y == 0);
    const int klast = kb + 1;  // last plane any iteration reads

    auto issue = [&](int plane, int stage) {  // producer only
        const int kc = min(max(plane, A.kmin), A.kmax) - A.kbase;
        mbar_expect_tx(&full[stage], G::STAGE_BYTES);
        tma_load_tile(tiles + stage * G::STAGE_FLOATS, &tmap, &full[stage], c0, ib - 1, kc);
    };

    if (producer) {
        for (int s = 0; s < 3; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    if (producer)
        for (int s = 0; s < 3; ++s) issue(ks + s, s);

    float qa[8], qb[8], qc[8], ha[8], hb[8], hc[8], ia[8], ib_[8], ic[8];
    constexpr int VS = G::TR * kTC;
    mbar_wait(&full[0], 0);
    mbar_wait(&full[1], 0);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        qa[v] = tiles[v * VS + T.own];
        qb[v] = tiles[G::STAGE_FLOATS + v * VS + T.own];
        ia[v] = 1.0f;
        ib_[v] = 1.0f;
    }
    hflux(qb, hb);
    if (first) ldg8(A.qlo, T.lcol, P.plane, ib_);  // Qint(ka-1)

    float* outp = A.Qout + (long long)(ka - A.kbase) * P.plane + T.lcol;
    float* t0 = tiles;
    float* t1 = tiles + G::STAGE_FLOATS;
    float* t2 = tiles + 2 * G::STAGE_FLOATS;
    float* x0 = xch;
    float* x1 = xch + 8 * TI * 32;

    // one step of the march; roles (Q(k),Q(k+1),Q(k+2)) / (H(k+1),H(k+2)) / (Qint(k-1),Qint(k),Qint(k+1)) / tiles rotate by renaming
#define IMHD_MARCH(IT, Q0, Q1, QN, H1, HN, IM, IC, IP, TFREE, TQ1, TQ2, SFREE, S2)                       \
    if (k < kb) {                                                                                         \
        float* xq = ((IT)&1) ? x1 : x0;                                                                   \
        _Pragma("unroll") for (int v = 0; v < 8; ++v) xq[v * TI * 32 + T.so] = IC[v];                     \
        mbar_wait(&full[S2], par##S2);                                                                    \
        par##S2 ^= 1;                                                                                     \
        __syncthreads();                                                                                  \
        if (producer && k + 3 <= klast) issue(k + 3, SFREE);                                              \
        tma_plane<PATH, TI>(A, T, k, ka, TQ1, TQ2, xq, Q0, Q1, QN, H1, HN, IM, IC, IP, outp);            \
        ++k;                                                                                              \
    }

    // parity of the NEXT wait on each stage: stages 0,1 were waited once in the prologue
    uint32_t par0 = 1, par1 = 1, par2 = 0;
    int k = ks;
    while (k < kb) {
        // it % 6 pattern: exchange buffer alternates (period 2), stages rotate (period 3)
        IMHD_MARCH(0, qa, qb, qc, hb, hc, ia, ib_, ic, t0, t1, t2, 0, 2)
        IMHD_MARCH(1, qb, qc, qa, hc, ha, ib_, ic, ia, t1, t2, t0, 1, 0)
        IMHD_MARCH(2, qc, qa, qb, ha, hb, ic, ia, ib_, t2, t0, t1, 2, 1)
        IMHD_MARCH(3, qa, qb, qc, hb, hc, ia, ib_, ic, t0, t1, t2, 0, 2)
        IMHD_MARCH(4, qb, qc, qa, hc, ha, ib_, ic, ia, t1, t2, t0, 1, 0)
        IMHD_MARCH(5, qc, qa, qb, ha, hb, ic, ia, ib_, t2, t0, t1, 2, 1)
    }
#undef IMHD_MARCH
}

// -----------------------------------------------------------------------------------------------
// The fused z-marching kernel, register-tiled (the hot path).  Same march, same TMA staging and the same device
// functions as k_fused_step_tma, but every thread owns TWO consecutive rows (i, i+1) of one column j, held as
// float2 (.x = first row, .y = second row), so that
//   - F(Q) and F'(Qint) of the first row are the i-1 neighbour flux of the second and the second row's flux is the
//     first row's i+1 neighbour (registers, no exchange); the x neighbours of the Laplacians likewise,
//   - the primitives of Q(k+1) are carried in the register queue from the iteration that evaluated H(Q(k+1)),
//   - everything that is symmetric in the two rows and has at most two distinct register operands issues as ONE packed
//     fp32x2 instruction (imhd_math.cuh),
//   - addressing, predicates, loop control and the barrier are paid once per two cells.
// The thread tile is 2*TI rows x 32 lanes: 8 warps give the 16x32 tile of the one-row kernel with half the threads
// (<= 255 registers, no spills).  The march is ONE rolled copy of the plane body -- it has to fit the instruction
// cache: with 8 warps per SM nobody hides a fetch miss (measured: 3x unrolled 29.9, rolled 31.6 GLUPS) -- so the
// register queue rotates by moves.
// -----------------------------------------------------------------------------------------------
template <int PATH, int TI>
struct PairGeo {
    static constexpr int NR = TI * 2;                                   // rows of the thread tile
    static constexpr int TR = NR + 2;                                   // tile rows (one ring row either side)
    static constexpr int WI = PATH == IMHD_PATH_A ? NR - 1 : NR - 2;    // output rows per tile
    static constexpr int WJ = PATH == IMHD_PATH_A ? 31 : 30;            // output lanes per tile
    static constexpr int STAGE_FLOATS = 8 * TR * kTC;
    static constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
    static constexpr int XBUF = 2 * 8 * TI * 32;                        // one exchange buffer: [first,last][v][ti][lane]
    static constexpr int NSTAGE = 4;                                    // planes k+1, k+2 in use, k+3, k+4 in flight
    static constexpr int NXBUF = 2;                                     // exchange buffers
    static constexpr size_t SMEM = NSTAGE * STAGE_BYTES + NXBUF * XBUF * 4 + 64;
};

struct PairThread {  // per-thread constants
    unsigned rows;          // per row rr: bit rr = bottom, 8+rr = interior i, 16+rr = corrector-updated row, 24+rr = owner row
    unsigned lanes;         // bit 0 = right, 1 = interior j, 2 = corrector-updated column, 3 = owner column
    int own;                // tile offset of the cell of the first row
    int xs, xsm, xsp;       // exchange slots: own, thread row above, thread row below
    int i0, jc;             // first row; column clamped into the domain
    __device__ __forceinline__ bool bottom(int rr) const { return (rows >> rr) & 1u; }
    __device__ __forceinline__ bool interior_i(int rr) const { return (rows >> (8 + rr)) & 1u; }
    __device__ __forceinline__ bool upd_i(int rr) const { return (rows >> (16 + rr)) & 1u; }
    __device__ __forceinline__ bool owner_i(int rr) const { return (rows >> (24 + rr)) & 1u; }
    __device__ __forceinline__ bool right() const { return lanes & 1u; }
    __device__ __forceinline__ bool interior_j() const { return lanes & 2u; }
    __device__ __forceinline__ bool upd_j() const { return lanes & 4u; }
    __device__ __forceinline__ bool owner_j() const { return lanes & 8u; }
};

// One plane of the march for the two rows of a thread, in two phases: predictor plane k+1, corrector plane k.
// Straight-line on purpose: with 8 warps per SM every taken branch is an instruction-fetch bubble nobody hides, so the
// tile's ring rows and the warm-up planes run the corrector too and only their STORES are predicated off (`store_ok`).
// W = false: the block holds no wall cell, no cell outside the domain and no cell next to one (an "interior" block): every
// wall / edge select below is decided at compile time.
template <int PATH, int TI, bool W = true>
__device__ __forceinline__ void pair_predict(const FusedArgs& A, const PairThread& T, bool hi, const float* tq1, const float* tq2,
                                             const float2 (&q0)[8], const float2 (&q1)[8], float2 (&qn)[8],
                                             const float2 (&h1)[8], float2 (&hn)[8], const PrimT<float2>& p1,
                                             PrimT<float2>& pn, float2 (&qip)[8]) {
    using G = PairGeo<PATH, TI>;
    const Params& P = A.P;
    constexpr int VS = G::TR * kTC;   // variable stride inside a tile
    const bool rt = W && T.right(), b0 = W && T.bottom(0), b1 = W && T.bottom(1);
    const bool in0 = !W || (T.interior_i(0) && T.interior_j()), in1 = !W || (T.interior_i(1) && T.interior_j());
    const FlagT<float2> right = {rt, rt};
#pragma unroll
    for (int v = 0; v < 8; ++v) qn[v] = make_float2(tq2[v * VS + T.own], tq2[v * VS + T.own + kTC]);
    pn = make_prim(qn);
    flux_idx<DIR_Z>(qn, pn, hn);
    {
        float xlast[8], xfirst[8], fl[8];
        float2 yp[8], ym[8], f[8], g[8], gy[8], dF[8], xsum[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            xlast[v] = tq1[v * VS + T.own + 2 * kTC];                       // Q(k+1) one row below the thread's rows
            yp[v] = make_float2(tq1[v * VS + T.own + 1], tq1[v * VS + T.own + kTC + 1]);
            if (PATH == IMHD_PATH_B) {
                xfirst[v] = tq1[v * VS + T.own - kTC];                      // ... and one row above
                ym[v] = make_float2(tq1[v * VS + T.own - 1], tq1[v * VS + T.own + kTC - 1]);
            }
        }
        flux_idx<DIR_X>(q1, p1, f);
        flux_idx<DIR_X>(xlast, make_prim(xlast), fl);
        flux_idx<DIR_Y>(q1, p1, g);
        flux_idx<DIR_Y>(yp, make_prim(yp), gy);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            dF[v].x = b0 ? -f[v].x : f[v].y - f[v].x;   // the second row's flux is the first row's i+1 neighbour
            dF[v].y = b1 ? -f[v].y : fl[v] - f[v].y;
            if (PATH == IMHD_PATH_B) xsum[v] = make_float2(q1[v].y + xfirst[v], xlast[v] + q1[v].x);
        }
        qint_combine<PATH, float2>(q1, dF, g, gy, h1, hn, xsum, ym, yp, q0, qn, {b0, b1}, right, {in0, in1}, P, qip);
    }
    if (__builtin_expect(hi, 0)) {  // the plane above the slab comes from the neighbour (or is the periodic image): once per slab
        float a[8], b[8];
        ldg8(A.qhi, (long long)min(T.i0, P.Nx - 1) * P.Ny + T.jc, P.plane, a);
        ldg8(A.qhi, (long long)min(T.i0 + 1, P.Nx - 1) * P.Ny + T.jc, P.plane, b);
#pragma unroll
        for (int v = 0; v < 8; ++v) qip[v] = make_float2(a[v], b[v]);
    }
}

template <int PATH, int TI, bool W = true>
__device__ __forceinline__ void pair_correct(const FusedArgs& A, const PairThread& T, bool store_ok, const float* xq,
                                             const float2 (&q0)[8], const float2 (&qim)[8], const float2 (&qic)[8],
                                             const float2 (&qip)[8], float* outp) {
    const Params& P = A.P;
    constexpr int XV = TI * 32;       // variable stride inside an exchange buffer
    float xfirst[8], xlast[8], ff[8];
    float2 fc[8], gc[8], hc[8], gj[8], hk[8], ym[8], yp[8], dF[8], xsum[8], out[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        xfirst[v] = xq[(8 + v) * XV + T.xsm];                          // Qint(k), second row of the thread row above
        if (PATH == IMHD_PATH_B) xlast[v] = xq[v * XV + T.xsp];        // ... first row of the thread row below
    }
    const PrimT<float2> sc = make_prim(qic);
    flux_loc<DIR_X>(qic, sc, fc);
    flux_loc<DIR_Y>(qic, sc, gc);
    flux_loc<DIR_Z>(qic, sc, hc);
    flux_loc<DIR_X>(xfirst, make_prim(xfirst), ff);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        ym[v] = make_float2(__shfl_up_sync(0xffffffffu, qic[v].x, 1), __shfl_up_sync(0xffffffffu, qic[v].y, 1));
        if (PATH == IMHD_PATH_B)
            yp[v] = make_float2(__shfl_down_sync(0xffffffffu, qic[v].x, 1), __shfl_down_sync(0xffffffffu, qic[v].y, 1));
        gj[v] = v == BY ? make_float2(0.0f, 0.0f)
                        : make_float2(__shfl_up_sync(0xffffffffu, gc[v].x, 1), __shfl_up_sync(0xffffffffu, gc[v].y, 1));
        // F'(Qint(i-1)): the thread row above for the first row, the first row's own flux for the second (B-6: rho)
        const float fa = (PATH == IMHD_PATH_B && v == RHO) ? xfirst[RHO] : ff[v];
        const float fb = (PATH == IMHD_PATH_B && v == RHO) ? qic[RHO].x : fc[v].x;
        dF[v] = make_float2(fc[v].x - fa, fc[v].y - fb);
        if (PATH == IMHD_PATH_B) xsum[v] = make_float2(qic[v].y + xfirst[v], xlast[v] + qic[v].x);
    }
    hflux_km1<float2>(qim, make_float2(xfirst[BX], qic[BX].x), ym[BY], ym[MY], hk);
    corr_combine<PATH, float2>(q0, qic, dF, gc, gj, hc, hk, xsum, ym, yp, qim, qip, P, out);
    if (store_ok && T.owner_j()) {
        const bool ua = !W || (T.upd_i(0) && T.upd_j()), ub = !W || (T.upd_i(1) && T.upd_j());   // an interior block's owners are all updated
        if (T.owner_i(0)) {
            char* o = reinterpret_cast<char*>(outp);
#pragma unroll
            for (int v = 0; v < 8; ++v)  // untouched cells are carried over
                *reinterpret_cast<float*>(o + (unsigned long long)A.vs32 * (unsigned)(4 * v)) = ua ? out[v].x : q0[v].x;
        }
        if (T.owner_i(1)) {
            char* o = reinterpret_cast<char*>(outp + P.Ny);
#pragma unroll
            for (int v = 0; v < 8; ++v)
                *reinterpret_cast<float*>(o + (unsigned long long)A.vs32 * (unsigned)(4 * v)) = ub ? out[v].y : q0[v].y;
        }
    }
}

// keeps a per-thread constant in its register: without this ptxas re-derives it from S2R / LDC every plane
__device__ __forceinline__ void keep(int& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void keep(unsigned& x) { asm volatile("" : "+r"(x)); }

template <int PATH, int TI, int MINB = 1>
__global__ void __launch_bounds__(TI * 32, MINB) k_fused_pair(const FusedArgs A, const __grid_constant__ CUtensorMap tmap) {
    using G = PairGeo<PATH, TI>;
    constexpr int NS = G::NSTAGE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                       // [NS][8][TR][kTC]
    float* xch = tiles + NS * G::STAGE_FLOATS;                                // [2][first,last][8][TI][32]  Qint exchange
    uint64_t* full = reinterpret_cast<uint64_t*>(xch + G::NXBUF * G::XBUF);   // [NS] tile landed

    const Params& P = A.P;
    const int lane = threadIdx.x, ti = threadIdx.y;
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int ib = bi * G::WI, jb = bj * G::WJ;
    const int t0 = ti * 2, i0 = ib + t0, j = jb + lane;
    const int oi_lo = bi == 0 ? 0 : ib + 1, oi_hi = bi == A.ntile_i - 1 ? P.Nx : ib + G::WI + 1;
    const int oj_lo = bj == 0 ? 0 : jb + 1, oj_hi = bj == A.ntile_j - 1 ? P.Ny : jb + G::WJ + 1;
    PairThread T;
    T.rows = 0;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int i = i0 + rr, t = t0 + rr;
        const bool corr_row = t >= 1 && (PATH == IMHD_PATH_A || t <= G::NR - 2);  // rows with valid Qint neighbours
        if (i == P.Nx - 1) T.rows |= 1u << rr;
        if (i > 0 && i < P.Nx - 1) T.rows |= 1u << (8 + rr);
        if (corr_row && i > 0 && (PATH == IMHD_PATH_A ? i < P.Nx : i < P.Nx - 1)) T.rows |= 1u << (16 + rr);
        if (i < P.Nx && i >= oi_lo && i < oi_hi) T.rows |= 1u << (24 + rr);
    }
    const bool interior_j = j > 0 && j < P.Ny - 1;
    T.lanes = (j == P.Ny - 1 ? 1u : 0u) | (interior_j ? 2u : 0u) |
              ((PATH == IMHD_PATH_A ? (j > 0 && j < P.Ny) : interior_j) ? 4u : 0u) | ((j < P.Ny && j >= oj_lo && j < oj_hi) ? 8u : 0u);
    // the box starts at the 4-column boundary at or below jb-1 (measured: a misaligned inner coordinate traps)
    const int c0 = ((jb - 1 + 4) / 4) * 4 - 4;
    T.own = (t0 + 1) * kTC + (jb - 1 - c0) + lane + 1;
    T.xs = ti * 32 + lane;
    T.xsm = max(ti - 1, 0) * 32 + lane;
    T.xsp = min(ti + 1, TI - 1) * 32 + lane;
    T.i0 = i0;
    T.jc = min(j, P.Ny - 1);
    keep(T.rows); keep(T.lanes); keep(T.own); keep(T.xs); keep(T.xsm); keep(T.xsp);

    int ka, kb;
    chunk_range(A, blockIdx.z, ka, kb);
    const bool first = ka == A.ka0;  // the plane below is the slab's qint_lo; otherwise re-derive it in a warm-up plane
    const int ks = first ? ka - 1 : ka - 2;
    const bool producer = (threadIdx.x == 0 && threadIdx.y == 0);
    const int klast = kb + 1;  // last plane any iteration reads

    auto issue = [&](int plane, int stage) {  // producer only
        const int kc = min(max(plane, A.kmin), A.kmax) - A.kbase;
        mbar_expect_tx(&full[stage], G::STAGE_BYTES);
        tma_load_tile(tiles + stage * G::STAGE_FLOATS, &tmap, &full[stage], c0, ib - 1, kc);
    };

    if (producer) {
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    if (producer)
        for (int s = 0; s < NS; ++s) issue(ks + s, s);

    float2 qa[8], qb[8], qc[8], hb[8], hc[8], ia[8], ib_[8], ic[8];
    PrimT<float2> pb, pc;
    constexpr int VS = G::TR * kTC;
    mbar_wait(&full[0], 0);
    mbar_wait(&full[1], 0);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        qa[v] = make_float2(tiles[v * VS + T.own], tiles[v * VS + T.own + kTC]);
        qb[v] = make_float2(tiles[G::STAGE_FLOATS + v * VS + T.own], tiles[G::STAGE_FLOATS + v * VS + T.own + kTC]);
        ia[v] = make_float2(1.0f, 1.0f);
        ib_[v] = make_float2(1.0f, 1.0f);
    }
    pb = make_prim(qb);
    flux_idx<DIR_Z>(qb, pb, hb);
    if (first) {  // Qint(ka-1); rows beyond the domain read the last row (they never produce output)
        float a[8], b[8];
        ldg8(A.qlo, (long long)min(i0, P.Nx - 1) * P.Ny + T.jc, P.plane, a);
        ldg8(A.qlo, (long long)min(i0 + 1, P.Nx - 1) * P.Ny + T.jc, P.plane, b);
#pragma unroll
        for (int v = 0; v < 8; ++v) ib_[v] = make_float2(a[v], b[v]);
    }

    // output pointer of the first row at plane ks; stores are predicated off below plane ka
    float* outp = A.Qout + (long long)(ks - A.kbase) * P.plane + (long long)i0 * P.Ny + T.jc;
    int xsel = 0;        // exchange buffer of this plane (floats): alternates between 0 and XBUF
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        xch[v * TI * 32 + T.xs] = ib_[v].x;
        xch[(8 + v) * TI * 32 + T.xs] = ib_[v].y;
    }
    int s2 = 2;          // stage of plane k+2
    uint32_t par = 0x3;  // bit s = parity of the next wait on stage s: stages 0,1 were waited once in the prologue
#pragma unroll 1
    for (int k = ks; k < kb; ++k) {
        const int s1 = (s2 + NS - 1) % NS, s0 = (s2 + NS - 2) % NS;
        const float* xq = xch + xsel;   // Qint(k) rows, stored during the previous iteration
        xsel ^= G::XBUF;
        mbar_wait(&full[s2], (par >> s2) & 1u);
        par ^= 1u << s2;
        __syncthreads();
        // stage s0 held plane k: every warp finished reading it (iteration k-1) before the barrier above; refill it two
        // planes ahead of its use
        if (producer && k + NS <= klast) issue(k + NS, s0);
        pair_predict<PATH, TI>(A, T, k + 1 == A.hi_plane, tiles + s1 * G::STAGE_FLOATS, tiles + s2 * G::STAGE_FLOATS, qa, qb, qc, hb, hc,
                               pb, pc, ic);
        {   // publish the rows of Qint(k+1) for the next plane now, among the arithmetic, instead of in front of the barrier:
            // that buffer was last read in iteration k-1, which every warp finished before this iteration's barrier
            float* xn = xch + xsel;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xn[v * TI * 32 + T.xs] = ic[v].x;
                xn[(8 + v) * TI * 32 + T.xs] = ic[v].y;
            }
        }
        pair_correct<PATH, TI>(A, T, k >= ka, xq, qa, ia, ib_, ic, outp);
        outp += P.plane;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            qa[v] = qb[v]; qb[v] = qc[v]; hb[v] = hc[v];
            ia[v] = ib_[v]; ib_[v] = ic[v];
        }
        pb = pc;
        s2 = (s2 + 1) % NS;
    }
}

// Split-phase variant of k_fused_pair: the block-wide __syncthreads per plane becomes an mbarrier pair (arrive right after
// a warp has published its rows of Qint(k+1), wait right before it reads its neighbours' rows of Qint(k) one iteration
// later), with four exchange buffers so that a warp may run up to one plane ahead of the slowest one.
template <int PATH, int TI>
struct SplitGeo : PairGeo<PATH, TI> {
    using B = PairGeo<PATH, TI>;
    static constexpr int NXBUF = 4;
    static constexpr size_t SMEM = B::NSTAGE * B::STAGE_BYTES + NXBUF * B::XBUF * 4 + 64;
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The march of one block of k_fused_split (prologue + plane loop).  W = true handles wall cells and cells beyond the
// domain; W = false would be the body of an interior block, with every wall / edge select folded away at compile time.
template <int PATH, int TI, bool W>
__device__ __forceinline__ void split_march(const FusedArgs& A, const CUtensorMap& tmap, const PairThread& T, float* tiles, float* xch,
                                            uint64_t* full, uint64_t* xbar, int i0, int ib, int c0) {
    using G = SplitGeo<PATH, TI>;
    constexpr int NS = G::NSTAGE;
    const Params& P = A.P;
    const int ti = threadIdx.y;
    (void)ti;
    int ka, kb;
    chunk_range(A, blockIdx.z, ka, kb);
    const bool first = ka == A.ka0;  // the plane below is the slab's qint_lo; otherwise re-derive it in a warm-up plane
    const int ks = first ? ka - 1 : ka - 2;
    const bool producer = (threadIdx.x == 0 && threadIdx.y == 0);
    const int klast = kb + 1;  // last plane any iteration reads

    auto issue = [&](int plane, int stage) {  // producer only
        const int kc = min(max(plane, A.kmin), A.kmax) - A.kbase;
        mbar_expect_tx(&full[stage], G::STAGE_BYTES);
        tma_load_tile(tiles + stage * G::STAGE_FLOATS, &tmap, &full[stage], c0, ib - 1, kc);
    };

    if (producer) {
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
        mbar_init(&xbar[0], TI * 32);
        mbar_init(&xbar[1], TI * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    if (producer)
        for (int s = 0; s < NS; ++s) issue(ks + s, s);

    float2 qa[8], qb[8], qc[8], hb[8], hc[8], ia[8], ib_[8], ic[8];
    PrimT<float2> pb, pc;
    constexpr int VS = G::TR * kTC;
    mbar_wait(&full[0], 0);
    mbar_wait(&full[1], 0);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        qa[v] = make_float2(tiles[v * VS + T.own], tiles[v * VS + T.own + kTC]);
        qb[v] = make_float2(tiles[G::STAGE_FLOATS + v * VS + T.own], tiles[G::STAGE_FLOATS + v * VS + T.own + kTC]);
        ia[v] = make_float2(1.0f, 1.0f);
        ib_[v] = make_float2(1.0f, 1.0f);
    }
    pb = make_prim(qb);
    flux_idx<DIR_Z>(qb, pb, hb);
    if (first) {  // Qint(ka-1); rows beyond the domain read the last row (they never produce output)
        float a[8], b[8];
        ldg8(A.qlo, (long long)min(i0, P.Nx - 1) * P.Ny + T.jc, P.plane, a);
        ldg8(A.qlo, (long long)min(i0 + 1, P.Nx - 1) * P.Ny + T.jc, P.plane, b);
#pragma unroll
        for (int v = 0; v < 8; ++v) ib_[v] = make_float2(a[v], b[v]);
    }

    // output pointer of the first row at plane ks; stores are predicated off below plane ka
    float* outp = A.Qout + (long long)(ks - A.kbase) * P.plane + (long long)i0 * P.Ny + T.jc;
    int xr = 0;          // exchange buffer of this plane (index): plane ks+it reads buffer it & 3 and writes (it+1) & 3
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        xch[v * TI * 32 + T.xs] = ib_[v].x;
        xch[(8 + v) * TI * 32 + T.xs] = ib_[v].y;
    }
    mbar_arrive(&xbar[0]);
    int s2 = 2;          // stage of plane k+2
    uint32_t par = 0x3;  // bit s = parity of the next wait on stage s: stages 0,1 were waited once in the prologue
    uint32_t xpar = 0;   // bit b = parity of the next wait on xbar[b]
#pragma unroll 1
    for (int k = ks; k < kb; ++k) {
        const int s1 = (s2 + NS - 1) % NS, s0 = (s2 + NS - 2) % NS;
        const float* xq = xch + xr * G::XBUF;   // Qint(k) rows, stored during the previous iteration
        const int xb = xr & 1;
        xr = (xr + 1) & 3;
        mbar_wait(&full[s2], (par >> s2) & 1u);
        par ^= 1u << s2;
        pair_predict<PATH, TI, W>(A, T, k + 1 == A.hi_plane, tiles + s1 * G::STAGE_FLOATS, tiles + s2 * G::STAGE_FLOATS, qa, qb, qc, hb, hc,
                               pb, pc, ic);
        {   // publish the rows of Qint(k+1) for the next plane and say so (release): that buffer last held Qint(k-3), read in
            // iteration k-3, and every warp has been seen in iteration k-2 (its arrival for Qint(k-1), which this warp waited on
            // one iteration ago)
            float* xn = xch + xr * G::XBUF;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xn[v * TI * 32 + T.xs] = ic[v].x;
                xn[(8 + v) * TI * 32 + T.xs] = ic[v].y;
            }
            mbar_arrive(&xbar[xb ^ 1]);
        }
        // the neighbours' rows of Qint(k): published during their iteration k-1 -- normally long ago, so the warps of a block
        // drift apart by up to half a plane instead of meeting at a block-wide barrier every plane
        mbar_wait(&xbar[xb], (xpar >> xb) & 1u);
        xpar ^= 1u << xb;
        // every warp has finished the predictor of iteration k-1, the last reader of the tile of plane k: refill its stage
        if (producer && k + NS <= klast) issue(k + NS, s0);
        pair_correct<PATH, TI, W>(A, T, k >= ka, xq, qa, ia, ib_, ic, outp);
        outp += P.plane;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            qa[v] = qb[v]; qb[v] = qc[v]; hb[v] = hc[v];
            ia[v] = ib_[v]; ib_[v] = ic[v];
        }
        pb = pc;
        s2 = (s2 + 1) % NS;
    }
}

template <int PATH, int TI, int MINB = 1>
__global__ void __launch_bounds__(TI * 32, MINB) k_fused_split(const FusedArgs A, const __grid_constant__ CUtensorMap tmap) {
    using G = SplitGeo<PATH, TI>;
    constexpr int NS = G::NSTAGE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                       // [NS][8][TR][kTC]
    float* xch = tiles + NS * G::STAGE_FLOATS;                                // [2][first,last][8][TI][32]  Qint exchange
    uint64_t* full = reinterpret_cast<uint64_t*>(xch + G::NXBUF * G::XBUF);   // [NS] tile landed
    uint64_t* xbar = full + NS;                                               // [2] Qint rows of a plane published (even / odd planes)

    const Params& P = A.P;
    const int lane = threadIdx.x, ti = threadIdx.y;
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int ib = bi * G::WI, jb = bj * G::WJ;
    const int t0 = ti * 2, i0 = ib + t0, j = jb + lane;
    const int oi_lo = bi == 0 ? 0 : ib + 1, oi_hi = bi == A.ntile_i - 1 ? P.Nx : ib + G::WI + 1;
    const int oj_lo = bj == 0 ? 0 : jb + 1, oj_hi = bj == A.ntile_j - 1 ? P.Ny : jb + G::WJ + 1;
    PairThread T;
    T.rows = 0;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int i = i0 + rr, t = t0 + rr;
        const bool corr_row = t >= 1 && (PATH == IMHD_PATH_A || t <= G::NR - 2);  // rows with valid Qint neighbours
        if (i == P.Nx - 1) T.rows |= 1u << rr;
        if (i > 0 && i < P.Nx - 1) T.rows |= 1u << (8 + rr);
        if (corr_row && i > 0 && (PATH == IMHD_PATH_A ? i < P.Nx : i < P.Nx - 1)) T.rows |= 1u << (16 + rr);
        if (i < P.Nx && i >= oi_lo && i < oi_hi) T.rows |= 1u << (24 + rr);
    }
    const bool interior_j = j > 0 && j < P.Ny - 1;
    T.lanes = (j == P.Ny - 1 ? 1u : 0u) | (interior_j ? 2u : 0u) |
              ((PATH == IMHD_PATH_A ? (j > 0 && j < P.Ny) : interior_j) ? 4u : 0u) | ((j < P.Ny && j >= oj_lo && j < oj_hi) ? 8u : 0u);
    // the box starts at the 4-column boundary at or below jb-1 (measured: a misaligned inner coordinate traps)
    const int c0 = ((jb - 1 + 4) / 4) * 4 - 4;
    T.own = (t0 + 1) * kTC + (jb - 1 - c0) + lane + 1;
    T.xs = ti * 32 + lane;
    T.xsm = max(ti - 1, 0) * 32 + lane;
    T.xsp = min(ti + 1, TI - 1) * 32 + lane;
    T.i0 = i0;
    T.jc = min(j, P.Ny - 1);
    keep(T.rows); keep(T.lanes); keep(T.own); keep(T.xs); keep(T.xsm); keep(T.xsp);

    // One instantiation for every block.  A second one with the wall / edge selects folded away for interior blocks
    // (W = false; 82 % of the blocks at 304 x 304) is bit-identical and measures SLOWER (1.604 vs 1.594 ms per step): the
    // 68 FSEL it saves per warp and plane are worth 3.4 % when every block runs that body alone, but two 17 KB bodies in
    // flight on neighbouring SMs cost more in instruction fetch (tools/experiments/README.md).
    split_march<PATH, TI, true>(A, tmap, T, tiles, xch, full, xbar, i0, ib, c0);
}

// -----------------------------------------------------------------------------------------------
// Remainder strip.  The tiles of the hot kernel produce 30 (path A: 31) output columns each; when Ny leaves a
// remainder of a few columns (304 -> 10 x 30 + 2) a whole extra tile column -- 9 % of the launch at 304 -- would
// compute them with 2 of 32 lanes.  This kernel takes the last columns instead, TRANSPOSED: lanes run along i and
// the R = blockDim.y thread rows along j (one row of predictor-only ring, the output columns, and for path B the
// wall column), so a block is 32 x R cells and four such blocks share an SM.  The Q planes come through a rotating
// three-plane staging tile [v][i][8 columns] filled by cp.async with a loader mapping in which 8 consecutive
// threads read the 8 consecutive columns of one i-row (one 32-byte sector).  Same device functions, same inputs
// -> the same bits as any other variant (tests).
// -----------------------------------------------------------------------------------------------
constexpr int kStripCols = 12;  // staging-tile columns: three aligned 4-column groups cover the compute rows, one column either
                                // side and the offset of the first column inside its group
constexpr int kStripRows = 6;   // at most this many compute rows

#ifndef IMHD_STRIP_REGS
#define IMHD_STRIP_REGS 168
#endif
template <int PATH>
__global__ void __maxnreg__(IMHD_STRIP_REGS) k_fused_strip(const FusedArgs A, const __grid_constant__ CUtensorMap tmap) {
    constexpr int O = Ring<PATH>::O;
    constexpr int WL = 32 - 2 * O;
    // The staging tile of a plane is ONE TMA box (12 columns x 32 rows x 8 variables, dense: row pitch 12 floats), four
    // stages with an mbarrier each, filled two planes ahead.  The pitch costs 4-way bank conflicts on the ~24 tile reads
    // per thread and plane: negligible.
    constexpr int PITCH = kStripCols;
    constexpr int NB = 8 * 32 * PITCH;
    extern __shared__ __align__(128) float strip_smem[];
    float* sN = strip_smem;                                            // Q planes k+1 (neighbour reads), k+2 (own read), k+3, k+4 (in flight)
    float* sQi = strip_smem + 4 * NB;                                  // exchange of Qint(k): [2][v][tj][lane]
    uint64_t* full = reinterpret_cast<uint64_t*>(sQi + 2 * 8 * kStripRows * 32);  // [4]

    const Params& P = A.P;
    const int lane = threadIdx.x, tj = threadIdx.y, R = blockDim.y;
    const int bi = blockIdx.x;
    const int i = bi * WL + 1 - O + lane, j = A.jstrip + 1 + tj;   // tile column tj+1
    const bool in_dom = i >= 0 && i < P.Nx && j < P.Ny;
    const int ic = min(max(i, 0), P.Nx - 1), jc = min(j, P.Ny - 1);
    const long long lcol = (long long)ic * P.Ny + jc;
    const bool top = (i == 0), bottom = (i == P.Nx - 1), right = (j == P.Ny - 1);
    const bool interior_ij = in_dom && !top && !bottom && !right;   // j > 0 always: the strip sits at the far wall
    const bool upd = PATH == IMHD_PATH_A ? (in_dom && !top) : interior_ij;
    const int oi_lo = bi == 0 ? 0 : 1 + bi * WL, oi_hi = bi == A.ntile_i - 1 ? P.Nx : 1 + (bi + 1) * WL;
    const bool owner = in_dom && i >= oi_lo && i < oi_hi && tj >= 1;  // row 0 is the predictor-only ring
    const int so = tj * 32 + lane, sjm = max(tj - 1, 0) * 32 + lane, sjp = min(tj + 1, R - 1) * 32 + lane;
    const int c0 = A.jstrip & ~3;                      // first column of the staging tile (16-byte aligned in every row)
    const int st_own = lane * PITCH + (A.jstrip - c0) + tj + 1;

    int ka, kb;
    chunk_range(A, blockIdx.z, ka, kb);
    const bool first = ka == A.ka0;
    const int ks = first ? ka - 1 : ka - 2;
    auto plane_off = [&](int k) -> long long {
        const int kc = min(max(k, A.kmin), A.kmax);
        return (long long)(kc - A.kbase) * P.plane;
    };
    const bool producer = lane == 0 && tj == 0;
    auto fill = [&](int k, int stage) {  // producer only: plane k of the strip into a staging buffer (rows / columns outside
                                         // the domain are zero-filled and only feed lanes whose results are discarded)
        const int kc = min(max(k, A.kmin), A.kmax) - A.kbase;
        mbar_expect_tx(&full[stage], NB * 4);
        tma_load_tile(sN + stage * NB, &tmap, &full[stage], c0, bi * WL + 1 - O, kc);
    };
    if (producer) {
        for (int q = 0; q < 4; ++q) mbar_init(&full[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    // stage of plane p (counted from ks + 1) is p % 4
    if (producer)
        for (int q = 0; q < 3; ++q) fill(ks + 1 + q, q);

    float q0[8], q1[8], h1[8], qim[8], qic[8];
    ldg8(A.Qin + plane_off(ks) + lcol, 0, A.vs, q0);
    ldg8(A.Qin + plane_off(ks + 1) + lcol, 0, A.vs, q1);
    hflux(q1, h1);
#pragma unroll
    for (int v = 0; v < 8; ++v) { qim[v] = 1.0f; qic[v] = 1.0f; }
    if (first) ldg8(A.qlo, lcol, P.plane, qic);

    float* outp = A.Qout + (long long)(ka - A.kbase) * P.plane + lcol;
    int b1 = 0;  // staging buffer of plane k+1; k+2, k+3 follow, k+4 goes to b1+3 (mod 4)
    for (int k = ks; k < kb; ++k) {
        const int buf = (k - ks) & 1;
        const int b2 = (b1 + 1) & 3, b4 = (b1 + 3) & 3;
        float qn[8], hn[8], qip[8];
        const float* bQ = sN + b1 * NB;
        const float* bN = sN + b2 * NB;
        float* bQi = sQi + buf * 8 * kStripRows * 32;
#pragma unroll
        for (int v = 0; v < 8; ++v) bQi[v * kStripRows * 32 + so] = qic[v];
        {   // planes k+1 and k+2 have landed?  plane p (from ks+1) is the (p/4)-th use of stage p%4
            const int p1 = k - ks, p2 = p1 + 1;
            if (k == ks) mbar_wait(&full[p1 & 3], 0);   // later iterations waited for it as their plane k+2
            mbar_wait(&full[p2 & 3], (uint32_t)(p2 >> 2) & 1u);
        }
        __syncthreads();
        if (producer && k + 2 < kb) fill(k + 4, b4);  // two planes ahead; stage b4 held plane k, which nobody reads any more
        b1 = b2;
#pragma unroll
        for (int v = 0; v < 8; ++v) qn[v] = bN[v * 32 * PITCH + st_own];
        hflux(qn, hn);
        if (k + 1 == A.hi_plane) {
            ldg8(A.qhi, lcol, P.plane, qip);
        } else {
            float xp[8], yp[8], xm[8], ym[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xp[v] = __shfl_down_sync(0xffffffffu, q1[v], 1);
                yp[v] = bQ[v * 32 * PITCH + st_own + 1];
            }
            if (PATH == IMHD_PATH_B) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    xm[v] = __shfl_up_sync(0xffffffffu, q1[v], 1);
                    ym[v] = bQ[v * 32 * PITCH + st_own - 1];
                }
            }
            qint_cell<PATH>(q1, xp, yp, h1, hn, xm, ym, q0, qn, bottom, right, false, interior_ij, P, qip);
        }
        if (k >= ka) {
            float out[8], xm[8], ym[8], xp[8], yp[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xm[v] = __shfl_up_sync(0xffffffffu, qic[v], 1);
                ym[v] = bQi[v * kStripRows * 32 + sjm];
            }
            if (PATH == IMHD_PATH_B) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    xp[v] = __shfl_down_sync(0xffffffffu, qic[v], 1);
                    yp[v] = bQi[v * kStripRows * 32 + sjp];
                }
            }
            corr_cell<PATH, true>(q0, qic, xm, ym, qim, xp, yp, qip, P, out);
            if (owner) {
#pragma unroll
                for (int v = 0; v < 8; ++v) outp[v * A.vs] = upd ? out[v] : q0[v];
            }
            outp += P.plane;
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            q0[v] = q1[v]; q1[v] = qn[v]; h1[v] = hn[v];
            qim[v] = qic[v]; qic[v] = qip[v];
        }
    }
}

// -----------------------------------------------------------------------------------------------
// Remainder strip, warp-autonomous (the one the launcher prefers when the strip is at most four columns wide).
// One WARP owns a 16-row x 4-column tile of the strip -- lane = 4 * r + c: thread row r = 0..7 (two rows each, packed as
// in k_fused_pair), column c = 0..3 (c = 0 is the predictor-only ring column = the last output column of the marching
// kernel) -- and marches along k on its own: no shared memory, no block barrier, no TMA.  Everything a cell needs from
// its neighbours inside the tile travels by warp shuffle (rows: +-4 lanes, columns: +-1 lane): the neighbour STATES of
// Q(k+1) and Qint(k), and the neighbour FLUXES G(j+-1), F(i+1), F'(i-1), which the neighbouring lane evaluates as its
// own anyway; the one row above / below the tile and the one column left of it are loaded by the lanes at that edge.
// Four lanes read 16 contiguous bytes of a row, so a load instruction touches 8 sectors instead of the 32 of the
// lanes-along-i strip; the loads go straight to registers and their latency is hidden by the other warps of the SM
// (eight independent warps).  Same device functions on the same values -> the same bits (tests).
// -----------------------------------------------------------------------------------------------
__device__ __forceinline__ float shdn(float x, int d) { return __shfl_down_sync(0xffffffffu, x, d); }
__device__ __forceinline__ float shup(float x, int d) { return __shfl_up_sync(0xffffffffu, x, d); }
__device__ __forceinline__ float2 shdn2(float2 x, int d) { return make_float2(shdn(x.x, d), shdn(x.y, d)); }
__device__ __forceinline__ float2 shup2(float2 x, int d) { return make_float2(shup(x.x, d), shup(x.y, d)); }

constexpr int kWsCols = 5;                    // staged columns: the one left of the tile + the tile's four
constexpr int kWsRows = 18;                   // staged rows: the tile's sixteen + one above + one below
constexpr int kWsVar = kWsRows * kWsCols;     // variable stride inside a stage
constexpr int kWsStage = 8 * kWsVar;          // floats per staged plane
constexpr int kWsNst = 4;                     // planes k+1, k+2 in use, k+3, k+4 in flight
constexpr size_t kWsSmem = 4 * kWsNst * kWsStage * sizeof(float);   // four warps per block

template <int PATH>
__global__ void __launch_bounds__(128, 2) k_fused_wstrip(const FusedArgs A) {
    constexpr int NR = 16;
    constexpr int WI = PATH == IMHD_PATH_A ? NR - 1 : NR - 2;
    extern __shared__ __align__(16) float ws_smem[];
    const Params& P = A.P;
    const int lane = threadIdx.x & 31;
    float* ring = ws_smem + (threadIdx.x >> 5) * (kWsNst * kWsStage);   // this warp's private staging ring
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp tile: tile row bi, z chunk cz
    const int bi = w % A.ntile_i, cz = w / A.ntile_i;
    int ka, kb;
    chunk_range(A, cz, ka, kb);
    if (ka >= kb) return;                                    // warp-uniform: a warp beyond the last chunk
    const int r = lane >> 2, c = lane & 3;
    const int ib = bi * WI, t0 = 2 * r, i0 = ib + t0, j = A.jstrip + 1 + c;
    const int oi_lo = bi == 0 ? 0 : ib + 1, oi_hi = bi == A.ntile_i - 1 ? P.Nx : ib + WI + 1;
    PairThread T;
    T.rows = 0;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int i = i0 + rr, t = t0 + rr;
        const bool corr_row = t >= 1 && (PATH == IMHD_PATH_A || t <= NR - 2);
        if (i == P.Nx - 1) T.rows |= 1u << rr;
        if (i > 0 && i < P.Nx - 1) T.rows |= 1u << (8 + rr);
        if (corr_row && i > 0 && (PATH == IMHD_PATH_A ? i < P.Nx : i < P.Nx - 1)) T.rows |= 1u << (16 + rr);
        if (i < P.Nx && i >= oi_lo && i < oi_hi) T.rows |= 1u << (24 + rr);
    }
    const bool interior_j = j > 0 && j < P.Ny - 1;
    T.lanes = (j == P.Ny - 1 ? 1u : 0u) | (interior_j ? 2u : 0u) |
              ((PATH == IMHD_PATH_A ? (j > 0 && j < P.Ny) : interior_j) ? 4u : 0u) | ((j < P.Ny && c >= 1) ? 8u : 0u);
    T.i0 = i0;
    T.jc = min(j, P.Ny - 1);
    T.own = (t0 + 1) * kWsCols + c + 1;                      // first row of the thread inside a staged plane
    keep(T.rows); keep(T.lanes); keep(T.own);
    const int ic0 = min(i0, P.Nx - 1), ic1 = min(i0 + 1, P.Nx - 1);
    const FlagT<float2> right = {T.right(), T.right()};

    const bool first = ka == A.ka0;
    const int ks = first ? ka - 1 : ka - 2;
    const int klast = kb + 1;  // last plane any iteration reads

    // Loader: lanes 0..29 copy 6 rows x 5 columns of one variable per pass (three passes cover the 18 staged rows), four
    // bytes per cp.async; rows / columns outside the domain repeat the edge (they only feed discarded results).
    int goff[3];
    {
        const int lr = lane / kWsCols, lc = lane % kWsCols;
#pragma unroll
        for (int p = 0; p < 3; ++p)
            goff[p] = min(max(ib - 1 + p * 6 + lr, 0), P.Nx - 1) * P.Ny + min(A.jstrip + lc, P.Ny - 1);
    }
    const uint32_t ring_u32 = smem_u32(ring) + 4u * lane;
    auto fill = [&](int k) {   // plane k -> stage (k - ks) & 3; always one commit group
        if (k <= klast && lane < 30) {
            const float* src = A.Qin + (long long)(min(max(k, A.kmin), A.kmax) - A.kbase) * P.plane;
            const uint32_t dst = ring_u32 + 4u * (((k - ks) & (kWsNst - 1)) * kWsStage);
#pragma unroll
            for (int v = 0; v < 8; ++v)
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * (v * kWsVar + p * 30)),
                                 "l"(src + v * A.vs + goff[p])
                                 : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto stage = [&](int k) -> const float* { return ring + ((k - ks) & (kWsNst - 1)) * kWsStage; };

    for (int q = 0; q < kWsNst; ++q) fill(ks + q);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncwarp();

    float2 qa[8], qb[8], qc[8], hb[8], hc[8], ia[8], ib_[8], ic[8];
    PrimT<float2> pb, pc;
    {
        const float* sa = stage(ks);
        const float* sb = stage(ks + 1);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            qa[v] = make_float2(sa[v * kWsVar + T.own], sa[v * kWsVar + T.own + kWsCols]);
            qb[v] = make_float2(sb[v * kWsVar + T.own], sb[v * kWsVar + T.own + kWsCols]);
            ia[v] = make_float2(1.0f, 1.0f);
            ib_[v] = make_float2(1.0f, 1.0f);
        }
    }
    pb = make_prim(qb);
    flux_idx<DIR_Z>(qb, pb, hb);
    if (first) {
        float a[8], b[8];
        ldg8(A.qlo, (long long)ic0 * P.Ny + T.jc, P.plane, a);
        ldg8(A.qlo, (long long)ic1 * P.Ny + T.jc, P.plane, b);
#pragma unroll
        for (int v = 0; v < 8; ++v) ib_[v] = make_float2(a[v], b[v]);
    }
    float* outp = A.Qout + (long long)(ks - A.kbase) * P.plane + (long long)i0 * P.Ny + T.jc;

#pragma unroll 1
    for (int k = ks; k < kb; ++k) {
        // plane k+2 has landed (the two fills behind it may still be in flight); every lane is past its reads of plane k
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        fill(k + kWsNst);
        const float* tq1 = stage(k + 1);
        const float* tq2 = stage(k + 2);
        // ---- predictor of plane k+1 -------------------------------------------------------------------------
#pragma unroll
        for (int v = 0; v < 8; ++v) qc[v] = make_float2(tq2[v * kWsVar + T.own], tq2[v * kWsVar + T.own + kWsCols]);
        pc = make_prim(qc);
        flux_idx<DIR_Z>(qc, pc, hc);
        {
            float xlast[8], xfirst[8], fl[8];
            float2 yp[8], ym[8], f[8], g[8], gy[8], dF[8], xsum[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                xlast[v] = tq1[v * kWsVar + T.own + 2 * kWsCols];       // Q(k+1) one row below the thread's rows
                yp[v] = make_float2(tq1[v * kWsVar + T.own + 1], tq1[v * kWsVar + T.own + kWsCols + 1]);
                if (PATH == IMHD_PATH_B) {
                    xfirst[v] = tq1[v * kWsVar + T.own - kWsCols];      // ... and one row above
                    ym[v] = make_float2(tq1[v * kWsVar + T.own - 1], tq1[v * kWsVar + T.own + kWsCols - 1]);
                }
            }
            flux_idx<DIR_X>(qb, pb, f);
            flux_idx<DIR_X>(xlast, make_prim(xlast), fl);
            flux_idx<DIR_Y>(qb, pb, g);
            // G(Q(k+1)) one column right: the lane to the right has just evaluated it as its own (the tile's last column never
            // needs it: it is the wall or outside the domain)
#pragma unroll
            for (int v = 0; v < 8; ++v) gy[v] = v == RHO ? yp[MY] : (v == BY ? make_float2(0.0f, 0.0f) : shdn2(g[v], 1));
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                dF[v].x = T.bottom(0) ? -f[v].x : f[v].y - f[v].x;
                dF[v].y = T.bottom(1) ? -f[v].y : fl[v] - f[v].y;
                if (PATH == IMHD_PATH_B) xsum[v] = make_float2(qb[v].y + xfirst[v], xlast[v] + qb[v].x);
            }
            qint_combine<PATH, float2>(qb, dF, g, gy, hb, hc, xsum, ym, yp, qa, qc, {T.bottom(0), T.bottom(1)}, right,
                                       {T.interior_i(0) && T.interior_j(), T.interior_i(1) && T.interior_j()}, P, ic);
        }
        if (__builtin_expect(k + 1 == A.hi_plane, 0)) {
            float a[8], b[8];
            ldg8(A.qhi, (long long)ic0 * P.Ny + T.jc, P.plane, a);
            ldg8(A.qhi, (long long)ic1 * P.Ny + T.jc, P.plane, b);
#pragma unroll
            for (int v = 0; v < 8; ++v) ic[v] = make_float2(a[v], b[v]);
        }
        // ---- corrector of plane k ---------------------------------------------------------------------------
        {
            float xfirst[8], xlast[8];
            float2 fc[8], gc[8], hcc[8], gj[8], hk[8], ym[8], yp[8], dF[8], xsum[8], out[8];
            const PrimT<float2> sc = make_prim(ib_);
            flux_loc<DIR_X>(ib_, sc, fc);
            flux_loc<DIR_Y>(ib_, sc, gc);
            flux_loc<DIR_Z>(ib_, sc, hcc);
#pragma unroll
            for (int v = 0; v < 8; ++v) xfirst[v] = shup(ib_[v].y, 4);   // Qint(k), second row of the thread row above
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                if (PATH == IMHD_PATH_B) xlast[v] = shdn(ib_[v].x, 4);
                ym[v] = shup2(ib_[v], 1);
                if (PATH == IMHD_PATH_B) yp[v] = shdn2(ib_[v], 1);
                gj[v] = v == BY ? make_float2(0.0f, 0.0f) : shup2(gc[v], 1);
                // F'(Qint(i-1)): the thread row above evaluated it as the flux of its second row (B-6: rho)
                const float ff = (v == RHO) ? xfirst[MX] : (v == BX ? 0.0f : shup(fc[v].y, 4));
                const float fa = (PATH == IMHD_PATH_B && v == RHO) ? xfirst[RHO] : ff;
                const float fb = (PATH == IMHD_PATH_B && v == RHO) ? ib_[RHO].x : fc[v].x;
                dF[v] = make_float2(fc[v].x - fa, fc[v].y - fb);
                if (PATH == IMHD_PATH_B) xsum[v] = make_float2(ib_[v].y + xfirst[v], xlast[v] + ib_[v].x);
            }
            hflux_km1<float2>(ia, make_float2(xfirst[BX], ib_[BX].x), ym[BY], ym[MY], hk);
            corr_combine<PATH, float2>(qa, ib_, dF, gc, gj, hcc, hk, xsum, ym, yp, ia, ic, P, out);
            if (k >= ka && T.owner_j()) {
                const bool ua = T.upd_i(0) && T.upd_j(), ub = T.upd_i(1) && T.upd_j();
                if (T.owner_i(0)) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) outp[v * A.vs] = ua ? out[v].x : qa[v].x;
                }
                if (T.owner_i(1)) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) outp[v * A.vs + P.Ny] = ub ? out[v].y : qa[v].y;
                }
            }
        }
        outp += P.plane;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            qa[v] = qb[v]; qb[v] = qc[v]; hb[v] = hc[v];
            ia[v] = ib_[v]; ib_[v] = ic[v];
        }
        pb = pc;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Predictor plane k (global index, k <= Nz-2) into an (8,Nx,Ny) buffer: same device function, same
// values as the fused kernel computes on chip.
template <int PATH>
__device__ __forceinline__ void qint_at(const FusedArgs& A, int i, int j, int k, float r[8]) {
    const Params& P = A.P;
    const bool bottom = i == P.Nx - 1, right = j == P.Ny - 1;
    const bool lap = PATH == IMHD_PATH_B && i > 0 && !bottom && j > 0 && !right && k >= 1;
    auto off = [&](int ii, int jj, int kk) -> long long {
        ii = min(max(ii, 0), P.Nx - 1); jj = min(max(jj, 0), P.Ny - 1); kk = min(max(kk, A.kmin), A.kmax);
        return (long long)(kk - A.kbase) * P.plane + (long long)ii * P.Ny + jj;
    };
    float c[8], xp[8], yp[8], zp[8], xm[8], ym[8], zm[8], hc[8], hp[8];
    ldg8(A.Qin, off(i, j, k), A.vs, c);
    ldg8(A.Qin, off(i + 1, j, k), A.vs, xp);
    ldg8(A.Qin, off(i, j + 1, k), A.vs, yp);
    ldg8(A.Qin, off(i, j, k + 1), A.vs, zp);
    ldg8(A.Qin, off(i - 1, j, k), A.vs, xm);
    ldg8(A.Qin, off(i, j - 1, k), A.vs, ym);
    ldg8(A.Qin, off(i, j, k - 1), A.vs, zm);
    hflux(c, hc);
    hflux(zp, hp);
    qint_cell<PATH>(c, xp, yp, hc, hp, xm, ym, zm, zp, bottom, right, k == 0, lap, P, r);
}

template <int PATH>
__global__ void __launch_bounds__(256) k_qint_plane(const FusedArgs A, int k, float* __restrict__ out) {
    const Params& P = A.P;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= P.Nx || j >= P.Ny) return;
    float r[8];
    qint_at<PATH>(A, i, j, k, r);
    const long long l = (long long)i * P.Ny + j;
#pragma unroll
    for (int v = 0; v < 8; ++v) out[l + v * P.plane] = r[v];
}

// Path B, plane 0 of the new state: BoundaryConditions (kernels_fluidbcs.cu:32-235), one thread per (i,j).
//   interior: corrector with indexed fluxes of Qint(0), k-1 -> Nz-2 (qwrap), + dt*numericalDiffusionFront
//   i = 0, Nx-1: wall values; j = 0, Ny-1 (i interior): carried over (dead code in the reference, B-8)
__global__ void __launch_bounds__(256) k_front_plane_B(const FusedArgs A) {
    const Params& P = A.P;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= P.Nx || j >= P.Ny) return;
    const long long l = (long long)i * P.Ny + j;
    const long long o = (long long)(0 - A.kbase) * P.plane + l;
    float q[8], out[8];
    ldg8(A.Qin, o, A.vs, q);
#pragma unroll
    for (int v = 0; v < 8; ++v) out[v] = q[v];
    if (i > 0 && i < P.Nx - 1) {
        if (j > 0 && j < P.Ny - 1) {
            float c[8], xm[8], ym[8], zm[8], xp[8], yp[8], zp[8];
            ldg8(A.qlo, l, P.plane, c);           // Qint(0)
            ldg8(A.qlo, l - P.Ny, P.plane, xm);
            ldg8(A.qlo, l - 1, P.plane, ym);
            ldg8(A.qlo, l + P.Ny, P.plane, xp);
            ldg8(A.qlo, l + 1, P.plane, yp);
            ldg8(A.qwrap, l, P.plane, zm);        // Qint(Nz-2)
            qint_at<IMHD_PATH_B>(A, i, j, 1, zp);  // Qint(1)
            front_cell_exact(q, c, xm, ym, zm, xp, yp, zp, P, out);
        }
    } else {
        out[RHO] = 1.0f;
#pragma unroll
        for (int v = 1; v < 7; ++v) out[v] = 0.0f;
        float e = q[EN];
        for (int rep = 0; rep < P.Nx; ++rep) {
            const float e2 = wall_e(e);
            if (e2 == e) break;
            e = e2;
        }
        out[EN] = e;
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) A.Qout[o + v * A.vs] = out[v];
}

// Path B, plane Nz-1 of the new state: carried over, except the column (Nx-1,Ny-1) which receives the wall
// value of (Nx-1,Ny-1,0) (kernels_fluidbcs.cu:227-231, B-8).
__global__ void __launch_bounds__(256) k_back_plane_B(const FusedArgs A) {
    const Params& P = A.P;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.plane) return;
    const long long o = (long long)(P.Nz - 1 - A.kbase) * P.plane + c;
    const bool corner = c == P.plane - 1;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        float x = __ldg(A.Qin + o + v * A.vs);
        if (corner) x = v == RHO ? 1.0f : (v == EN ? A.corner_e : 0.0f);
        A.Qout[o + v * A.vs] = x;
    }
}

// dst plane <- src plane of the same array, 8 variables (path A PBCs on the new state)
__global__ void __launch_bounds__(256) k_plane_copy(float* __restrict__ Q, long long dst, long long src, long long plane, long long vs) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plane) return;
#pragma unroll
    for (int v = 0; v < 8; ++v) Q[dst + c + v * vs] = Q[src + c + v * vs];
}

}  // namespace imhd

using namespace imhd;

static int fill_args(FusedArgs& A, const float* Qin, float* Qout, const float* qlo, const float* qhi,
                     const float* qwrap, const imhd_slab* s) {
    if (!s) { set_error("null slab descriptor"); return IMHD_E_INVALID; }
    if (int e = bad_dims(s->Nx, s->Ny, s->Nz)) return e;
    if (s->path != IMHD_PATH_A && s->path != IMHD_PATH_B) { set_error("bad path %d", s->path); return IMHD_E_INVALID; }
    if (s->k0 < 0 || s->nzl < 3 || s->k0 + s->nzl > s->Nz) {
        set_error("bad slab: k0=%d nzl=%d Nz=%d (need nzl >= 3)", s->k0, s->nzl, s->Nz);
        return IMHD_E_INVALID;
    }
    if (!s->ghosts && (s->k0 != 0 || s->nzl != s->Nz)) {
        set_error("a slab that is not the whole domain needs ghosts=1");
        return IMHD_E_INVALID;
    }
    A.P = make_params(s->path, s->D, s->dt, s->dx, s->dy, s->dz, s->Nx, s->Ny, s->Nz);
    A.Qin = Qin; A.Qout = Qout;
    const int g = s->ghosts ? 1 : 0;
    A.vs = (long long)(s->nzl + 2 * g) * A.P.plane;
    if (A.vs >= (1ll << 32)) { set_error("slab of %lld cells per variable exceeds the 2^32 addressing range", A.vs); return IMHD_E_INVALID; }
    A.vs32 = (unsigned)A.vs;
    A.kbase = s->k0 - g;
    A.kmin = max(s->k0 - g, 0);
    A.kmax = min(s->k0 + s->nzl - 1 + g, s->Nz - 1);
    A.k0 = s->k0; A.k1 = s->k0 + s->nzl;
    A.ka0 = max(A.k0, 1);
    A.kb0 = min(A.k1, s->path == IMHD_PATH_A ? s->Nz : s->Nz - 1);
    A.kfrom = A.ka0; A.kto = A.kb0;
    A.qlo = qlo; A.qhi = qhi; A.qwrap = qwrap;
    A.hi_plane = min(A.k1, s->Nz - 1);
    A.corner_e = s->corner_e;
    A.chunk = A.clen = 0; A.ntile_i = A.ntile_j = 0; A.jstrip = 0;
    return 0;
}

// block_rows: rows of 32 threads per block.  8 by default; the slab engine asks for 4 (128 threads x 80 registers) so that the
// blocks find room beside a resident block of the marching kernel instead of waiting for an SM of their own.
int imhd_qint_plane_rows(const float* Q, float* out_plane, int k, const imhd_slab* s, int block_rows, void* stream) {
    FusedArgs A;
    if (int e = fill_args(A, Q, nullptr, nullptr, nullptr, nullptr, s)) return e;
    if (k < 0 || k > s->Nz - 2 || k < A.kmin || k + 1 > A.kmax) {
        set_error("imhd_qint_plane: plane %d not computable from array planes [%d,%d]", k, A.kmin, A.kmax);
        return IMHD_E_INVALID;
    }
    const dim3 grid((s->Ny + 31) / 32, (s->Nx + block_rows - 1) / block_rows), block(32, block_rows);
    if (s->path == IMHD_PATH_A) k_qint_plane<IMHD_PATH_A><<<grid, block, 0, (cudaStream_t)stream>>>(A, k, out_plane);
    else                        k_qint_plane<IMHD_PATH_B><<<grid, block, 0, (cudaStream_t)stream>>>(A, k, out_plane);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

extern "C" int imhd_qint_plane(const float* Q, float* out_plane, int k, const imhd_slab* s, void* stream) {
    return imhd_qint_plane_rows(Q, out_plane, k, s, 8, stream);
}

extern "C" float imhd_wall_energy_fixed_point(float e, int max_iter) {
    for (int rep = 0; rep < max_iter; ++rep) {  // kernels_fluidbcs.cu:173 (B-12), host copy of imhd::wall_e
        const float p = (float)(kGm1 * ((e - 0.0f) - 0.0f / 2.0));
        const float e2 = (float)(p / kGm1);
        if (e2 == e) break;
        e = e2;
    }
    return e;
}

static int g_chunk_override = 0;
extern "C" void imhd_set_chunk(int planes) { g_chunk_override = planes; }

// ---- TMA descriptor ----------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled get_encode() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled)p;
    }
    return fn;
}

static int g_force_ldg = 0, g_no_strip = 0, g_force_strip = 0, g_no_under = 0, g_kernel = 0, g_block_strip = 0;
extern "C" void imhd_set_kernel_variant(int flags) {
    g_force_ldg = flags & 1; g_no_strip = (flags >> 1) & 1; g_force_strip = (flags >> 2) & 1; g_no_under = (flags >> 3) & 1;
    g_kernel = (flags >> 4) & 15; g_block_strip = (flags >> 8) & 1;
}

// ---- work that runs UNDER the path B marching kernel ------------------------------------------------------------------
// A block of the path B marching kernel (8 warps x 208 registers, one block per SM) leaves 12288 registers, ~60 KB of shared
// memory and -- being bound by its own dependency latency -- issue slots of its SM idle.  The small launches of a step that
// depend only on the OLD state (the remainder strip: a latency chain of its own; the k = 0 face and the k = Nz-1 plane of
// path B) are therefore forked onto a per-device side stream in blocks small enough to be co-resident with a marching block
// (one 32-thread strip block of 228 registers per SM; 64-thread face blocks) and joined behind the marching kernel: they
// cost the step their share of the issue slots instead of their own latency (0.12 ms of 1.64 at 304x304x592).
namespace {
struct UnderStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
UnderStream g_under[64];
std::mutex g_under_mu;   // a fork or join is an event record + a stream wait on that event: one atomic pair per host thread
}  // namespace

static UnderStream* under_stream() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> g(g_under_mu);
    UnderStream* u = &g_under[dev & 63];
    if (!u->s) {
        if (cudaStreamCreateWithFlags(&u->s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&u->fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&u->join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            u->s = nullptr;
            return nullptr;
        }
    }
    return u;
}
// `to` continues after everything enqueued on `from` so far
static int stream_follows(cudaStream_t to, cudaStream_t from, cudaEvent_t ev) {
    std::lock_guard<std::mutex> g(g_under_mu);
    IMHD_CUDA(cudaEventRecord(ev, from));
    IMHD_CUDA(cudaStreamWaitEvent(to, ev, 0));
    return 0;
}

// 4-D view (j, i, plane, variable) of a state array for the tile loads of the TMA kernel.
static bool make_tile_map(CUtensorMap* map, const FusedArgs& A, int nplanes, int tile_rows, int tile_cols = kTC,
                          CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B) {
    const Params& P = A.P;
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc || g_force_ldg) return false;
    if (P.Ny % 4 != 0 || ((uintptr_t)A.Qin & 15) != 0) return false;  // TMA: 16-byte base and strides
    const cuuint64_t dims[4] = {(cuuint64_t)P.Ny, (cuuint64_t)P.Nx, (cuuint64_t)nplanes, 8};
    const cuuint64_t strides[3] = {(cuuint64_t)P.Ny * 4, (cuuint64_t)P.plane * 4, (cuuint64_t)A.vs * 4};
    const cuuint32_t box[4] = {(cuuint32_t)tile_cols, (cuuint32_t)tile_rows, 1, 8};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)A.Qin, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// cudaFuncSetAttribute is per device: remember which devices already have the opt-in shared-memory size set
template <class K>
static int ensure_smem(K kernel, size_t bytes, unsigned long long& done_mask) {
    int dev = 0;
    IMHD_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done_mask & bit)) {
        IMHD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        done_mask |= bit;
    }
    return 0;
}

// A launch over two plane ranges arrives with chunk (the distance between the ranges' first planes) and clen (their common
// length) set by step_ranges and runs one z-chunk per range; otherwise the launcher picks the chunk length.
static bool two_ranges(const FusedArgs& A) { return A.clen > 0 && A.clen < A.chunk; }
static int longest_range(const FusedArgs& A) { return two_ranges(A) ? A.clen : A.kto - A.kfrom; }
static int count_chunks(const FusedArgs& A) { return two_ranges(A) ? 2 : (A.kto - A.kfrom + A.chunk - 1) / A.chunk; }

static int pick_chunk(FusedArgs& A, int nz, long long tiles = 0, int blocks_per_sm = 1) {
    if (two_ranges(A)) return 2;
    // z-chunks: one block per SM is resident (register-limited), so the launch runs in waves of `sms` blocks.  Pick the
    // chunk count that minimises  waves x (chunk length + warm-up)  -- i.e. fill the last wave -- with chunks long
    // enough to amortise the two warm-up planes (each costs about half a plane).
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sms *= blocks_per_sm;   // resident blocks per wave
    if (tiles <= 0) tiles = (long long)A.ntile_i * A.ntile_j;
    int best_n = 1;
    double best = 1e300;
    const int nmax = nz / 12 > 1 ? nz / 12 : 1;
    for (int n = 1; n <= nmax; ++n) {
        const int len = (nz + n - 1) / n;
        const long long blocks = tiles * ((nz + len - 1) / len);
        const double cost = (double)((blocks + sms - 1) / sms) * (len + 1.0);
        if (cost < best * 0.999) { best = cost; best_n = n; }
    }
    int chunk = (nz + best_n - 1) / best_n;
    if (g_chunk_override > 0) chunk = g_chunk_override;
    if (chunk < 2) chunk = 2;
    if (chunk > nz) chunk = nz;
    A.chunk = A.clen = chunk;
    return (nz + chunk - 1) / chunk;
}

// The one-row kernel and the register-tiled kernels share the launch logic: tile counts from the geometry, the
// remainder strip, wave-aware z-chunks.
template <int PATH, int TI>
struct OneRowLaunch {
    using G = TmaGeo<PATH, TI>;
    static constexpr int THREAD_ROWS = TI;
    static constexpr int BLOCKS_PER_SM = 1;
    static auto kernel() { return k_fused_step_tma<PATH, TI>; }
};
template <int PATH, int TI, int MINB = 1>
struct PairLaunch {
    using G = PairGeo<PATH, TI>;
    static constexpr int THREAD_ROWS = TI;
    static constexpr int BLOCKS_PER_SM = MINB;
    static auto kernel() { return k_fused_pair<PATH, TI, MINB>; }
};

template <int PATH, int TI, int MINB = 1>
struct SplitLaunch {
    using G = SplitGeo<PATH, TI>;
    static constexpr int THREAD_ROWS = TI;
    static constexpr int BLOCKS_PER_SM = MINB;
    static auto kernel() { return k_fused_split<PATH, TI, MINB>; }
};
// ---- optional per-launch timing of the hot kernel (bench.py's roofline leg) ---------------------------------------------
// When enabled, every launch of the marching kernel that covers >= 64 planes (+ its remainder strip) is bracketed by
// a CUDA event pair on the launching stream; imhd_fused_timing_read sums them after the caller has synchronised.
namespace {
struct TimedLaunch { cudaEvent_t a, b; int dev; long long cells; };
std::vector<TimedLaunch>* g_timed = nullptr;
std::mutex g_timed_mu;
bool g_timing = false;
}  // namespace

extern "C" void imhd_fused_timing(int enable) {
    std::lock_guard<std::mutex> g(g_timed_mu);
    if (!g_timed) { g_timed = new std::vector<TimedLaunch>(); g_timed->reserve(100000); }  // pointers into it stay valid
    for (TimedLaunch& t : *g_timed) { cudaSetDevice(t.dev); cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    g_timed->clear();
    g_timing = enable != 0;
}

// total_ms / launches / cell_updates of the bracketed launches since imhd_fused_timing(1); synchronise the streams first
extern "C" int imhd_fused_timing_read(double* total_ms, int* launches, long long* cell_updates) {
    std::lock_guard<std::mutex> g(g_timed_mu);
    double ms = 0.0;
    long long cells = 0;
    int n = 0;
    if (g_timed)
        for (TimedLaunch& t : *g_timed) {
            float x = 0.f;
            cudaSetDevice(t.dev);
            if (cudaEventElapsedTime(&x, t.a, t.b) != cudaSuccess) { cudaGetLastError(); continue; }
            ms += x; cells += t.cells; ++n;
        }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    if (cell_updates) *cell_updates = cells;
    return 0;
}

static TimedLaunch* timing_begin(const FusedArgs& A, int nz, cudaStream_t st) {
    if (!g_timing || nz < 64) return nullptr;
    std::lock_guard<std::mutex> g(g_timed_mu);
    if (!g_timed || g_timed->size() >= 100000) return nullptr;
    TimedLaunch t;
    cudaGetDevice(&t.dev);
    t.cells = (long long)A.P.Nx * A.P.Ny * nz;
    if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return nullptr;
    cudaEventRecord(t.a, st);
    g_timed->push_back(t);
    return &g_timed->back();
}
static void timing_end(TimedLaunch* t, cudaStream_t st) {
    if (t) cudaEventRecord(t->b, st);
}

template <int PATH, class L>
static int launch_tma(FusedArgs& A, const CUtensorMap& tmap, int nplanes_array, cudaStream_t st, UnderStream* under) {
    using G = typename L::G;
    const Params& P = A.P;
    const int nz = longest_range(A);
    const int ni = PATH == IMHD_PATH_A ? P.Nx - 1 : P.Nx - 2, nj = PATH == IMHD_PATH_A ? P.Ny - 1 : P.Ny - 2;
    A.ntile_i = (ni + G::WI - 1) / G::WI;
    A.ntile_j = (nj + G::WJ - 1) / G::WJ;
    // a remainder of a few columns goes to the transposed strip kernel instead of a whole extra tile column
    constexpr int O = Ring<PATH>::O;
    const int rem = nj % G::WJ;
    const int strip_rows = PATH == IMHD_PATH_A ? 1 + rem : 1 + rem + 1;  // predictor-only ring row, outputs (, wall column)
    // (not for short plane ranges such as the slab-end launches of the multi-GPU loop: there the strip's fixed
    // latency costs more than the extra tile column; the test hook bit 2 forces it on for any length)
    const bool strip = !g_no_strip && rem > 0 && strip_rows <= kStripRows && nj / G::WJ >= 2 && (nz >= 64 || g_force_strip);
    int grid_j = A.ntile_j;
    if (strip) {
        grid_j = nj / G::WJ;
        A.ntile_j = grid_j + 1;  // no hot-kernel tile is the last one: none widens its window to the domain edge
    }
    pick_chunk(A, nz, (long long)A.ntile_i * grid_j, L::BLOCKS_PER_SM);
    const int nchunk = count_chunks(A);
    static unsigned long long done = 0;
    if (int e = ensure_smem(L::kernel(), G::SMEM, done)) return e;
    TimedLaunch* timed = timing_begin(A, nz, st);
    // The fork goes IMMEDIATELY in front of the marching kernel: the side stream's work must reach the block scheduler after
    // the marching blocks have taken their SMs and fill what they leave.  With anything between the fork and the launch (the
    // timing event of the roofline leg was enough) the one-warp strip blocks get there first, eight to an SM, and the marching
    // blocks wait for them: the strip runs IN FRONT of the kernel instead of under it (measured: 1.58 instead of 1.47 ms).
    if (under)
        if (int e = stream_follows(under->s, st, under->fork)) return e;
    L::kernel()<<<dim3(grid_j, A.ntile_i, nchunk), dim3(32, L::THREAD_ROWS), G::SMEM, st>>>(A, tmap);
    IMHD_LAUNCH_CHECK(1);
    if (strip && strip_rows <= 4 && !g_block_strip) {
        // warp-autonomous strip: one warp per 16-row x 4-column tile and z chunk, eight warps per SM
        FusedArgs S = A;
        constexpr int WIw = PATH == IMHD_PATH_A ? 15 : 14;
        S.ntile_i = (ni + WIw - 1) / WIw;
        S.jstrip = grid_j * G::WJ - 1;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (!two_ranges(A)) {   // (two ranges: one chunk per range, as the marching kernel)
            // on its own stream: one wave of warps, two 4-warp blocks per SM (230 registers).  Under the marching kernel one
            // warp per SM is resident and the strip has the whole launch to finish: half as many, longer chunks (fewer warm-up
            // planes; measured 37.0 / 37.2 / 37.1 / 36.9 / 36.8 GLUPS for 3 / 4 / 6 / 8 / 16 warps' worth per SM)
            const int wpsm = under && L::BLOCKS_PER_SM == 1 ? 4 : 8;
            int n = (wpsm * sms) / S.ntile_i;
            n = n > nz / 8 ? nz / 8 : n;
            n = n < 1 ? 1 : n;
            S.chunk = (nz + n - 1) / n;
            if (g_chunk_override > 0) S.chunk = g_chunk_override;
            S.chunk = S.chunk < 2 ? 2 : (S.chunk > nz ? nz : S.chunk);
            S.clen = S.chunk;
        }
        const int warps = S.ntile_i * count_chunks(S);
        static unsigned long long wdone = 0;
        if (int e = ensure_smem(k_fused_wstrip<PATH>, kWsSmem, wdone)) return e;
        if (under && L::BLOCKS_PER_SM == 1) {
            // one-warp blocks on the side stream, launched after the marching kernel: one fits beside each marching block
            k_fused_wstrip<PATH><<<warps, 32, kWsSmem / 4, under->s>>>(S);
            IMHD_LAUNCH_CHECK(1);
            if (timed) {   // the bracket of the roofline leg covers the strip
                if (int e = stream_follows(st, under->s, under->join)) return e;
            }
        } else {
            k_fused_wstrip<PATH><<<(warps + 3) / 4, 128, kWsSmem, st>>>(S);
            IMHD_LAUNCH_CHECK(1);
        }
    } else if (strip) {
        FusedArgs S = A;
        constexpr int WL = 32 - 2 * O;
        S.ntile_i = (ni + WL - 1) / WL;
        S.jstrip = grid_j * G::WJ - 1;          // tile column 0; the ring row sits at jstrip + 1 = the last hot-kernel output
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int n = (3 * sms + S.ntile_i - 1) / S.ntile_i;      // a few small blocks per SM
        n = n > nz / 8 ? nz / 8 : n;
        n = n < 1 ? 1 : n;
        if (!two_ranges(A)) {
            S.chunk = (nz + n - 1) / n;
            if (g_chunk_override > 0) S.chunk = g_chunk_override;
            S.chunk = S.chunk < 2 ? 2 : (S.chunk > nz ? nz : S.chunk);
            S.clen = S.chunk;
        }
        constexpr size_t strip_smem = (2 * 8 * kStripRows * 32 + 4 * 8 * 32 * kStripCols) * sizeof(float) + 64;
        static unsigned long long sdone = 0;
        if (int e = ensure_smem(k_fused_strip<PATH>, strip_smem, sdone)) return e;
        CUtensorMap smap;
        // 48-byte row segments of 1216-byte rows: 64-byte L2 promotion fetches 167 MB per launch at 304x304x592 where
        // 128-byte promotion (and, measured, "none") fetch 277 MB -- for 46 MB needed.  The kernel's duration does not
        // depend on it (nor on cp.async vs TMA staging, nor on its register cap): see tools/experiments/README.md.
        if (!make_tile_map(&smap, A, nplanes_array, 32, kStripCols, CU_TENSOR_MAP_L2_PROMOTION_L2_64B)) { set_error("remainder strip: tensor map"); return IMHD_E_STATE; }
        k_fused_strip<PATH><<<dim3(S.ntile_i, 1, count_chunks(S)), dim3(32, strip_rows), strip_smem, st>>>(S, smap);
        IMHD_LAUNCH_CHECK(1);
    }
    timing_end(timed, st);
    return 0;
}

template <int PATH>
static int launch_fused(FusedArgs& A, int nplanes_array, cudaStream_t st, UnderStream* under) {
    const Params& P = A.P;
    const int nz = longest_range(A);
    if (nz <= 0) return 0;
    CUtensorMap tmap;
    // kernel choice (imhd_set_kernel_variant bits 4..7): 0 = default
    switch (g_kernel) {
        case 1:
            if (make_tile_map(&tmap, A, nplanes_array, TmaGeo<PATH, 16>::TR)) return launch_tma<PATH, OneRowLaunch<PATH, 16>>(A, tmap, nplanes_array, st, under);
            break;
        case 2:  // the 8-warp tile for either path
            if (make_tile_map(&tmap, A, nplanes_array, PairGeo<PATH, 8>::TR)) return launch_tma<PATH, PairLaunch<PATH, 8>>(A, tmap, nplanes_array, st, under);
            break;
        case 4:  // split-phase exchange barrier on the 8-warp tile
            if (make_tile_map(&tmap, A, nplanes_array, PairGeo<PATH, 8>::TR)) return launch_tma<PATH, SplitLaunch<PATH, 8>>(A, tmap, nplanes_array, st, under);
            break;
        default:
            // Path B: 8 warps x 2 rows = a 16x32 tile, one block per SM (its two-row ring makes smaller tiles too wasteful:
            // 4-warp tiles, 2 or 3 blocks per SM, measure 30.3 / 28.2 GLUPS against 32.0).  Path A: one ring row and ~170
            // registers suffice, so three INDEPENDENT 4-warp blocks per SM (8x32 tiles, 12 warps) win: 50.6 against 44.5.
            if (PATH == IMHD_PATH_A) {
                if (make_tile_map(&tmap, A, nplanes_array, PairGeo<PATH, 4>::TR)) return launch_tma<PATH, PairLaunch<PATH, 4, 3>>(A, tmap, nplanes_array, st, under);
            } else {
                if (make_tile_map(&tmap, A, nplanes_array, PairGeo<PATH, 8>::TR)) return launch_tma<PATH, SplitLaunch<PATH, 8>>(A, tmap, nplanes_array, st, under);
            }
            break;
    }
    constexpr int TI = 16;
    constexpr int O = Ring<PATH>::O;
    constexpr int WI = TI - 2 * O, WJ = 32 - 2 * O;
    // cells the corrector updates, counted from index 1: i in [1, Nx-1] (A) / [1, Nx-2] (B); edge cells ride along
    const int ni = PATH == IMHD_PATH_A ? P.Nx - 1 : P.Nx - 2, nj = PATH == IMHD_PATH_A ? P.Ny - 1 : P.Ny - 2;
    A.ntile_i = (ni + WI - 1) / WI;
    A.ntile_j = (nj + WJ - 1) / WJ;
    pick_chunk(A, nz);
    const int nchunk = count_chunks(A);
    const size_t smem = 2 * 2 * 8 * TI * 32 * sizeof(float);
    static unsigned long long done = 0;
    if (int e = ensure_smem(k_fused_step_ldg<PATH, TI>, smem, done)) return e;
    if (under)
        if (int e = stream_follows(under->s, st, under->fork)) return e;
    k_fused_step_ldg<PATH, TI><<<dim3(A.ntile_j, A.ntile_i, nchunk), dim3(32, TI), smem, st>>>(A);
    IMHD_LAUNCH_CHECK(1);
    return 0;
}

// Output planes [kfrom, kmid1) and [kmid2, kto) of the slab (global indices, clipped to the owned range; kmid1 == kmid2: one
// range).  Any split of the owned range into such calls writes the same bits as one imhd_step_fused call; the slab solver
// launches the planes next to BOTH slab ends first so their exchange overlaps the interior launch -- in ONE launch of the
// marching kernel when the two ranges hold the same number of its planes (at most kMaxEndPlanes), else in two.
constexpr int kMaxEndPlanes = 32;
static int step_ranges(const float* Qin, float* Qout, const float* qint_lo, const float* qint_hi, const float* qint_wrap,
                       const imhd_slab* s, int kfrom, int kmid1, int kmid2, int kto, void* stream) {
    FusedArgs A;
    if (int e = fill_args(A, Qin, Qout, qint_lo, qint_hi, qint_wrap, s)) return e;
    kfrom = max(kfrom, A.k0); kto = min(kto, A.k1);
    kmid1 = min(max(kmid1, kfrom), kto); kmid2 = min(max(kmid2, kmid1), kto);
    if (kfrom >= kto) return 0;
    if (kmid1 < kmid2 && kfrom < kmid1 && kmid2 < kto) {   // two non-empty ranges with a gap between them
        const int a0 = max(kfrom, A.ka0), a1 = min(kmid1, A.kb0), b0 = max(kmid2, A.ka0), b1 = min(kto, A.kb0);   // marching planes
        const bool one_launch = a1 - a0 == b1 - b0 && a1 - a0 >= 1 && a1 - a0 <= kMaxEndPlanes && b0 > a1;
        if (!one_launch) {
            if (int e = step_ranges(Qin, Qout, qint_lo, qint_hi, qint_wrap, s, kfrom, kmid1, kmid1, kmid1, stream)) return e;
            return step_ranges(Qin, Qout, qint_lo, qint_hi, qint_wrap, s, kmid2, kto, kto, kto, stream);
        }
        A.chunk = b0 - a0;
        A.clen = a1 - a0;
    }
    const bool do_front = kfrom == 0, do_back = kto == s->Nz;
    A.kfrom = max(kfrom, A.ka0); A.kto = min(kto, A.kb0);
    if (Qin == Qout) { set_error("imhd_step_fused: Qin and Qout must be distinct buffers"); return IMHD_E_INVALID; }
    if (!qint_lo || !qint_hi || (s->path == IMHD_PATH_B && s->k0 == 0 && !qint_wrap)) {
        set_error("imhd_step_fused: missing predictor plane (qint_lo=%p qint_hi=%p qint_wrap=%p)", (const void*)qint_lo,
                  (const void*)qint_hi, (const void*)qint_wrap);
        return IMHD_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const Params& P = A.P;
    const unsigned pb = (unsigned)((P.plane + 255) / 256);
    if (s->path == IMHD_PATH_A) {
        if (int e = launch_fused<IMHD_PATH_A>(A, s->nzl + 2 * (s->ghosts ? 1 : 0), st, nullptr)) return e;
        if (A.k0 == 0 && A.k1 == P.Nz && do_back) {  // PBCs on one GPU; across slabs the ghost exchange carries this plane
            k_plane_copy<<<pb, 256, 0, st>>>(Qout, (long long)(0 - A.kbase) * P.plane, (long long)(P.Nz - 1 - A.kbase) * P.plane,
                                             P.plane, A.vs);
            IMHD_LAUNCH_CHECK(1);
        }
        return 0;
    }
    // a launch long enough for the remainder strip (the same bound) carries its small launches under the marching kernel
    UnderStream* under = !g_no_under && !two_ranges(A) && A.kto - A.kfrom >= 64 ? under_stream() : nullptr;   // (forked by the launcher)
    if (int e = launch_fused<IMHD_PATH_B>(A, s->nzl + 2 * (s->ghosts ? 1 : 0), st, under)) return e;
    cudaStream_t small = under ? under->s : st;
    if (do_front) {
        if (under) k_front_plane_B<<<dim3((P.Ny + 31) / 32, (P.Nx + 1) / 2), dim3(32, 2), 0, small>>>(A);
        else       k_front_plane_B<<<dim3((P.Ny + 31) / 32, (P.Nx + 7) / 8), dim3(32, 8), 0, small>>>(A);
        IMHD_LAUNCH_CHECK(1);
    }
    if (do_back) {
        k_back_plane_B<<<pb, 256, 0, small>>>(A);
        IMHD_LAUNCH_CHECK(1);
    }
    if (under)
        if (int e = stream_follows(st, under->s, under->join)) return e;
    return 0;
}

extern "C" int imhd_step_fused_planes(const float* Qin, float* Qout, const float* qint_lo, const float* qint_hi,
                                      const float* qint_wrap, const imhd_slab* s, int kfrom, int kto, void* stream) {
    return step_ranges(Qin, Qout, qint_lo, qint_hi, qint_wrap, s, kfrom, kto, kto, kto, stream);
}

extern "C" int imhd_step_fused_ends(const float* Qin, float* Qout, const float* qint_lo, const float* qint_hi,
                                    const float* qint_wrap, const imhd_slab* s, int kfrom, int kmid1, int kmid2, int kto, void* stream) {
    return step_ranges(Qin, Qout, qint_lo, qint_hi, qint_wrap, s, kfrom, kmid1, kmid2, kto, stream);
}

extern "C" int imhd_step_fused(const float* Qin, float* Qout, const float* qint_lo, const float* qint_hi,
                               const float* qint_wrap, const imhd_slab* s, void* stream) {
    if (!s) { set_error("null slab descriptor"); return IMHD_E_INVALID; }
    return imhd_step_fused_planes(Qin, Qout, qint_lo, qint_hi, qint_wrap, s, s->k0, s->k0 + s->nzl, stream);
}
