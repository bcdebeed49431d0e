// Device math of the ideal-MHD Lax-Wendroff update: thermo helpers, the two flux families,
// diffusion stencil.  Two arithmetic recipes, chosen by the template flag EXACT:
//
//   EXACT = true   mirrors the reference's fp32/fp64 rounding points operation by operation
//                  (SURVEY.md A.6).  Used by the parity-granular operators (imhd_granular.cu,
//                  compiled with -fmad=false) so they can be compared bit for bit.
//   EXACT = false  the fast recipe of the fused kernels: fp32 flux evaluation with one
//                  reciprocal per state, and the O(1) accumulations arranged so that the
//                  last rounding is the only significant one (DESIGN.md "Precision").
//
// Reference formulas (file:line relative to the reference root):
//   helpers          lib/on-device/helper_functions.cu:7-62
//   indexed fluxes   lib/on-device/kernels_od_fluxes.cu:112-275  (predictor, BoundaryConditions)
//   local fluxes     lib/on-device/kernels_od_fluxes.cu:8-104    (corrector)
//   diffusion        lib/on-device/diffusion.cu:8-19
// The live quirks of those formulas are reproduced on purpose.
#pragma once
#include <cuda_runtime.h>

namespace imhd {

enum { RHO = 0, MX = 1, MY = 2, MZ = 3, BX = 4, BY = 5, BZ = 6, EN = 7 };
enum { DIR_X = 0, DIR_Y = 1, DIR_Z = 2 };

constexpr double kGamma = 5.0 / 3.0;  // include/on-device/kernels_od.cuh:16 (a double)
constexpr double kGm1 = kGamma - 1.0;
constexpr float kGm1f = (float)(kGamma - 1.0);

__device__ __forceinline__ double sqd(float x) { return (double)x * (double)x; }  // pow(float,2)

// ------------------------------------------------------------------------------------------
// Derived quantities of one state (helper_functions.cu).
// ------------------------------------------------------------------------------------------
template <bool EXACT>
struct Aux {
    float Bsq, ke, p, Bdotu;
    double inv;   // EXACT: 1.0/rho as the indexed family uses it
    float invf;   // fast: fp32 reciprocal of rho
};

template <bool EXACT>
__device__ __forceinline__ float h_Bsq(float bx, float by, float bz) {
    if (EXACT) return (float)(sqd(bx) + sqd(by) + sqd(bz));
    return fmaf(bz, bz, fmaf(by, by, bx * bx));
}
template <bool EXACT>
__device__ __forceinline__ float h_KE(float rho, float mx, float my, float mz, float invf) {
    if (EXACT) return (float)((1.0 / rho) * (sqd(mx) + sqd(my) + sqd(mz)));  // no 1/2 (B-1)
    return invf * fmaf(mz, mz, fmaf(my, my, mx * mx));
}
template <bool EXACT>
__device__ __forceinline__ float h_p(float e, float Bsq, float ke) {
    if (EXACT) return (float)(kGm1 * ((e - ke) - Bsq / 2.0));
    return kGm1f * fmaf(-0.5f, Bsq, e - ke);
}
template <bool EXACT>
__device__ __forceinline__ float h_Bdotu(float rho, float mx, float my, float mz, float bx, float by,
                                         float bz, float invf) {
    if (EXACT) return (float)((1.0 / rho) * (mx * bx + my * by + mz * bz));
    return invf * fmaf(mz, bz, fmaf(my, by, mx * bx));
}

#ifndef IMHD_RCP_NEWTON
#define IMHD_RCP_NEWTON 0   // 1: one Newton step on the hardware approximation
#endif
__device__ __forceinline__ float fast_rcp(float x) {
    // The hardware approximation (rcp.approx.ftz.f32: relative error <= 2^-23, i.e. one ulp).  A Newton step on top of it
    // (IMHD_RCP_NEWTON=1) costs two fp32x2 instructions per state and changes nothing measurable: the reciprocal only
    // enters the flux differences, which carry dt/dx ~ 5e-3, and the 100-step normalised L-inf against the reference stays
    // at 2e-7 .. 1.2e-6 either way (measured, tests/test_gpu_parity.py).
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
#if IMHD_RCP_NEWTON
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return r;
#endif
}

template <bool EXACT>
__device__ __forceinline__ Aux<EXACT> make_aux(const float U[8]) {
    Aux<EXACT> a;
    a.inv = EXACT ? 1.0 / U[RHO] : 0.0;
    a.invf = EXACT ? 0.0f : fast_rcp(U[RHO]);
    a.Bsq = h_Bsq<EXACT>(U[BX], U[BY], U[BZ]);
    a.ke = h_KE<EXACT>(U[RHO], U[MX], U[MY], U[MZ], a.invf);
    a.p = h_p<EXACT>(U[EN], a.Bsq, a.ke);
    a.Bdotu = h_Bdotu<EXACT>(U[RHO], U[MX], U[MY], U[MZ], U[BX], U[BY], U[BZ], a.invf);
    return a;
}

// ------------------------------------------------------------------------------------------
// INDEXED flux family (kernels_od_fluxes.cu:112-275): d-direction flux of all 8 variables of
// state U.  c = component index of the direction (MX+d, BX+d).
// ------------------------------------------------------------------------------------------
template <bool EXACT, int DIR>
__device__ __forceinline__ void flux_indexed(const float U[8], const Aux<EXACT>& a, float f[8]) {
    const float rho = U[RHO];
    const float md = U[MX + DIR], bd = U[BX + DIR];
    f[RHO] = md;  // :112-126
    if (EXACT) {
        const double inv = a.inv;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float mc = U[MX + c], bc = U[BX + c];
            if (c == DIR) {
                // :129-137,159-167,184-192   (1/rho)*m^2 - B^2 + p + Bsq/2
                f[MX + c] = (float)(inv * sqd(md) - sqd(bd) + a.p + a.Bsq / 2.0);
                f[BX + c] = 0.0f;  // :195,217,237
            } else {
                // :138-149,168-173 and their aliases :152-183: (1/rho)*m_lo*m_hi - B_lo*B_hi with
                // lo < hi component order (x before y before z), the product order of the source
                const int lo = c < DIR ? c : DIR, hi = c < DIR ? DIR : c;
                f[MX + c] = (float)(inv * U[MX + lo] * U[MX + hi] - U[BX + lo] * U[BX + hi]);
                // induction (B-2): the primary definitions are YFluxBX, ZFluxBX, ZFluxBY
                //   G(Bx) = (1/rho) mx By - Bx my ; H(Bx) = (1/rho) mx Bz - Bx mz ;
                //   H(By) = (1/rho) my Bz - By mz ; the transposed ones are -1.0 * those.
                if (c < DIR) f[BX + c] = (float)(inv * mc * bd - bc * md);
                else         f[BX + c] = (float)(-1.0 * (float)(inv * md * bc - bd * mc));
            }
        }
        // :243-275 all fp32 (B-3): e + p + Bsq*(m_d/rho) - Bdotu*B_d
        f[EN] = U[EN] + a.p + a.Bsq * (md / rho) - a.Bdotu * bd;
    } else {
        const float inv = a.invf;
        const float ud = md * inv;
        const float ptot = fmaf(0.5f, a.Bsq, a.p);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float mc = U[MX + c], bc = U[BX + c];
            if (c == DIR) {
                f[MX + c] = fmaf(ud, md, fmaf(-bd, bd, ptot));
                f[BX + c] = 0.0f;
            } else {
                f[MX + c] = fmaf(ud, mc, -(bc * bd));
                if (c < DIR) f[BX + c] = fmaf(inv * mc, bd, -(bc * md));
                else         f[BX + c] = -fmaf(ud, bc, -(bd * mc));
            }
        }
        f[EN] = fmaf(-a.Bdotu, bd, fmaf(a.Bsq, ud, U[EN] + a.p));
    }
}

// ------------------------------------------------------------------------------------------
// LOCAL flux family (kernels_od_fluxes.cu:8-104): d-direction flux of state U with
// EXPLICIT p, Bsq, Bdotu (the corrector feeds mixed-neighbour values at k-1: B-4, B-5).
// ------------------------------------------------------------------------------------------
template <bool EXACT, int DIR>
__device__ __forceinline__ void flux_local(const float U[8], float p, float Bsq, float Bdotu, float invf,
                                           float f[8]) {
    const float rho = U[RHO];
    const float md = U[MX + DIR], bd = U[BX + DIR];
    f[RHO] = md;
    if (EXACT) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float mc = U[MX + c], bc = U[BX + c];
            if (c == DIR) {
                f[MX + c] = (float)(sqd(md) / rho - sqd(bd) + p + 0.5 * Bsq);  // :12-14,49-51,82-84
                f[BX + c] = 0.0f;
            } else {
                f[MX + c] = (mc * md) / rho - bc * bd;              // fp32 (:16-22,45-47,53-55,78-80...)
                f[BX + c] = (mc / rho) * bd - (md / rho) * bc;      // fp32 (:28-34,57-67,90-96)
            }
        }
        f[EN] = (float)((U[EN] + p + 0.5 * Bsq) * (md / rho) - Bdotu * bd);  // :36-38,69-71,102-104
    } else {
        const float ud = md * invf;
        const float ptot = fmaf(0.5f, Bsq, p);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float mc = U[MX + c], bc = U[BX + c];
            if (c == DIR) {
                f[MX + c] = fmaf(ud, md, fmaf(-bd, bd, ptot));
                f[BX + c] = 0.0f;
            } else {
                f[MX + c] = fmaf(ud, mc, -(bc * bd));
                f[BX + c] = fmaf(mc * invf, bd, -(ud * bc));
            }
        }
        f[EN] = fmaf(U[EN] + ptot, ud, -(Bdotu * bd));
    }
}

// ------------------------------------------------------------------------------------------
// FAST recipe, flux tensors from once-per-state primitives.  The two families differ only in the
// induction and energy fluxes; momentum fluxes and the aliases F(my)=G(mx), F(mz)=H(mx),
// G(mz)=H(my) (kernels_od_fluxes.cu:152-183) are shared.  ~19 ops per state + ~13 per direction.
//
// Generic over the value type V:
//   V = float    one cell
//   V = float2   the two rows a thread of the register-tiled kernel owns (.x = first row, .y = second).  Operations with
//                at most two distinct register operands (the third a broadcast scalar, an immediate or a repeated
//                operand) are ONE packed fp32x2 instruction (FFMA2 / FMUL2 / FADD2): measured on B200
//                (tools/probe/packed_const_probe.cu) such an instruction holds the fp32 pipe for two cycles but the
//                issue port for one, so the slot it frees goes to the loads, shuffles and moves of the march.  A
//                multiply-add of three distinct register pairs would run at half rate as FFMA2 (three pairs, four
//                banks) and stays two scalar FFMA (`vfma3`).
// Each half of a packed instruction rounds exactly like the scalar instruction and both instantiations are the same
// text, so they give the same bits.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vfma3(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vfmac(float k, float a, float c) { return fmaf(k, a, c); }   // constant multiplier
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float vmulc(float k, float a) { return k * a; }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vrcp(float a) { return fast_rcp(a); }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }  // folds into operand modifiers
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 vfma3(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 vfmac(float k, float2 a, float2 c) { return __ffma2_rn(make_float2(k, k), a, c); }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 vmulc(float k, float2 a) { return __fmul2_rn(make_float2(k, k), a); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, vneg(b)); }
__device__ __forceinline__ float2 vrcp(float2 a) {
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(a.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(a.y));
#if IMHD_RCP_NEWTON
    return vfma(r, vfma(vneg(a), r, make_float2(1.0f, 1.0f)), r);
#else
    return r;
#endif
}
template <class V> __device__ __forceinline__ V vset(float s);
template <> __device__ __forceinline__ float vset<float>(float s) { return s; }
template <> __device__ __forceinline__ float2 vset<float2>(float s) { return make_float2(s, s); }

// per-half predicates of a value
template <class V> struct FlagT;
template <> struct FlagT<float> { bool x; };
template <> struct FlagT<float2> { bool x, y; };
__device__ __forceinline__ float vsel(FlagT<float> f, float a, float b) { return f.x ? a : b; }
__device__ __forceinline__ float2 vsel(FlagT<float2> f, float2 a, float2 b) { return make_float2(f.x ? a.x : b.x, f.y ? a.y : b.y); }
__device__ __forceinline__ FlagT<float> fand(FlagT<float> a, FlagT<float> b) { return {a.x && b.x}; }
__device__ __forceinline__ FlagT<float2> fand(FlagT<float2> a, FlagT<float2> b) { return {a.x && b.x, a.y && b.y}; }

template <class V>
struct PrimT {
    V ux, uy, uz;   // m / rho
    V Bsq, p, ptot, Bdotu;
};
typedef PrimT<float> Prim;

template <class V>
__device__ __forceinline__ PrimT<V> make_prim(const V U[8]) {
    PrimT<V> s;
    const V inv = vrcp(U[RHO]);
    s.ux = vmul(U[MX], inv); s.uy = vmul(U[MY], inv); s.uz = vmul(U[MZ], inv);
    s.Bsq = vfma(U[BZ], U[BZ], vfma(U[BY], U[BY], vmul(U[BX], U[BX])));
    const V ke = vfma3(s.uz, U[MZ], vfma3(s.uy, U[MY], vmul(s.ux, U[MX])));   // (m.m)/rho, no 1/2 (B-1)
    s.p = vmulc(kGm1f, vfmac(-0.5f, s.Bsq, vsub(U[EN], ke)));
    s.ptot = vfmac(0.5f, s.Bsq, s.p);
    s.Bdotu = vfma3(s.uz, U[BZ], vfma3(s.uy, U[BY], vmul(s.ux, U[BX])));
    return s;
}

// momentum part of the d-direction flux, common to both families
template <int DIR, class V>
__device__ __forceinline__ void flux_mom(const V U[8], const PrimT<V>& s, V f[8]) {
    const V ud = DIR == DIR_X ? s.ux : (DIR == DIR_Y ? s.uy : s.uz);
    const V bd = U[BX + DIR];
    f[RHO] = U[MX + DIR];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c == DIR) {
            f[MX + c] = vfma3(ud, U[MX + c], vfma(vneg(bd), bd, s.ptot));
        } else {
            // (m_lo/rho) * m_hi - B_lo*B_hi with lo < hi: one expression for both aliases
            const int lo = c < DIR ? c : DIR, hi = c < DIR ? DIR : c;
            const V ulo = lo == 0 ? s.ux : s.uy;
            f[MX + c] = vfma3(ulo, U[MX + hi], vneg(vmul(U[BX + lo], U[BX + hi])));
        }
    }
    f[BX + DIR] = vset<V>(0.0f);
}

// INDEXED family (predictor): induction with the undivided term (B-2), energy without parentheses (B-3)
template <int DIR, class V>
__device__ __forceinline__ void flux_idx(const V U[8], const PrimT<V>& s, V f[8]) {
    flux_mom<DIR>(U, s, f);
    const V ud = DIR == DIR_X ? s.ux : (DIR == DIR_Y ? s.uy : s.uz);
    const V md = U[MX + DIR], bd = U[BX + DIR];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c == DIR) continue;
        const V uc = c == 0 ? s.ux : (c == 1 ? s.uy : s.uz);
        // c < DIR: (m_c/rho) B_d - B_c m_d ;  c > DIR: -( (m_d/rho) B_c - B_d m_c )
        if (c < DIR) f[BX + c] = vfma3(uc, bd, vneg(vmul(U[BX + c], md)));
        else         f[BX + c] = vfma3(bd, U[MX + c], vneg(vmul(ud, U[BX + c])));
    }
    f[EN] = vfma3(vneg(s.Bdotu), bd, vfma3(s.Bsq, ud, vadd(U[EN], s.p)));
}

// LOCAL family (corrector): (m_c/rho) B_d - (m_d/rho) B_c ; (e + p + Bsq/2)(m_d/rho) - Bdotu B_d
template <int DIR, class V>
__device__ __forceinline__ void flux_loc(const V U[8], const PrimT<V>& s, V f[8]) {
    flux_mom<DIR>(U, s, f);
    const V ud = DIR == DIR_X ? s.ux : (DIR == DIR_Y ? s.uy : s.uz);
    const V bd = U[BX + DIR];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c == DIR) continue;
        const V uc = c == 0 ? s.ux : (c == 1 ? s.uy : s.uz);
        f[BX + c] = vfma3(uc, bd, vneg(vmul(ud, U[BX + c])));
    }
    f[EN] = vfma3(vadd(U[EN], s.ptot), ud, vneg(vmul(s.Bdotu, bd)));
}

// ------------------------------------------------------------------------------------------
// Diffusion stencil D * lap(q) (diffusion.cu:8-19); cx = 1/dx^2 etc. precomputed in fp64 on the
// host exactly as `1.0 / pow(dx, 2)`.
// ------------------------------------------------------------------------------------------
struct DiffCoef {
    double cx, cy, cz;     // EXACT
    float cxf, cyf, czf;   // fast
    float c0f;             // fast: -2 (cx + cy + cz)
    float dtD;             // fast: dt * D
};

// fast: q + dt*D*lap folded into one chain: r + dtD * (cx (xp+xm) + cy (yp+ym) + cz (zp+zm) + c0 q), from the three
// neighbour sums (the caller forms them: inside a register-tiled thread the x sum mixes its own rows).
// The cancellation error of this form is ~1e-7 * (2/dx^2) * dt*D * |q| << ulp(q) for every grid of interest.
template <class V>
__device__ __forceinline__ V add_diffusion(V r, V q, V xsum, V ysum, V zsum, const DiffCoef& c) {
    const V lap = vfmac(c.cxf, xsum, vfmac(c.cyf, ysum, vfmac(c.czf, zsum, vmulc(c.c0f, q))));
    return vfmac(c.dtD, lap, r);
}

template <bool EXACT>
__device__ __forceinline__ float num_diff(float q, float qip1, float qjp1, float qkp1, float qim1, float qjm1,
                                          float qkm1, float D, const DiffCoef& c) {
    if (EXACT)
        return (float)(D * (c.cx * (qip1 - 2.0 * q + qim1) + c.cy * (qjp1 - 2.0 * q + qjm1) +
                            c.cz * (qkp1 - 2.0 * q + qkm1)));
    const float m2q = -2.0f * q;
    return D * fmaf(c.czf, (qkp1 + m2q) + qkm1, fmaf(c.cyf, (qjp1 + m2q) + qjm1, c.cxf * ((qip1 + m2q) + qim1)));
}

// wall energy e <- p(e,0,0)/(gamma-1): kernels_fluidbcs.cu:173,451,461 (B-12)
__device__ __forceinline__ float wall_e(float e) {
    float p = (float)(kGm1 * ((e - 0.0f) - 0.0f / 2.0));
    return (float)(p / kGm1);
}

}  // namespace imhd
