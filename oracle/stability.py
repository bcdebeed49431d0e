"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's CFL / stability scanner
(/root/reference/src/on-device/utils/compute_stability.cpp), SURVEY.md row 8(f3).

The reference scans every cell on the host: builds the three 8x8 flux Jacobians A, B, C of the conserved
state (:183-462), takes the largest |eigenvalue| of each with Eigen (:165-181), forms

    LHS = (dt/dx)|lambda_A| + (dt/dy)|lambda_B| + (dt/dz)|lambda_C|            (:157-163)

counts cells with LHS >= 1, remembers the largest, and proposes dt_new = 0.1 * dt / max LHS (:139-141).
**Parity unpinned**: Eigen is not in this image and the reference ships no fixture for the scanner (it is disabled
in the shipped input, eigen_bin_name=none), so this restatement uses numpy's LAPACK eigenvalues in Eigen's place
and is anchored on theory instead: for the x-direction the reference's matrix is the textbook 8-wave Jacobian and
its spectrum is {0, u, u +- c_a, u +- c_s, u +- c_f} (tests/test_stability.py checks that to fp32 eigen-solver
accuracy).  The y and z matrices carry transcription slips (quirk B-26 below), so their spectra are NOT the MHD wave
speeds; `jacobians()` reproduces them verbatim for the record, `wave_speed_lhs()` is the exact bound the CUDA
scan implements.

B-26 (found while restating; located by differentiating the flux numerically, tests/test_stability.py):
  * B rows 5 and 7 (:340-356) and C rows 5 and 6 (:427-443): the induction rows have every entry with the sign of
    the x matrix's convention, i.e. the whole row is the negative of d(flux)/dU;
  * B(7,5) (:370) has `+ By * Bdotu` where the pattern of A(7,4) / C(7,6) is `- Bdotu`;
  * C row 3 (:409-416) repeats row 2's magnetic part: C(2,4) = -Bz, C(2,6) = -Bx instead of C(2,5) = -Bz,
    C(2,6) = -By;
  * C(3,6) (:425) is `-gamma * By` where the pattern is `-gamma * Bz`.
"""
from __future__ import annotations

import numpy as np

GAMMA = 5.0 / 3.0


def _prims(U):
    """U: (..., 8) conserved states -> primitives the way computeA/B/C derive them (:185-206), in fp32."""
    U = np.asarray(U, np.float32)
    rho = U[..., 0]
    with np.errstate(divide="ignore", invalid="ignore"):  # rho == 0 cells become inf / NaN, as in the reference
        u, v, w = U[..., 1] / rho, U[..., 2] / rho, U[..., 3] / rho
    Bx, By, Bz, e = U[..., 4], U[..., 5], U[..., 6], U[..., 7]
    usq = u * u + v * v + w * w
    Bsq = Bx * Bx + By * By + Bz * Bz
    Bdotu = Bx * u + By * v + Bz * w
    return rho, u, v, w, Bx, By, Bz, e, usq, Bsq, Bdotu


def jacobians(U):
    """The reference's A, B, C (verbatim, slips included) for states U (..., 8) -> three (..., 8, 8) fp32 arrays."""
    rho, u, v, w, Bx, By, Bz, e, usq, Bsq, Bdotu = _prims(U)
    g = np.float32(GAMMA)
    shape = rho.shape + (8, 8)
    A, B, C = (np.zeros(shape, np.float32) for _ in range(3))
    H = (g * e + (2 - g) * 0.5 * Bsq) / rho  # the bracket shared by the three energy rows

    def fill(M, rows):
        for (r, c), val in rows.items():
            M[..., r, c] = val

    # ---- A (:208-274) -------------------------------------------------------------------------------------------
    fill(A, {(0, 1): 1.0,
             (1, 0): 0.5 * (g - 1) * usq - u * u, (1, 1): u * (3 - g), (1, 2): v * (1 - g), (1, 3): w * (1 - g),
             (1, 4): -g * Bx, (1, 5): (2 - g) * By, (1, 6): (2 - g) * Bz, (1, 7): g - 1,
             (2, 0): -u * v, (2, 1): v, (2, 2): u, (2, 4): -By, (2, 5): -Bx,
             (3, 0): -u * w, (3, 1): w, (3, 3): u, (3, 4): -Bz, (3, 6): -Bx,
             (5, 0): (v * Bx - u * By) / rho, (5, 1): By / rho, (5, 2): -Bx / rho, (5, 4): -v, (5, 5): u,
             (6, 0): (w * Bx - u * Bz) / rho, (6, 1): Bz / rho, (6, 3): -Bx / rho, (6, 4): -w, (6, 6): u,
             (7, 0): u * ((g - 1) * usq - H) + Bx * Bdotu / rho,
             (7, 1): H + (1 - g) * (u * u + 0.5 * usq) - Bx * Bx / rho,
             (7, 2): (1 - g) * u * v - By * Bx / rho, (7, 3): (1 - g) * u * w - Bx * Bz / rho,
             (7, 4): (1 - g) * u * Bx - Bdotu, (7, 5): (2 - g) * u * By - v * Bx, (7, 6): (2 - g) * u * Bz - w * Bx,
             (7, 7): u * g})
    # ---- B (:296-373) -------------------------------------------------------------------------------------------
    fill(B, {(0, 2): 1.0,
             (1, 0): -u * v, (1, 1): v, (1, 2): u, (1, 4): -By, (1, 5): -Bx,
             (2, 0): 0.5 * (g - 1) * usq - v * v, (2, 1): (1 - g) * u, (2, 2): (3 - g) * v, (2, 3): (1 - g) * w,
             (2, 4): (2 - g) * Bx, (2, 5): -g * By, (2, 6): (2 - g) * Bz, (2, 7): g - 1,
             (3, 0): -v * w, (3, 2): w, (3, 3): v, (3, 5): -Bz, (3, 6): -By,
             (4, 0): (v * Bx - u * By) / rho, (4, 1): By / rho, (4, 2): -Bx / rho, (4, 4): -v, (4, 5): u,   # slip
             (6, 0): (v * Bz - w * By) / rho, (6, 2): -Bz / rho, (6, 3): By / rho, (6, 5): w, (6, 6): -v,
             (7, 0): v * ((g - 1) * usq - H) + By * Bdotu / rho,
             (7, 1): (1 - g) * u * v - Bx * By / rho,
             (7, 2): H + (1 - g) * (v * v + 0.5 * usq) - By * By / rho,
             (7, 3): (1 - g) * v * w - By * Bz / rho,
             (7, 4): (2 - g) * v * Bx - u * By, (7, 5): (1 - g) * v * By + By * Bdotu,                        # slip
             (7, 6): (2 - g) * v * Bz - w * By, (7, 7): v * g})
    # ---- C (:395-459) -------------------------------------------------------------------------------------------
    fill(C, {(0, 3): 1.0,
             (1, 0): -u * w, (1, 1): w, (1, 3): u, (1, 4): -Bz, (1, 6): -Bx,
             (2, 0): -v * w, (2, 2): w, (2, 3): v, (2, 4): -Bz, (2, 6): -Bx,                                  # slip
             (3, 0): 0.5 * (g - 1) * usq - w * w, (3, 1): (1 - g) * u, (3, 2): (1 - g) * v, (3, 3): (3 - g) * w,
             (3, 4): (2 - g) * Bx, (3, 5): (2 - g) * By, (3, 6): -g * By, (3, 7): g - 1,                      # slip
             (4, 0): (w * Bx - u * Bz) / rho, (4, 1): Bz / rho, (4, 3): -Bx / rho, (4, 4): -w, (4, 6): u,
             (5, 0): (w * By - v * Bz) / rho, (5, 2): Bz / rho, (5, 3): -By / rho, (5, 5): -w, (5, 6): v,
             (7, 0): w * ((g - 1) * usq - H) + Bz * Bdotu / rho,
             (7, 1): (1 - g) * u * w - Bx * Bz / rho, (7, 2): (1 - g) * v * w - By * Bz / rho,
             (7, 3): H + (1 - g) * (w * w + 0.5 * usq) - Bz * Bz / rho,
             (7, 4): (2 - g) * w * Bx - u * Bz, (7, 5): (2 - g) * w * By - v * Bz, (7, 6): (1 - g) * w * Bz - Bdotu,
             (7, 7): w * g})
    return A, B, C


def spectral_radii(U):
    """max |eigenvalue| of the reference's A, B, C per state (getLargestEVs, :165-181) -> (..., 3)."""
    return np.stack([np.abs(np.linalg.eigvals(M)).max(axis=-1) for M in jacobians(U)], axis=-1).astype(np.float32)


def wave_speeds(U):
    """Exact spectral radius of the 8-wave ideal-MHD flux Jacobian per direction -> (..., 3), in fp64.

    Spectrum in direction d: {0, u_d, u_d +- c_a, u_d +- c_s, u_d +- c_f} with c_a^2 = B_d^2/rho and c_f^2, c_s^2 the
    roots of x^2 - (a^2 + b^2) x + a^2 B_d^2/rho = 0, a^2 = gamma p / rho, b^2 = B^2/rho.  A negative root (p < 0,
    a non-physical state) is an imaginary speed: the pair u_d +- i sqrt(-x) has modulus sqrt(u_d^2 - x), which is
    what Eigen's complex abs() returns in the reference (:173).  p uses the proper kinetic energy rho u^2 / 2 as
    the scanner does (:206), unlike the solver's own helper."""
    rho, u, v, w, Bx, By, Bz, e, usq, Bsq, _ = (np.asarray(q, np.float64) for q in _prims(U))
    p = (GAMMA - 1.0) * (e - 0.5 * rho * usq - 0.5 * Bsq)
    out = []
    with np.errstate(divide="ignore", invalid="ignore"):  # rho == 0 -> NaN cells, skipped by scan()
        a2, b2 = GAMMA * p / rho, Bsq / rho
        for ud, Bd in ((u, Bx), (v, By), (w, Bz)):
            ca2 = Bd * Bd / rho
            disc = np.sqrt(np.maximum((a2 + b2) ** 2 - 4.0 * a2 * ca2, 0.0))
            best = np.abs(ud)
            for x in (0.5 * (a2 + b2 + disc), 0.5 * (a2 + b2 - disc), ca2):
                mod = np.where(x >= 0, np.abs(ud) + np.sqrt(np.abs(x)), np.sqrt(ud * ud + np.abs(x)))
                best = np.maximum(best, mod)
            out.append(best)
    return np.stack(out, axis=-1)


def wave_speed_lhs(Q, dt, dx, dy, dz):
    """LHS field (Nz, Nx, Ny) of the stability criterion with the exact wave speeds, fp64."""
    U = np.moveaxis(np.asarray(Q, np.float32), 0, -1)
    lam = wave_speeds(U)
    return dt / dx * lam[..., 0] + dt / dy * lam[..., 1] + dt / dz * lam[..., 2]


def reference_lhs(Q, dt, dx, dy, dz):
    """LHS field (Nz, Nx, Ny) exactly as the reference forms it (:157-181): max |eigenvalue| of ITS A, B, C (slips
    included) per cell -- what the library's IMHD_STABILITY_REFERENCE_QUIRKS mode reports."""
    U = np.moveaxis(np.asarray(Q, np.float32), 0, -1)
    lam = spectral_radii(U).astype(np.float64)
    return dt / dx * lam[..., 0] + dt / dy * lam[..., 1] + dt / dz * lam[..., 2]


def scan(lhs, dt, alpha=0.1):
    """The scanner's summary (:112-141): number of cells with LHS >= 1, the largest LHS and its (i, j, k), the
    proposed dt = alpha * dt / max LHS.  `lhs` has shape (Nz, Nx, Ny); NaN cells (rho = 0) are ignored, as a NaN
    never satisfies `>= 1.0` in the reference either."""
    clean = np.where(np.isnan(lhs), -np.inf, lhs)
    k, i, j = np.unravel_index(int(np.argmax(clean)), clean.shape)
    mx = float(clean[k, i, j])
    return {"violations": int((clean >= 1.0).sum()), "max_lhs": mx, "argmax_ijk": (int(i), int(j), int(k)),
            "dt_new": float(alpha * dt / mx) if mx > 0 else float("inf")}
