// Drop-in drivers: `imhd-cuda` (path B, 37 positional arguments, src/on-device/main.cu:20-65) and
// `imhd-cuda_nodiff` (path A, 49 positional arguments, src/on-device/no_diffusion.cu:20-82), built from this one
// file (-DIMHD_NODIFF selects path A).  Same argv lists, same semantics: Nt-1 steps (it = 1..Nt-1), frame 0 = the
// initial condition written with attributes, grid.h5, one fluidvars_<it>.h5 per output step into path_to_data.
// What changed is underneath: one fused kernel per step through libimhd_b200.so, no per-phase device sync, and the
// per-step blocking D2H + fork of mpirun/PHDF5 (main.cu:216-226) is an asynchronous snapshot -> D2H -> writer thread.
// The execution-configuration arguments (block dims, SM multipliers) are accepted and ignored: the library picks its
// own tiling.  The writer/attribute/grid binary names and num_proc are accepted and unused (output is in-process);
// eigen_bin_name other than "none" runs the CFL scan on the device (imhd_ctx_stability) and prints the scanner's report.
// Optional environment: IMHD_OUTPUT_EVERY=n (default 1 = the reference's behaviour), IMHD_DEVICE=d, IMHD_GPUS=N (the
// domain as N z-slabs on devices 0..N-1 of this box, exchanged over NVLink: imhd_create_multi; output is then gathered
// and written synchronously),
// IMHD_STABILITY_QUIRKS=1 makes the CFL report use the reference's own (slipped) y / z Jacobians,
// IMHD_IC=<registry key>[:p0[,p1]] selects the initial condition by the reference's registry key
// (include/on-device/utils/configurers.hpp:21-29) instead of the one each shipped driver hard-codes; parameters left
// out come from argv (J0, r_max_coeff, A, k_harmonic) where the argv list has them.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/imhd_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc__ = (call);                                                            \
        if (rc__) { fprintf(stderr, "%s\n", imhd_last_error()); return EXIT_FAILURE; } \
    } while (0)

int main(int argc, char* argv[]) {
#ifdef IMHD_NODIFF
    const int need = 49, path = IMHD_PATH_A;
#else
    const int need = 37, path = IMHD_PATH_B;
#endif
    if (argc - 1 < need) {
        fprintf(stderr, "usage: %s <%d positional arguments, see %s>\n", argv[0], need,
                path == IMHD_PATH_A ? "src/on-device/no_diffusion.cu:20-82" : "src/on-device/main.cu:20-65");
        return EXIT_FAILURE;
    }
    int a = 1;
    const int Nt = atoi(argv[a++]), Nx = atoi(argv[a++]), Ny = atoi(argv[a++]), Nz = atoi(argv[a++]);
    const float J0 = atof(argv[a++]), D = atof(argv[a++]);
#ifdef IMHD_NODIFF
    const float r_max_coeff = atof(argv[a++]);  // parsed and unused by the shipped initial condition (no_diffusion.cu:27, B-24)
#else
    const float r_max_coeff = 0.25f;
#endif
    const float x_min = atof(argv[a++]), x_max = atof(argv[a++]), y_min = atof(argv[a++]), y_max = atof(argv[a++]);
    const float z_min = atof(argv[a++]), z_max = atof(argv[a++]), dt = atof(argv[a++]);
    const std::string path_to_data = argv[a++];
    a += 3;  // phdf5_bin_name, attr_bin_name, write_grid_bin_name
    const std::string eigen_bin_name = argv[a++];
    a++;     // num_proc
#ifdef IMHD_NODIFF
    const float A = atof(argv[48]);
    const int n_harmonic = atoi(argv[49]);
#endif
    const char* env = getenv("IMHD_OUTPUT_EVERY");
    const int every = env && atoi(env) > 0 ? atoi(env) : 1;
    env = getenv("IMHD_DEVICE");
    const int device = env ? atoi(env) : 0;

    env = getenv("IMHD_GPUS");
    const int n_gpus = env && atoi(env) > 1 ? atoi(env) : 1;
    imhd_ctx* ctx = n_gpus > 1 ? imhd_create_multi(Nx, Ny, Nz, n_gpus, nullptr) : imhd_create(Nx, Ny, Nz, device);
    if (!ctx) { fprintf(stderr, "%s\n", imhd_last_error()); return EXIT_FAILURE; }
    CHECK(imhd_ctx_init_grids(ctx, x_min, x_max, y_min, y_max, z_min, z_max));
#ifdef IMHD_NODIFF
    const float k = 2 * M_PI * n_harmonic / (z_max - z_min);  // no_diffusion.cu:166
    std::string ic = "cubic-bennett-vortex-m0";                // no_diffusion.cu:168
#else
    const float k = 0.f, A = 0.f;
    std::string ic = "screwpinch-stride";                      // main.cu:105
#endif
    float ic_params[2] = {0.f, 0.f};
    int ic_given = 0;
    if (const char* sel = getenv("IMHD_IC")) {
        ic = sel;
        const size_t colon = ic.find(':');
        if (colon != std::string::npos) {
            ic_given = sscanf(ic.c_str() + colon + 1, "%f,%f", &ic_params[0], &ic_params[1]);
            ic.resize(colon);
        }
    }
    const int ic_n = imhd_registry_initializer_nparams(ic.c_str());
    if (ic_given < ic_n) {  // defaults from argv
        const float d0 = ic == "zpinch" ? r_max_coeff : ic == "cubic-bennett-vortex-m0" ? k : J0;
        const float d1 = ic == "cubic-bennett-vortex-m0" ? A : r_max_coeff;
        if (ic_given < 1) ic_params[0] = d0;
        if (ic_given < 2) ic_params[1] = d1;
    }
    printf("Initial condition: %s\n", ic.c_str());
    CHECK(imhd_ctx_initialize(ctx, ic.c_str(), ic_params, ic_n < 0 ? 0 : ic_n));
    CHECK(imhd_ctx_prime(ctx, path, D, dt));

    printf("Writing initial conditions and grid to %s\n", path_to_data.c_str());
    CHECK(imhd_ctx_write_frame(ctx, path_to_data.c_str(), 0));
    CHECK(imhd_ctx_write_grid(ctx, path_to_data.c_str()));
    if (eigen_bin_name != "none") {
        // the reference forks its host scanner here (no_diffusion.cu:259-276, compute_stability.cpp); same report, on the device
        imhd_stability st;
        if (getenv("IMHD_STABILITY_QUIRKS")) imhd_stability_mode(IMHD_STABILITY_REFERENCE_QUIRKS);  // the reference's own y / z matrices (B-26)
        CHECK(imhd_ctx_stability(ctx, dt, &st));
        printf("Old timestep: %g\nLargest violation: %g at (i,j,k) = (%d,%d,%d)\nNew timestep: %g\n"
               "Total number of stability violations detected: %llu\n",
               dt, st.max_lhs, st.i, st.j, st.k, st.dt_new, st.violations);
    }

    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 1; it < Nt; it++) {
        CHECK(imhd_ctx_step(ctx, 1));
        if (it % every == 0 || it == Nt - 1) {
            printf("Timestep %d complete, queueing fluidvars_%d.h5\n", it, it);
            CHECK(imhd_ctx_write_frame(ctx, path_to_data.c_str(), it));
        }
    }
    CHECK(imhd_ctx_synchronize(ctx));
    const double compute_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CHECK(imhd_ctx_flush_output(ctx));
    const double total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%d steps of %dx%dx%d: time loop %.3f s (%.2f Mcell-updates/s), with output drained %.3f s\n", Nt - 1, Nx, Ny, Nz,
           compute_s, 1e-6 * (double)Nx * Ny * Nz * (Nt - 1) / compute_s, total_s);
    imhd_destroy(ctx);
    return 0;
}
