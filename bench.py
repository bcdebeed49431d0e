#!/usr/bin/env python
"""Benchmark of the hot path: cell-updates/s of the fused Lax-Wendroff time step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1] / [2]): 304x304x592 fp32 Bennett screw pinch (ScrewPinchStride, J0=1,
domain +-3.14159), numerical diffusion ON (reference pipeline of src/on-device/main.cu, D=0.01, dt=1e-4).
At N GPUs the domain is 304x304x(592*N), z-slab decomposed, one slab per rank (weak scaling).  A "step" is
one time step of the whole domain.  Prints ONE JSON line on rank 0.

--impl reference times the reference's own kernels on the host cores (oracle/_ref, all threads) on a bounded
sample of the same workload; see DESIGN.md "Measurement".
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY, NZ_PER_GPU = 304, 304, 592
STRONG = (1024, 1024, 2048)  # BASELINE.json configs[3]: total work fixed, split across 2/4/8 GPUs (--workload strong)
BOUNDS = (-3.14159, 3.14159) * 3
J0, D, DT = 1.0, 0.01, 1e-4
BYTES_PER_CELL_UPDATE = 64  # 8 fp32 read + 8 fp32 written


def spacing(nz_per_gpu):
    """Grid spacing of the workload.  Weak scaling EXTENDS the domain in z (z_max = z_min + (Nz-1) dz with dz the
    N = 1 spacing) instead of refining it: squeezing 592 N planes into the same box would push the explicit
    diffusion past its stability limit dt D (2/dx^2 + 2/dy^2 + 2/dz^2) < 1/2 (lib/on-device/diffusion.cu:8-19)."""
    import numpy as np

    f = lambda lo, hi, n: float(np.float32((np.float32(hi) - np.float32(lo)) / np.float32(n - 1)))  # noqa: E731
    return f(BOUNDS[0], BOUNDS[1], NX), f(BOUNDS[2], BOUNDS[3], NY), f(BOUNDS[4], BOUNDS[5], nz_per_gpu)


def diffusion_number(dx, dy, dz):
    return DT * D * (2.0 / dx**2 + 2.0 / dy**2 + 2.0 / dz**2)


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[0]) for r in rows]
        out["samples"] = len(rows)
        out["sm_mhz"] = statistics.median(sm)
        out["sm_max_mhz"] = float(rows[0][1])
        out["power_w_max"] = max(float(r[2]) for r in rows if r[2].strip().replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for c, n in enumerate(names, start=3) if any("Active" in r[c] and "Not" not in r[c] for r in rows)]
        return out


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
def cpu_reference_rate(nplanes, nsteps, threads=None):
    """The reference's own kernels on the host cores (oracle/_ref; falls back to the C port), on a 304x304xnplanes
    sub-slab of the workload.  Returns (cell-updates/s, kind, cores, sample description)."""
    import numpy as np
    from oracle import oracle as om

    ref = om.best_available()
    cores = os.cpu_count() or 1
    if ref.kind == "reference":
        ref.T = threads or cores
    else:
        cores = int(ref.lib.oracle_num_threads())
    g = ref.init_grids(BOUNDS, NX, NY, nplanes)
    dx, dy, dz = spacing(NZ_PER_GPU)  # spacing of the full workload; the sample is a thin slab of it
    Q = ref.screwpinch_stride(J0, *g)
    Qi = np.zeros_like(Q)
    ref.prime(Q, Qi, om.PATH_B, D, DT, dx, dy, dz)
    t0 = time.perf_counter()
    ref.steps(Q, Qi, om.PATH_B, nsteps, D, DT, dx, dy, dz)
    dt = time.perf_counter() - t0
    assert np.isfinite(Q).all()
    sample = f"{NX}x{NY}x{nplanes} sub-slab of the workload, {nsteps} steps, path B (D={D}), {ref.kind} kernels on {cores} host threads"
    return NX * NY * nplanes * nsteps / dt, ref.kind, cores, sample, dt


def run_reference_arm(args, rank):
    if rank != 0:
        return
    nplanes = 16
    # each bench "step" = one time step of the bounded sample
    rate_w, kind, cores, sample, _ = cpu_reference_rate(nplanes, max(args.warmup, 1))
    rate, kind, cores, sample, secs = cpu_reference_rate(nplanes, args.steps)
    glups = rate / 1e9
    line = {
        "impl": "reference", "metric": "cell_updates_per_sec", "value": glups, "unit": "GLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.workload if args.workload in WORKLOADS else "weak") | {"sample": sample},
        "cpu_baseline": {"value": glups, "unit": "GLUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# The no-diffusion pipeline is not stable on this grid at dt = 1e-4: the REFERENCE'S OWN kernels (recompiled for sm_100,
# tools/check_pathA_blowup.py) leave the finite numbers between steps 101 and 125 on the screw pinch and before step 50
# on its Bennett vortex, and so does this library (same arithmetic) -- the blow-up comes at a fixed physical time
# (t ~ 0.012), not at a CFL limit.  The 1000-step legs of that pipeline therefore run dt = 1e-5 (t_end = 0.01, finite).
DT_NODIFF = 1e-5
WORKLOADS = {
    # name: (path, D, dt, description) -- BASELINE.json configs[1]/[2] (weak), [3] (strong), the no-diffusion pipeline, [4] (c5)
    "weak": ("B", D, DT, "numerical diffusion on (D=%g), reference pipeline src/on-device/main.cu (path B)" % D),
    "strong": ("B", D, DT, "numerical diffusion on (D=%g), reference pipeline src/on-device/main.cu (path B)" % D),
    "pathA": ("A", 0.0, DT_NODIFF, "diffusion off, reference pipeline src/on-device/no_diffusion.cu (path A)"),
    "c5": ("A", 0.0, DT_NODIFF, "diffusion off (path A) + fluidvars_<it>.h5 every 50 steps through the asynchronous writer"),
}


def workload_config(n, workload="weak"):
    weak = workload != "strong"
    dx, dy, dz = spacing(NZ_PER_GPU if weak else NZ_PER_GPU * n)
    path, Dw, dtw, what = WORKLOADS[workload]
    return {"workload": f"Bennett screw pinch {NX}x{NY}x{NZ_PER_GPU * n} fp32, dt={dtw}, {what}",
            "name": workload, "grid": [NX, NY, NZ_PER_GPU * n], "decomposition": f"z-slabs x{n}" if n > 1 else "single GPU",
            "spacing": [dx, dy, dz], "z_extent": dz * (NZ_PER_GPU * n - 1),
            "diffusion_number": dtw * Dw * (2.0 / dx**2 + 2.0 / dy**2 + 2.0 / dz**2),  # explicit limit 1/2
            "l2": f"inputs ({8 * 4 * NX * NY * NZ_PER_GPU / 1e9:.2f} GB per array per GPU) larger than L2, no flush needed"}


def source_hash():
    """Blob hashes of the hot kernel's sources: an ncu traffic figure is only quoted for the build it was taken on."""
    import hashlib

    h = hashlib.sha1()
    for f in ("imhd_fused.cu", "imhd_math.cuh"):
        h.update(open(os.path.join(ROOT, "imhd-cuda_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def ncu_traffic_per_launch(cells):
    """dram__bytes_read+write of the fused kernel per launch from the committed ncu capture -- only if that capture was
    taken on THIS workload and on the current kernel sources (profiles/fused_traffic.json stores their hash)."""
    p = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        if t.get("cells") == cells and t.get("source_sha1") == source_hash():
            return t.get("dram_bytes_per_launch")
    return None


class DevArray:
    """A device pointer of the library as a torch tensor (through __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000, help="timed steps (default: the 1000 steps of BASELINE.json configs[1], ~1.7 s)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="weak", choices=sorted(WORKLOADS),
                    help="weak (default, the driver's contract): 304x304x592 per GPU with diffusion; strong: 1024x1024x2048 in total; "
                         "pathA: the no-diffusion pipeline on the weak grid; c5: pathA + .h5 output every 50 steps (1 GPU)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (the strong workload pins 34 GB per rank at N=2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import ctypes as C

    import numpy as np
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if args.workload == "c5" and world != 1:
        raise SystemExit("--workload c5 (output every 50 steps) is the single-GPU configuration")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    pkg = importlib.import_module("imhd-cuda_b200")
    ops = pkg.ops
    lib = pkg._lib.load()
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    global NX, NY, NZ_PER_GPU
    if args.workload == "strong":
        free_b, _ = torch.cuda.mem_get_info()
        need = 2 * 8 * 4 * STRONG[0] * STRONG[1] * (STRONG[2] // world + 2) + (1 << 30)
        if free_b < need:
            raise SystemExit(f"--workload strong: 1024x1024x2048 on {world} GPU(s) needs {need / 2**30:.0f} GiB per GPU, "
                             f"{free_b / 2**30:.0f} GiB free")
        NX, NY = STRONG[0], STRONG[1]
        NZ_PER_GPU = STRONG[2] // world
    nz_global = NZ_PER_GPU * world
    path_name, Dw, DTw, _ = WORKLOADS[args.workload]
    path = pkg.PATH_B if path_name == "B" else pkg.PATH_A
    cfg = workload_config(world, args.workload)
    dx, dy, dz = cfg["spacing"]
    # weak scaling EXTENDS the z domain (constant dz): z_max = z_min + (Nz - 1) dz
    bounds = (BOUNDS[0], BOUNDS[1], BOUNDS[2], BOUNDS[3], BOUNDS[4], BOUNDS[4] + dz * (nz_global - 1))

    # ---- the solver: the C ABI context; at N > 1 one z-slab per process, the slab loop and its NCCL exchanges in C++ ----
    def make_ctx():
        if world == 1:
            return ops.Context(NX, NY, nz_global, device=local_rank)
        box = [ops.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return ops.Context.slab(NX, NY, nz_global, rank, world, local_rank, box[0])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(*vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    if os.environ.get("IMHD_EDGE"):  # tuning hook: planes launched ahead of the interior at each slab end
        lib.imhd_set_edge_planes(int(os.environ["IMHD_EDGE"]))
    ctx = make_ctx()
    k0, nzl, _ = ctx.slab_extent(0)
    cells_local = NX * NY * nzl
    cells_global = NX * NY * nz_global

    def reset_state():  # synthetic input: the screw pinch, initialised on the device(s)
        ctx.init_grids(*bounds)
        ctx.init_screwpinch_stride(J0)
        ctx.set_spacing(dx, dy, dz)
        ctx.prime(path, Dw, DTw)

    outdir = None
    frames = 0

    def run_steps(n, it0=0):
        nonlocal frames
        if args.workload != "c5":
            ctx.step(n)
            return
        done = 0
        while done < n:  # BASELINE configs[4]: a frame every 50 steps, queued without stopping the loop
            m = min(50 - (it0 + done) % 50, n - done)
            ctx.step(m)
            done += m
            if (it0 + done) % 50 == 0:
                ctx.write_frame(outdir, (it0 + done))
                frames += 1

    if args.workload == "c5":
        outdir = tempfile.mkdtemp(prefix="imhd_c5_", dir=os.environ.get("IMHD_C5_DIR"))
    reset_state()
    stream = torch.cuda.ExternalStream(lib.imhd_ctx_stream(ctx.h), device=torch.device("cuda", local_rank))
    # the clock sampler starts before the warm-up (nvidia-smi needs ~0.2 s to come up) and stops after the timed steps;
    # both run the same kernels, so every sample is a sample under this load
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3)
    run_steps(args.warmup)
    ctx.synchronize()
    if args.workload == "c5":
        ctx.flush_output()
        frames = 0
    barrier()

    # ---- device-resident throughput (`value`) + per-launch timing of the fused kernel (`roofline`) -----------------
    launches0 = ops.launch_count()
    lib.imhd_fused_timing(1)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record(stream)
    run_steps(args.steps, it0=args.warmup)
    t1.record(stream)
    ctx.synchronize()
    if args.workload == "c5":
        ctx.flush_output()  # the frames are part of this workload: the number includes draining them to disk
    wall_s = time.perf_counter() - w0
    barrier()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    loop_ms = t0.elapsed_time(t1)   # the time loop on the compute stream (c5: incl. the device-side snapshots, not the drain)
    ms = loop_ms if args.workload != "c5" else wall_s * 1e3
    tot_ms, n_l, cells_l = C.c_double(), C.c_int(), C.c_longlong()
    lib.imhd_fused_timing_read(C.byref(tot_ms), C.byref(n_l), C.byref(cells_l))
    lib.imhd_fused_timing(0)
    fused_ms = tot_ms.value / max(n_l.value, 1)
    cells_launch = cells_l.value // max(n_l.value, 1)
    ms, fused_ms = allmax(ms, fused_ms)
    value = cells_global * args.steps / (ms * 1e-3) / 1e9
    dev = torch.as_tensor(DevArray(lib.imhd_ctx_device_state(ctx.h), (8, nzl + (2 if world > 1 else 0), NX, NY)), device="cuda")
    finite = all(bool(torch.isfinite(dev[v]).all()) for v in range(8))  # per variable: bounded temporaries
    del dev
    out_bytes = 0
    if outdir:
        out_bytes = sum(os.path.getsize(os.path.join(outdir, f)) for f in os.listdir(outdir))
        for f in os.listdir(outdir):
            os.unlink(os.path.join(outdir, f))
        os.rmdir(outdir)

    # ---- end to end through host buffers (`e2e`): pinned host state -> device, K steps, result back to the host -----
    slab_bytes = 8 * cells_local * 4
    e2e = None
    if not args.no_e2e and args.workload != "c5":
        host_in = torch.empty((8, nzl, NX, NY), dtype=torch.float32, pin_memory=True)
        host_out = torch.empty_like(host_in, pin_memory=True)
        reset_state()
        ctx.get_state_local(0, host_in.numpy())

        def job(nsteps):
            if world == 1:  # ONE C-ABI call: imhd_run_host
                ctx.run_host(host_in.numpy(), host_out.numpy(), path, Dw, DTw, dx, dy, dz, nsteps)
            else:           # the same job per slab: owned planes in, prime (ghost refresh), K steps, owned planes out
                ctx.set_state_local(host_in.numpy(), 0)
                ctx.set_spacing(dx, dy, dz)
                ctx.prime(path, Dw, DTw)
                ctx.step(nsteps)
                ctx.get_state_local(0, host_out.numpy())

        job(2)  # warm-up
        barrier()
        w0 = time.perf_counter()
        job(args.steps)
        barrier()
        (e2e_s,) = allmax(time.perf_counter() - w0)
        # the copies alone, for the PCIe bound of this job
        c0 = time.perf_counter()
        ctx.set_state_local(host_in.numpy(), 0)
        c1 = time.perf_counter()
        ctx.get_state_local(0, host_out.numpy())
        c2 = time.perf_counter()
        finite = finite and bool(torch.isfinite(host_out).all())
        e2e = {"value": cells_global * args.steps / e2e_s / 1e9, "unit": "GLUPS",
               "h2d_bytes_per_step": slab_bytes * world / args.steps, "d2h_bytes_per_step": slab_bytes * world / args.steps,
               "h2d_GBps_per_gpu": slab_bytes / (c1 - c0) / 1e9, "d2h_GBps_per_gpu": slab_bytes / (c2 - c1) / 1e9,
               "note": f"one C-ABI job{'' if world == 1 else ' per slab'}: pinned host state -> device, prime, {args.steps} fused steps, "
                       f"state -> host; copies inside the timed region.  The steps need the whole state on the device, so the job is "
                       f"H2D + K steps + D2H in series: bounded by the two PCIe copies, not by the kernel"}
    ctx.close()

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved = BYTES_PER_CELL_UPDATE * cells_launch / (fused_ms * 1e-3) / 1e9 if n_l.value else None
        kern = (f"k_fused_split<PATH_B,8> (+ k_fused_wstrip<PATH_B> for the columns beyond the last full tile, co-resident under it on a side "
                f"stream; one event pair around both)" if path_name == "B" else
                f"k_fused_pair<PATH_A,4,3> (+ k_fused_wstrip<PATH_A> for the columns beyond the last full tile; timed together)")
        line = {
            "metric": "cell_updates_per_sec", "value": value, "unit": "GLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "strong" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "finite": finite,
            "host_loop": "C++ (libimhd_b200.so: imhd_ctx_step" + ("" if world == 1 else ", z-slab engine, planes exchanged by the copy engines (CUDA IPC) on a side stream; IMHD_SLAB_EXCHANGE=nccl: ncclSend/Recv") + ")",
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_cell_update": BYTES_PER_CELL_UPDATE,
                         "cell_updates_per_launch": cells_launch, "avg_launch_ms": fused_ms, "timed_launches": n_l.value,
                         "traffic": ncu_traffic_per_launch(cells_launch)},
        }
        if args.workload == "c5":
            line["output"] = {"frames": frames, "bytes": out_bytes, "every": 50, "time_loop_ms_per_step": loop_ms / args.steps,
                              "time_loop_glups": cells_global * args.steps / (loop_ms * 1e-3) / 1e9,
                              "note": "value = cell-updates / wall time of the loop INCLUDING draining the frames to disk"}
        if not args.no_cpu_baseline:
            rate, kind, cores, sample, _ = cpu_reference_rate(32, 2)
            line["cpu_baseline"] = {"value": rate / 1e9, "unit": "GLUPS", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not finite:
        raise SystemExit("bench.py: the state is not finite after the timed steps -- the number above is not a measurement")


if __name__ == "__main__":
    main()
