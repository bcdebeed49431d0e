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
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
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


def ncu_traffic_per_launch(cells):
    """dram__bytes_read+write of the fused kernel per launch from the committed ncu capture, if it was taken on this workload."""
    p = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        if t.get("cells") == cells:
            return t.get("dram_bytes_per_launch")
    return None


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
        "config": workload_config(args.gpus) | {"sample": sample},
        "cpu_baseline": {"value": glups, "unit": "GLUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n, weak=True):
    dx, dy, dz = spacing(NZ_PER_GPU if weak else NZ_PER_GPU * n)
    return {"workload": f"Bennett screw pinch {NX}x{NY}x{NZ_PER_GPU * n} fp32, numerical diffusion on (D={D}), dt={DT}, "
                        f"reference pipeline src/on-device/main.cu (path B)",
            "grid": [NX, NY, NZ_PER_GPU * n], "decomposition": f"z-slabs x{n}" if n > 1 else "single GPU",
            "spacing": [dx, dy, dz], "z_extent": dz * (NZ_PER_GPU * n - 1),
            "diffusion_number": diffusion_number(dx, dy, dz),  # dt D (2/dx^2+2/dy^2+2/dz^2), explicit limit 1/2
            "l2": f"inputs ({8 * 4 * NX * NY * NZ_PER_GPU / 1e9:.2f} GB per array per GPU) larger than L2, no flush needed"}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): 304x304x592 per GPU; strong: 1024x1024x2048 in total (needs >= 2 GPUs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (the strong workload pins 34 GB per rank at N=2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    pkg = importlib.import_module("imhd-cuda_b200")
    ops = pkg.ops
    slabmod = importlib.import_module("imhd-cuda_b200.slab")
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = slabmod.TorchComm()

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
    dx, dy, dz = spacing(NZ_PER_GPU) if args.workload == "weak" else spacing(nz_global)
    solver = slabmod.SlabSolver(NX, NY, nz_global, pkg.PATH_B, D, DT, dx, dy, dz, comm=comm, corner_e=0.0)
    L = solver.layout
    cells_local = NX * NY * L.nzl
    cells_global = NX * NY * nz_global

    # synthetic input: the screw pinch is z-invariant, so each rank initialises its own ghosted slab on device
    gx, gy, gz = ops.init_grids(BOUNDS, NX, NY, L.nzl + 2)

    def reset_state():
        solver.cur = 0
        check = ops._lib.load().imhd_init_screwpinch_stride
        ops.check(check(ops._dev(solver.Q[0]), J0, ops._dev(gx), ops._dev(gy), ops._dev(gz), NX, NY, L.nzl + 2, ops._stream()))

    reset_state()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) + per-launch timing of the fused kernel (`roofline`) -----------------
    orig_step_fused = ops.step_fused_planes
    ev_pairs = []

    def timed_step_fused(Qin, Qout, lo, hi, wrap, slab, kfrom, kto):
        # time the launch that covers the bulk of the slab (at N>1 the two 8-plane end launches go first, untimed)
        if kto - kfrom < L.nzl // 2:
            return orig_step_fused(Qin, Qout, lo, hi, wrap, slab, kfrom, kto)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_step_fused(Qin, Qout, lo, hi, wrap, slab, kfrom, kto)
        e1.record()
        ev_pairs.append((e0, e1, kto - kfrom))

    solver.compute = type("Compute", (), {"qint_plane": staticmethod(ops.qint_plane), "make_slab": staticmethod(ops.make_slab),
                                          "step_fused_planes": staticmethod(timed_step_fused)})
    solver.step(args.warmup)
    ev_pairs.clear()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ops.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    solver.step(args.steps)
    t1.record()
    barrier()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms = t0.elapsed_time(t1)
    fused_ms = sum(a.elapsed_time(b) for a, b, _ in ev_pairs) / len(ev_pairs)
    fused_planes = ev_pairs[0][2]
    if dist is not None:
        t = torch.tensor([ms, fused_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, fused_ms = t.tolist()
    value = cells_global * args.steps / (ms * 1e-3) / 1e9
    finite = all(bool(torch.isfinite(solver.state[v]).all()) for v in range(8))  # per variable: bounded temporaries

    # ---- end to end through host buffers (`e2e`): pinned host state -> device, K steps, result back to the host -----
    slab_bytes = 8 * cells_local * 4
    e2e_s = None
    if args.no_e2e:
        pass
    elif world == 1:
        host_in = torch.empty((8, L.nzl, NX, NY), dtype=torch.float32, pin_memory=True)
        host_out = torch.empty_like(host_in, pin_memory=True)
        reset_state()
        host_in.copy_(solver.Q[0][:, 1:-1])
        torch.cuda.synchronize()
        del solver  # the context owns its own buffers; free the slab solver's first
        torch.cuda.empty_cache()
        with pkg.Context(NX, NY, nz_global, device=local_rank) as ctx:
            ctx.run_host(host_in.numpy(), host_out.numpy(), pkg.PATH_B, D, DT, dx, dy, dz, 2)  # warm-up
            w0 = time.perf_counter()
            ctx.run_host(host_in.numpy(), host_out.numpy(), pkg.PATH_B, D, DT, dx, dy, dz, args.steps)  # synchronises
            e2e_s = time.perf_counter() - w0
        finite = finite and bool(torch.isfinite(host_out).all())
    else:
        host = torch.empty((8, L.nzl + 2, NX, NY), dtype=torch.float32, pin_memory=True)
        reset_state()
        host.copy_(solver.Q[0])
        solver.compute = ops
        barrier()
        w0 = time.perf_counter()
        solver.cur = 0
        solver.Q[0].copy_(host, non_blocking=True)
        solver.step(args.steps)
        host.copy_(solver.Q[solver.cur], non_blocking=True)
        barrier()
        e2e_s = time.perf_counter() - w0
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e = None if e2e_s is None else {"value": cells_global * args.steps / e2e_s / 1e9, "unit": "GLUPS",
           "h2d_bytes_per_step": slab_bytes * world / args.steps, "d2h_bytes_per_step": slab_bytes * world / args.steps,
           "note": f"one C-ABI job: pinned host state -> device, {args.steps} fused steps, state -> host; copies inside the timed region"}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        cells_launch = NX * NY * fused_planes
        achieved = BYTES_PER_CELL_UPDATE * cells_launch / (fused_ms * 1e-3) / 1e9
        line = {
            "metric": "cell_updates_per_sec", "value": value, "unit": "GLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.workload,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world, args.workload == "weak"), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "finite": finite,
            "roofline": {"bound": "hbm", "kernel": "k_fused_step_tma<PATH_B,16> (+ k_fused_strip<PATH_B> for the columns beyond the last full tile; timed together)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "algorithmic_bytes_per_cell_update": BYTES_PER_CELL_UPDATE,
                         "cell_updates_per_launch": cells_launch, "avg_launch_ms": fused_ms,
                         "traffic": ncu_traffic_per_launch(cells_launch)},
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, kind, cores, sample, _ = cpu_reference_rate(32, 2)
            line["cpu_baseline"] = {"value": rate / 1e9, "unit": "GLUPS", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if not finite:
        raise SystemExit("bench.py: the state is not finite after the timed steps -- the number above is not a measurement")


if __name__ == "__main__":
    main()
