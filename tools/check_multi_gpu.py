"""Run under torchrun on N GPUs: the z-slab decomposed run over NCCL must equal the single-GPU run BIT FOR BIT.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py
"""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = importlib.import_module("imhd-cuda_b200"); ops = pkg.ops
slab = importlib.import_module("imhd-cuda_b200.slab")
Nx, Ny, Nz, steps, dt = 96, 80, 24 * world, 12, 1e-4
b = (-3.14159, 3.14159) * 3
d = tuple(ops.grid_spacing(b[2 * a], b[2 * a + 1], n) for a, n in enumerate((Nx, Ny, Nz)))
gx, gy, gz = ops.init_grids(b, Nx, Ny, Nz)
ok = True
for path, D in ((pkg.PATH_A, 0.0), (pkg.PATH_B, 0.01)):
    Q0 = ops.init_cubic_bennett_vortex_m0(2.0, 0.5, gx, gy, gz)   # z-dependent IC, identical on every rank
    if path == pkg.PATH_A:
        ops.initial_bcs(Q0)
    ce = ops.wall_energy_fixed_point(float(Q0[7, 0, -1, -1]), Nx)
    s = slab.SlabSolver(Nx, Ny, Nz, path, D, dt, *d, comm=slab.TorchComm(), corner_e=ce)
    s.load_global(Q0)
    s.step(steps)
    mine = s.state.clone()
    # single-GPU reference on every rank
    A, B = Q0.clone(), torch.empty_like(Q0)
    for _ in range(steps):
        ops.step_full_domain(A, B, path, D, dt, *d, corner_e=ce); A, B = B, A
    L = s.layout
    same = bool((mine.view(torch.int32) == A[:, L.k0:L.k1].view(torch.int32)).all())
    t = torch.tensor([int(same)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"path {'AB'[path]}: {world} slabs over NCCL vs 1 GPU, {steps} steps on {Nx}x{Ny}x{Nz}: bit-identical = {bool(t.item())}", flush=True)
    ok = ok and bool(t.item())
dist.destroy_process_group()
sys.exit(0 if ok else 1)
