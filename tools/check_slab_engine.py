"""Run under torchrun on N GPUs: the C++ z-slab engine with ONE SLAB PER PROCESS (imhd_create_slab, the path bench.py
times at N > 1) must equal the single-GPU context BIT FOR BIT, both pipelines, uneven slabs, thin and overlapped schedules.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_slab_engine.py

IMHD_SLAB_EXCHANGE=nccl selects the ncclSend/ncclRecv exchange instead of the copy-engine one (same bits).
"""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = importlib.import_module("imhd-cuda_b200"); ops = pkg.ops
B = (-3.14159, 3.14159) * 3
ok = True
for (Nx, Ny, Nz, steps) in ((40, 36, 6 * world + 1, 7), (44, 64, 35 * world + 3, 9), (44, 64, 80 * world, 5)):   # thin slabs; overlapped (ends-first) schedule; interior long enough for the strip under the marching kernel
    for path, D in ((pkg.PATH_A, 0.0), (pkg.PATH_B, 0.01)):
        with ops.Context(Nx, Ny, Nz, device=lr) as c:
            c.init_grids(*B); c.init_cubic_bennett_vortex_m0(2.0, 0.5); c.prime(path, D, 1e-4); c.step(steps)
            ref = c.get_state()
        box = [ops.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        with ops.Context.slab(Nx, Ny, Nz, rank, world, lr, box[0]) as s:
            s.init_grids(*B); s.init_cubic_bennett_vortex_m0(2.0, 0.5); s.prime(path, D, 1e-4); s.step(steps)
            k0, nzl, _ = s.slab_extent(0)
            mine = s.get_state_local(0)
        same = bool(np.array_equal(mine.view(np.uint32), ref[:, k0:k0 + nzl].view(np.uint32)))
        t = torch.tensor([int(same)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{Nx}x{Ny}x{Nz} path {'AB'[path]}: {world} slabs in {world} processes vs 1 GPU, {steps} steps: bit-identical = {bool(t.item())}", flush=True)
        ok = ok and bool(t.item())
dist.destroy_process_group()
sys.exit(0 if ok else 1)
