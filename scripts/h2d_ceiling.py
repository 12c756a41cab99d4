"""What this box can move between pinned host memory and its GPUs with NO kernel running:

    python scripts/h2d_ceiling.py [points_per_gpu]                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_ceiling.py

Every rank copies the bench step's coordinates host -> device (16 B per point) and its values device -> host (8 B
per point) on two streams, all ranks at once, max over ranks — the ceiling bench.py reports as e2e.copy_ceiling and
divides its end-to-end figure by.  Also prints the one-direction rates."""
import os, sys, time
import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
xh, oh = torch.empty((n, 2), dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
xd, od = torch.empty((n, 2), dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    def once():
        if h2d:
            with torch.cuda.stream(s1):
                xd.copy_(xh, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                oh.copy_(od, non_blocking=True)
        s1.synchronize(); s2.synchronize()
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


for name, h, d, nbytes in (("H2D only", True, False, 16 * n), ("D2H only", False, True, 8 * n), ("H2D + D2H", True, True, 24 * n)):
    dt = run(h, d)
    if rank == 0:
        print(f"{world} GPU(s) {name:10s}: {dt * 1e3:8.2f} ms/step  {nbytes * world / dt / 1e9:7.1f} GB/s aggregate  "
              f"{n * world / dt / 1e9:6.3f} G points/s", flush=True)
if world > 1:
    dist.destroy_process_group()
