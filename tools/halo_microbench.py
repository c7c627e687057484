"""Times the host-orchestrated pieces of the slab decomposition (2+ ranks, NCCL):
    torchrun --nproc-per-node 2 tools/halo_microbench.py
"""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
left, right = (rank - 1) % world, (rank + 1) % world


def timeit(name, fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    if rank == 0:
        print(f"{name:60s} {dt:8.3f} ms", flush=True)


for n in (125000, 500000):
    coord = torch.rand(n, device=dev, dtype=torch.float64)
    m = coord < 0.06
    timeit(f"n={n} argsort(stable) of uint8 mask", lambda: torch.argsort((~m).to(torch.uint8), stable=True))
    timeit(f"n={n} nonzero (sync)", lambda: m.nonzero())
    timeit(f"n={n} mask + sum", lambda: (coord < 0.06).sum())
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    allc = torch.empty(2 * world, dtype=torch.int64, device=dev)
    timeit(f"n={n} all_gather 16 B + .cpu()", lambda: (dist.all_gather_into_tensor(allc, counts), allc.cpu()))
    for rows, width, dt in ((7250, 3, torch.float64), (7250, 256, torch.float32), (7250, 128, torch.float32)):
        a, b = torch.zeros((rows, width), dtype=dt, device=dev), torch.zeros((rows, width), dtype=dt, device=dev)
        ra, rb = torch.empty_like(a), torch.empty_like(b)

        def xchg():
            ops = [dist.P2POp(dist.isend, a, left), dist.P2POp(dist.isend, b, right),
                   dist.P2POp(dist.irecv, ra, right), dist.P2POp(dist.irecv, rb, left)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()

        timeit(f"n={n} batch_isend_irecv 2 x {rows} x {width} {str(dt)[6:]}", xchg)
    p = torch.zeros((n, 256), device=dev)
    idx = torch.randint(0, n, (7250,), device=dev)
    timeit(f"n={n} index_select 7250 rows of P", lambda: p.index_select(0, idx))
dist.destroy_process_group()
