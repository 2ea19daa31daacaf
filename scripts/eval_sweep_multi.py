"""BASELINE configs[4] at N GPUs: batched splfe on 1e8 / 1e9 / 1e10 queries IN TOTAL, sharded over the ranks (queries are
independent: no collective on the data path), 1-D..4-D splines, real64 and real32, uniform-random order.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/eval_sweep_multi.py [max_total]

Every rank evaluates its shard device-resident; the time is the MAX over ranks of the CUDA-event time (best of 3 after a
warm-up), bracketed by barriers.  Rank 0 prints one markdown row per (ndim, precision, total queries)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import splpak_b200 as sp
from splpak_b200 import synth

GRIDS = {1: [50], 2: [64, 64], 3: [24, 24, 24], 4: [12, 12, 12, 12]}
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
max_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000_000
HBM = 6534.5


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if rank == 0:
    print(f"| GPUs | ndim | nodes | real | total queries | per GPU | ms (max over ranks) | total Gq/s | per-GPU algorithmic GB/s | of HBM ({HBM:.0f} GB/s) |")
    print("|---:|---:|---|---|---:|---:|---:|---:|---:|---:|", flush=True)
for ndim in (1, 2, 3, 4):
    nodes = GRIDS[ndim]
    ncol = 1
    for n in nodes:
        ncol *= n
    for real32 in (False, True):
        dt = torch.float32 if real32 else torch.float64
        s = 4 if real32 else 8
        coef = torch.randn(ncol, dtype=dt, device=dev)
        for total in (100_000_000, 1_000_000_000, 10_000_000_000):
            if total > max_total:
                continue
            nq = total // world
            if nq * (ndim + 1) * s > 110e9:                 # shard must fit one 180 GB GPU with head-room
                if rank == 0:
                    print(f"| {world} | {ndim} | {'x'.join(map(str, nodes))} | {'real32' if real32 else 'real64'} | {total:.0e} | {nq:.2e} | "
                          f"— (shard of {nq * (ndim + 1) * s / 1e9:.0f} GB does not fit) | | | |", flush=True)
                continue
            q = synth.queries_torch(ndim, nq, start=rank * nq, seed=43, device=dev, dtype=dt)
            out = torch.empty(nq, dtype=dt, device=dev)
            best = 1e30
            for rep in range(4):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ierr = sp.eval_batch_device(ndim, q, ndim, nq, coef, [0.0] * ndim, [1.0] * ndim, nodes, out,
                                            stream=torch.cuda.current_stream(), real32=real32)
                e1.record()
                barrier()
                assert ierr == 0
                ms = rmax(e0.elapsed_time(e1))
                if rep:
                    best = min(best, ms)
            if rank == 0:
                gbs = nq * (ndim + 1) * s / best / 1e6
                print(f"| {world} | {ndim} | {'x'.join(map(str, nodes))} | {'real32' if real32 else 'real64'} | {total:.0e} | {nq:.2e} | "
                      f"{best:.3f} | {total / best / 1e6:.1f} | {gbs:.0f} | {100 * gbs / HBM:.1f} % |", flush=True)
            del q, out
if world > 1:
    dist.destroy_process_group()
