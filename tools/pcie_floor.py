"""PCIe floor of the end-to-end call at N concurrent ranks (one process per GPU, launched with torch.distributed.run):
every rank moves, at the same time, the bytes one step of each transport moves (H2D and D2H on two streams, pinned host
memory, 32 chunks like the pipelined sub-batches) and the slowest rank sets the time.  Prints, per transport, GB/s per
direction per rank and the bases/s no implementation of that transport can exceed on this box at this N.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_floor.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
BASES = 1_000_000_000
TRANSPORTS = {"u8": (BASES, BASES), "u16": (BASES, 2 * BASES), "bits": (BASES, BASES // 8), "packed_bits": (BASES // 4, BASES // 8)}
h_in = torch.empty(BASES, dtype=torch.uint8).pin_memory()
h_out = torch.empty(2 * BASES, dtype=torch.uint8).pin_memory()
d_in = torch.empty(BASES, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 * BASES, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(nin, nout, chunks=32):
    barrier()
    t = time.perf_counter()
    for i in range(chunks):
        a, b = i * nin // chunks, (i + 1) * nin // chunks
        with torch.cuda.stream(s1):
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        a, b = i * nout // chunks, (i + 1) * nout // chunks
        with torch.cuda.stream(s2):
            h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    barrier()
    return time.perf_counter() - t


out = {"n_ranks": world}
for name, (nin, nout) in TRANSPORTS.items():
    run(nin, nout)
    best = min(run(nin, nout) for _ in range(4))
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    out[name] = {"ms": sec * 1e3, "h2d_GBps_per_rank": nin / sec / 1e9, "d2h_GBps_per_rank": nout / sec / 1e9,
                 "floor_bases_per_s_all_ranks": BASES * world / sec}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
