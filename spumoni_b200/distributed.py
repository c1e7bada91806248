"""Multi-GPU plumbing: one process per GPU, reads sharded by bases, index replicated.

The reference's only parallelism is an OpenMP region over batches of reads sharing one read-only
index (/root/reference/src/compute_ms_pml.cpp:894-1024).  The multi-GPU equivalent keeps the
same shape: every rank holds the whole flat index (one broadcast at load: NCCL over
NVLink/NVSwitch for CUDA tensors, gloo for the CPU-side tests) and owns a contiguous range of the
read stream; there is no collective in the steady state.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import Index, flat_header_bytes


def shard_by_bases(offsets: np.ndarray, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous read range [r0, r1) of `rank`, balanced by bases rather than read count
    (mixed-length inputs, BASELINE config 5).  Ranges of all ranks tile [0, n_reads)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if n <= 0:
        return 0, 0
    total = int(offsets[-1])
    cuts = [int(np.searchsorted(offsets, (total * r) // world_size, side="left")) for r in range(world_size + 1)]
    cuts[0], cuts[-1] = 0, n
    for i in range(1, world_size + 1):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


def broadcast_index(index: Optional[Index], device: Optional[int] = None, src: int = 0, group=None) -> Index:
    """Replicate the flat index of rank `src` on every rank.

    device is None  -> host path (gloo): every rank ends up with an Index holding a host image.
    device is an int -> CUDA path (NCCL): the image is broadcast GPU-to-GPU straight into a torch
                        buffer that the returned Index adopts (spb_index_attach_device)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    on_gpu = device is not None
    dev = torch.device("cuda", device) if on_gpu else torch.device("cpu")
    size = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        blob = index.flat()
        size[0] = blob.nbytes
    dist.broadcast(size, src=src, group=group)
    nbytes = int(size.item())
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if rank == src:
        buf.copy_(torch.from_numpy(np.asarray(blob).copy()) if not on_gpu else torch.from_numpy(np.array(blob, copy=True)).to(dev))
    dist.broadcast(buf, src=src, group=group)
    if not on_gpu:
        return index if rank == src else Index.from_flat(buf.numpy())
    hdr = buf[: flat_header_bytes()].cpu().numpy()
    out = Index.from_header(hdr)
    out.attach_device(device, buf)
    return out
