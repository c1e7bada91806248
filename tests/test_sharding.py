"""CPU: the multi-GPU host logic under gloo with world_size 2 (sharding + index broadcast)."""
import os
import sys

import numpy as np
import pytest

import spumoni_b200 as sp
from spumoni_b200.distributed import shard_by_bases


def test_shard_by_bases_tiles_and_balances():
    rng = np.random.default_rng(1)
    lens = np.concatenate([rng.integers(100, 200, 500), rng.integers(5000, 10000, 40)])
    rng.shuffle(lens)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 4, 8):
        ranges = [shard_by_bases(off, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == len(lens)
        for a, b in zip(ranges, ranges[1:]):
            assert a[1] == b[0]
        bases = [int(off[r1] - off[r0]) for r0, r1 in ranges]
        assert max(bases) - min(bases) <= 2 * lens.max()
    assert shard_by_bases(np.zeros(1, dtype=np.uint64), 4, 2) == (0, 0)
    assert [shard_by_bases(np.array([0, 10], dtype=np.uint64), 2, r) for r in range(2)] in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])


def _worker(rank, world, ref, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from spumoni_b200.distributed import broadcast_index
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sp.Index.open(ref, sp.MODE_PML) if rank == 0 else None
    idx = broadcast_index(idx, device=None, src=0)
    np.save(os.path.join(out_dir, f"blob{rank}.npy"), np.array(idx.flat(), copy=True))
    np.save(os.path.join(out_dir, f"stats{rank}.npy"), np.array(idx.get_bwt_stats()))
    dist.barrier()
    dist.destroy_process_group()


def test_index_broadcast_gloo_world2(small_index, tmp_path):
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as sk:  # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, small_index["ref"], port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = np.load(tmp_path / "blob0.npy"), np.load(tmp_path / "blob1.npy")
    assert b0.nbytes > 4096 and np.array_equal(b0, b1)
    assert np.array_equal(np.load(tmp_path / "stats0.npy"), np.load(tmp_path / "stats1.npy"))
