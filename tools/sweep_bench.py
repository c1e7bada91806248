"""Read-length sweep (BASELINE configs[4]): PML + `-c` bin statistics for reads of 150 bp ... 10 kb, equal base counts per
length, each length alone and all of them shuffled into one batch (the case the longest-first order and the sub-batch
sizing by reads exist for).  Prints one JSON line.
  [SPB_BENCH_WORKLOAD=c2|c4s] python tools/sweep_bench.py [bases per length]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import spumoni_b200 as sp  # noqa: E402
import torch  # noqa: E402

name = os.environ.get("SPB_BENCH_WORKLOAD", "c2")
per_len = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
LENGTHS = [150, 500, 1000, 5000, 10000]
w = bench.WORKLOADS[name]
ref = bench.ensure_index(name, "pml")
d = bench.cache_dir(name)


def reads_of(L):
    n = per_len // L
    path = os.path.join(d, f"sweep_{n}_{L}.bin")
    if not os.path.exists(path + ".off"):
        if "synth" in w:
            subprocess.check_call([bench.MKINDEX, "synth", "--reads-only", "--prefix", os.path.join(d, "idx"), "--n", str(n), "--len", str(L), "--read-seed", str(L),
                                   "--bin", path, "--sub", "0.55", "--neg-frac", "0.5"] + w["synth"], stdout=subprocess.DEVNULL)
        else:
            subprocess.check_call([bench.MKINDEX, "reads", "--ref", ref, "--n", str(n), "--len", str(L), "--seed", str(100 + L), "--bin", path,
                                   "--sub", "0.06", "--ins", "0.02", "--del", "0.02", "--neg-frac", "0.5"], stdout=subprocess.DEVNULL)
    return np.fromfile(path, dtype=np.uint8), np.fromfile(path + ".off", dtype=np.uint64)


sets = {L: reads_of(L) for L in LENGTHS}
# all lengths in one batch, reads in random order
rng = np.random.default_rng(7)
chunks = []
for L, (b, o) in sets.items():
    for i in range(len(o) - 1):
        chunks.append((L, i))
order = rng.permutation(len(chunks))
lens = np.array([chunks[j][0] for j in order], dtype=np.int64)
moff = np.zeros(len(lens) + 1, dtype=np.uint64)
moff[1:] = np.cumsum(lens)
mbuf = np.empty(int(moff[-1]), dtype=np.uint8)
for pos, j in enumerate(order):
    L, i = chunks[j]
    b, o = sets[L]
    mbuf[int(moff[pos]):int(moff[pos + 1])] = b[int(o[i]):int(o[i + 1])]
sets["mixed"] = (mbuf, moff)

idx = sp.Index.open(ref, sp.MODE_PML).to_device(0)
dev = torch.device("cuda", 0)
out = {"workload": w["desc"], "bases_per_length": per_len, "stats": "bin_width 150, max_value_thr 9 (-c)", "results": {}}
for key, (buf, off) in sets.items():
    nb, nr = int(off[-1]), len(off) - 1
    h_reads = torch.from_numpy(buf).pin_memory()
    h_off = torch.from_numpy(off.astype(np.int64)).pin_memory()
    d_reads = torch.zeros(nb + 512, dtype=torch.uint8, device=dev)
    d_reads[:nb].copy_(h_reads)
    d_off = h_off.to(dev)
    d_len = torch.zeros(nb, dtype=torch.int16, device=dev)
    d_stats = torch.zeros(nr * 2, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    for _ in range(2):
        idx.query_device(d_reads, d_off, nr, d_len, d_stats=d_stats, bin_width=150, max_value_thr=9, stream=stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(3):
        idx.query_device(d_reads, d_off, nr, d_len, d_stats=d_stats, bin_width=150, max_value_thr=9, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    dev_rate = nb * 3 / (e0.elapsed_time(e1) / 1e3)
    L16 = torch.empty(nb, dtype=torch.int16).pin_memory().numpy().view(np.uint16)
    S = torch.zeros(nr * 2, dtype=torch.int64).pin_memory().numpy().view(sp.BinStats)  # pinned: a pageable output makes every D2H synchronous
    np_reads, np_off = h_reads.numpy(), h_off.numpy().view(np.uint64)
    for _ in range(2):
        idx.query_host(np_reads, np_off, bin_width=150, max_value_thr=9, out_lengths=L16, out_stats=S, length_bytes=2)
    t0 = time.perf_counter()
    for _ in range(3):
        idx.query_host(np_reads, np_off, bin_width=150, max_value_thr=9, out_lengths=L16, out_stats=S, length_bytes=2)
    e2e = nb * 3 / (time.perf_counter() - t0)
    same = bool(np.array_equal(L16, d_len.cpu().numpy().view(np.uint16))) and bool(np.array_equal(S.view(np.uint64).reshape(-1), d_stats.cpu().numpy().view(np.uint64)))
    out["results"][str(key)] = {"reads": nr, "bases": nb, "device_bases_per_s": dev_rate, "e2e_u16_stats_bases_per_s": e2e, "host_equals_device": same}
    del d_reads, d_len, d_stats, h_reads
print(json.dumps(out))
