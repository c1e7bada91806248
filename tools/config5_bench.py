"""BASELINE configs[4] in one process: PML + `-c` bin statistics on a large index replicated over N GPUs
(spb_index_replicate), read lengths 150 bp ... 10 kb with equal base counts per length, every length alone and all of them
mixed in one batch (blocks of 64 reads of one length, shuffled).  Prints one JSON line.
  SPB_BENCH_WORKLOAD=c4r python tools/config5_bench.py <n_gpus> [bases per length] [--dry]"""
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

args = [a for a in sys.argv[1:] if not a.startswith("--")]
dry = "--dry" in sys.argv
n_gpus = int(args[0]) if args else 1
per_len = int(args[1]) if len(args) > 1 else 400_000_000
name = os.environ.get("SPB_BENCH_WORKLOAD", "c4r")
LENGTHS = [150, 500, 1000, 5000, 10000]
w = bench.WORKLOADS[name]
t0 = time.time()
ref = bench.ensure_index(name, "pml")
d = bench.cache_dir(name)
t_index = time.time() - t0


def reads_of(L):
    n = per_len // L
    path = os.path.join(d, f"c5_{n}_{L}.bin")
    if not os.path.exists(path + ".off"):
        if "synth" in w:
            subprocess.check_call([bench.MKINDEX, "synth", "--reads-only", "--prefix", os.path.join(d, "idx"), "--n", str(n), "--len", str(L), "--read-seed", str(L),
                                   "--bin", path, "--sub", "0.55", "--neg-frac", "0.5"] + w["synth"], stdout=subprocess.DEVNULL)
        else:
            subprocess.check_call([bench.MKINDEX, "reads", "--ref", ref, "--n", str(n), "--len", str(L), "--seed", str(100 + L), "--bin", path,
                                   "--sub", "0.06", "--ins", "0.02", "--del", "0.02", "--neg-frac", "0.5"], stdout=subprocess.DEVNULL)
    return np.fromfile(path, dtype=np.uint8), np.fromfile(path + ".off", dtype=np.uint64)


t0 = time.time()
sets = {L: reads_of(L) for L in LENGTHS}
# mixed: blocks of 64 consecutive reads of one length, blocks in random order
BLK = 64
blocks = []
for L, (b, o) in sets.items():
    nr = len(o) - 1
    for s in range(0, nr, BLK):
        blocks.append((L, s, min(nr, s + BLK)))
rng = np.random.default_rng(7)
order = rng.permutation(len(blocks))
lens_parts, total = [], 0
for j in order:
    L, s, e = blocks[j]
    o = sets[L][1]
    lens_parts.append(np.diff(o[s:e + 1].astype(np.int64)))
lens = np.concatenate(lens_parts)
moff = np.zeros(len(lens) + 1, dtype=np.uint64)
moff[1:] = np.cumsum(lens)
lib = sp.load_library()


def pinned(nbytes, dtype):
    import ctypes
    if dry:
        return np.zeros(nbytes, dtype=np.uint8).view(dtype)
    p = ctypes.c_void_p()
    assert lib.spb_host_alloc(ctypes.byref(p), max(nbytes, 8)) == 0
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,)).view(dtype)


total = int(moff[-1])
mbuf = pinned(total, np.uint8)
pos = 0
for j in order:
    L, s, e = blocks[j]
    b, o = sets[L]
    a0, a1 = int(o[s]), int(o[e])
    mbuf[pos:pos + (a1 - a0)] = b[a0:a1]
    pos += a1 - a0
assert pos == total
t_reads = time.time() - t0
out = {"config": "BASELINE configs[4] shape: PML + -c statistics, read-length sweep, index replicated over the GPUs of one process",
       "workload": w["desc"], "n_gpus": n_gpus, "bases_per_length": per_len, "stats": "bin_width 150, max_value_thr 9",
       "setup_s": {"index": t_index, "reads": t_reads}, "results": {}}
if dry:
    print(json.dumps({"dry": True, "mixed_reads": len(lens), "mixed_bases": total, "blocks": len(blocks), **out}))
    sys.exit(0)
t0 = time.time()
idx = sp.Index.open(ref, sp.MODE_PML).replicate(list(range(n_gpus)))
out["setup_s"]["open_replicate"] = time.time() - t0
out["replicate_method"] = lib.spb_replicate_method().decode()
out["n_r"] = idx.get_bwt_stats()


def run(buf, off, key):
    nb, nr = int(off[-1]), len(off) - 1
    S = pinned(nr * 16, np.uint8).view(sp.BinStats)
    res = {"reads": nr, "bases": nb}
    for fmt in ("bits", "u16"):
        if fmt == "bits":
            L = pinned(sp.bits_bytes(nb), np.uint64)
            kw = dict(out_lengths=L, length_bytes=sp.LENGTH_BITS)
        else:
            L = pinned(nb * 2, np.uint16)
            kw = dict(out_lengths=L, length_bytes=2)
        for _ in range(2):
            idx.query_host(buf, off, bin_width=150, max_value_thr=9, out_stats=S, **kw)
        t = time.perf_counter()
        for _ in range(3):
            idx.query_host(buf, off, bin_width=150, max_value_thr=9, out_stats=S, **kw)
        res["e2e_%s_stats_bases_per_s" % fmt] = nb * 3 / (time.perf_counter() - t)
        if fmt == "bits":
            res["found_fraction"] = float(np.mean(S["above"] / np.maximum(S["above"] + S["below"], 1) > 0.5))
    out["results"][key] = res


for L in LENGTHS:
    b, o = sets[L]
    pb = pinned(b.nbytes, np.uint8)
    pb[:] = b
    run(pb, o, str(L))
run(mbuf, moff, "mixed")
# spot check of the mixed batch against one GPU
one = sp.Index.open(ref, sp.MODE_PML).to_device(0)
k = min(len(moff) - 1, 3000)
L1, _, S1 = one.query_host(mbuf[: int(moff[k])], moff[: k + 1], bin_width=150, max_value_thr=9, length_bytes=2)
L2, _, S2 = idx.query_host(mbuf[: int(moff[k])], moff[: k + 1], bin_width=150, max_value_thr=9, length_bytes=2)
out["equal_single_gpu"] = bool(np.array_equal(L1, L2)) and bool(np.array_equal(S1, S2))
print(json.dumps(out))
