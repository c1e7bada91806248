"""Multi-GPU inside ONE process through the C ABI: spb_index_replicate (one broadcast of the flat image) and host queries
that the library shards by bases over the GPUs, one host thread per GPU.  Prints one JSON line with bases/s per transport.
  python tools/replicate_bench.py <n_gpus> [workload] [reps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import spumoni_b200 as sp  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else sp.device_count()
name = sys.argv[2] if len(sys.argv) > 2 else "c2"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else n_gpus  # the read set is tiled so that every GPU gets a full workload
ref = bench.ensure_index(name, "pml")
buf, off = bench.ensure_reads(name, ref, seed=4)
lens = np.tile(np.diff(off.astype(np.int64)), reps)
goff = np.zeros(len(lens) + 1, dtype=np.uint64)
goff[1:] = np.cumsum(lens)
total = int(goff[-1])
lib = sp.load_library()


def pinned(nbytes, dtype):
    import ctypes
    p = ctypes.c_void_p()
    assert lib.spb_host_alloc(ctypes.byref(p), nbytes) == 0
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,)).view(dtype)


reads = pinned(total, np.uint8)
for i in range(reps):
    reads[i * len(buf):(i + 1) * len(buf)] = buf
t0 = time.time()
idx = sp.Index.open(ref, sp.MODE_PML).replicate(list(range(n_gpus)))
t_rep = time.time() - t0
out = {"n_gpus": n_gpus, "workload": name, "bases": total, "replicate_s": t_rep, "method": lib.spb_replicate_method().decode()}
L8 = pinned(total, np.uint8)
bits = pinned(sp.bits_bytes(total), np.uint64)
pk, ep, ev = sp.pack_reads_2bit(reads)
pkp = pinned(pk.nbytes, np.uint8)
pkp[:] = pk
for fmt in ("u8", "bits", "packed_bits"):
    kw = dict(out_lengths=L8, length_bytes=1) if fmt == "u8" else dict(out_lengths=bits, length_bytes=sp.LENGTH_BITS)
    if fmt == "packed_bits":
        kw["packed"] = (pkp, ep, ev)
    src = None if fmt == "packed_bits" else reads
    for _ in range(2):
        idx.query_host(src, goff, **kw)
    t0 = time.perf_counter()
    for _ in range(3):
        idx.query_host(src, goff, **kw)
    out[fmt] = total * 3 / (time.perf_counter() - t0)
# spot check against a single-GPU handle
one = sp.Index.open(ref, sp.MODE_PML).to_device(0)
k = min(len(off) - 1, 20000)
L1, _, _ = one.query_host(buf[: int(off[k])], off[: k + 1], length_bytes=1)
out["equal_single_gpu"] = bool(np.array_equal(L1, L8[: int(off[k])])) and bool(np.array_equal(sp.expand_bits(bits, goff[: k + 1], 1), L1))
print(json.dumps(out))
