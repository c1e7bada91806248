#!/usr/bin/env python
"""bench.py -- PML-mode classification throughput (bases/sec) of the B200 hot path.

Workload (BASELINE.json configs[1]): PML on a 100 Mb pangenome index (20 x 5 Mb haplotypes, 0.1 % SNPs,
0.01 % indels, + reverse complement, n ~ 2e8), 1 M synthetic 1 kb ONT-like reads (6 % sub, 2 % ins, 2 % del,
10 % negatives) per GPU.  A "step" is one pass of the hot path over the rank's 1 M reads (1e9 bases).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [--gpus N] ...              the reference's own CPU code (oracle/_ref/spumoni_ref: its translation
                                                               units over stand-in third-party containers) when that binary was
                                                               built, else the oracle port; all host threads

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
MKINDEX = os.path.join(ROOT, "tools", "bin", "spumoni_mkindex")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "spumoni_ref")  # built where /root/reference exists; travels prebuilt to the GPU box

BYTES_PER_BASE = {"pml": 69,   # 1 B read + 4 B length + 2 x 32 B index sectors
                  "ms": 117}   # + 8 B pointer + 8 B SA sample + 32 B text sector
METRICS = {"pml": "bases/sec PML-mode classification", "ms": "bases/sec MS-mode (-M) matching statistics"}
METRIC = METRICS["pml"]

WORKLOADS = {
    # name: (genome args, reads per rank, read length, read args)
    "c2": dict(genome=["--kind", "pangenome", "--len", "100000000", "--haps", "20", "--snp", "0.001", "--indel", "0.0001", "--seed", "3"],
               n_reads=1_000_000, read_len=1000, reads=["--sub", "0.06", "--ins", "0.02", "--del", "0.02", "--neg-frac", "0.1"],
               desc="PML, 100 Mb pangenome index (20x5 Mb, +revcomp), 1 M x 1 kb ONT-like reads per GPU"),
    # Stand-in for BASELINE configs[3] (3 Gb pangenome, n > 2^32), NOT a bench line of this round: a synthetic run-length BWT
    # (tools/mkindex synth: no suffix array, not the BWT of a text) with the table sizes of such an index
    # -- r = 1.5e8 runs: 16-byte move rows, 2.4 GB row table >> L2 -- and reads driven by the search itself, 55 % jumps like
    # the c2 reads. PML only (there is no text for MS). SPB_BENCH_WORKLOAD=c4s
    "c4s": dict(synth=["--runs", "150000000", "--mean-len", "40", "--seed", "5"],
                n_reads=1_000_000, read_len=1000, reads=["--sub", "0.55", "--neg-frac", "0.1"],
                desc="PML, synthetic RLBWT stand-in for config 4 (n ~ 6e9 > 2^32, r = 1.5e8), 1 M x 1 kb search-driven reads per GPU"),
    # Real-text stand-in for BASELINE configs[3]: a 500 Mb pangenome (5 haplotypes x 100 Mb, + reverse complement: n = 1e9 + 1),
    # suffix-sorted ON THE GPU (spb_build_parts_device, build_kernels.cu), serialised by the reference's build side (mkindex
    # from-parts).  The BWT of a real text with r ~ 1.5e8 runs: the row table (32-byte rows) is ~5 GB, far beyond the L2, like a
    # human pangenome index -- but n < 2^31 (the 32-bit device builder); positions beyond 2^32 are covered by c4s.  SPB_BENCH_WORKLOAD=c4r
    "c4r": dict(genome=["--kind", "pangenome", "--len", "500000000", "--haps", "5", "--snp", "0.001", "--indel", "0.0001", "--seed", "5"], gpu_build=True,
                n_reads=1_000_000, read_len=1000, reads=["--sub", "0.06", "--ins", "0.02", "--del", "0.02", "--neg-frac", "0.1"],
                desc="PML, 500 Mb pangenome index built on the GPU (5x100 Mb, +revcomp, n = 1e9, real BWT with r ~ 1.5e8), 1 M x 1 kb ONT-like reads per GPU"),
    "c4r_dev": dict(genome=["--kind", "pangenome", "--len", "4000000", "--haps", "4", "--snp", "0.001", "--indel", "0.0001", "--seed", "5"], gpu_build=True,
                    n_reads=20_000, read_len=1000, reads=["--sub", "0.06", "--ins", "0.02", "--del", "0.02", "--neg-frac", "0.1"],
                    desc="DEV ONLY: PML, 4 Mb pangenome index built on the GPU, 20 k x 1 kb reads per GPU"),
    "c4s_dev": dict(synth=["--runs", "2000000", "--mean-len", "40", "--seed", "5"],
                    n_reads=20_000, read_len=1000, reads=["--sub", "0.55", "--neg-frac", "0.1"],
                    desc="DEV ONLY: PML, synthetic RLBWT (r = 2e6), 20 k x 1 kb search-driven reads per GPU"),
    # development-size stand-in (NOT a bench line): SPB_BENCH_WORKLOAD=dev
    "dev": dict(genome=["--kind", "pangenome", "--len", "4000000", "--haps", "8", "--snp", "0.001", "--indel", "0.0001", "--seed", "3"],
                n_reads=50_000, read_len=1000, reads=["--sub", "0.06", "--ins", "0.02", "--del", "0.02", "--neg-frac", "0.1"],
                desc="DEV ONLY: PML, 4 Mb pangenome index, 50 k x 1 kb reads per GPU"),
}


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: everything else that writes to fd 1 (the NCCL version banner, child
# processes) is sent to stderr, and the result line goes to a private duplicate of the original stdout.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    print(json.dumps(obj), file=_RESULT_OUT, flush=True)


def cache_dir(name: str) -> str:
    w = WORKLOADS[name]
    key = hashlib.sha1(json.dumps([w.get("genome") or w["synth"], "fixture-v1"]).encode()).hexdigest()[:12]
    base = os.environ.get("SPUMONI_B200_CACHE", "/tmp/spumoni_b200_cache")
    d = os.path.join(base, f"{name}_{key}")
    os.makedirs(d, exist_ok=True)
    return d


def ensure_index(name: str, mode: str = "pml") -> str:
    """Untimed setup: synthetic reference + `spumoni build` stand-in (tools/). Returns ref_file (<prefix>.fa)."""
    d = cache_dir(name)
    ref = os.path.join(d, "idx.fa")
    ext = ".thrbv.spumoni" if mode == "pml" else ".thrbv.ms"
    if os.path.exists(ref + ext) and os.path.exists(os.path.join(d, "ok_" + mode)):
        return ref
    t0 = time.time()
    w = WORKLOADS[name]
    if "synth" in w:  # index and reads come out of one pass over the synthetic run arrays
        if mode != "pml":
            raise SystemExit("bench.py: synthetic workloads have no text: PML only")
        tmp = os.path.join(d, f"reads.tmp{os.getpid()}")
        subprocess.check_call([MKINDEX, "synth", "--prefix", os.path.join(d, "idx"), "--n", str(w["n_reads"]), "--len", str(w["read_len"]),
                               "--bin", tmp] + w["synth"] + w["reads"], stdout=subprocess.DEVNULL)
        path = os.path.join(d, f"reads_{w['n_reads']}_{w['read_len']}_s4.bin")
        os.replace(tmp + ".off", path + ".off")
        os.replace(tmp, path)
        open(os.path.join(d, "ok_" + mode), "w").write("ok")
        log(f"built synthetic index {name} in {time.time() - t0:.1f} s")
        return ref
    subprocess.check_call([MKINDEX, "genome", "--out", os.path.join(d, "raw.fa")] + WORKLOADS[name]["genome"], stdout=subprocess.DEVNULL)
    if w.get("gpu_build"):  # text -> run-length BWT / thresholds / samples on the GPU, then the reference's constructors + serialize
        import spumoni_b200 as sp
        txt = os.path.join(d, "idx.txt")
        subprocess.check_call([MKINDEX, "build", "--input", os.path.join(d, "raw.fa"), "--prefix", os.path.join(d, "idx"), "--text-only", "--dump-text", txt], stdout=subprocess.DEVNULL)
        t1 = time.time()
        P = sp.build_parts_device(np.fromfile(txt, dtype=np.uint8), int(os.environ.get("LOCAL_RANK", "0")), write_to=ref)
        log(f"GPU build of {name}: n={P['n']} r={P['r']} in {time.time() - t1:.1f} s (incl. file I/O of the intermediates)")
        os.remove(txt)
        subprocess.check_call([MKINDEX, "from-parts", "--prefix", os.path.join(d, "idx"), "--" + mode], stdout=subprocess.DEVNULL)
        for ext in (".bwt.heads", ".bwt.len", ".thr_pos", ".ssa", ".esa"):
            os.remove(ref + ext)
    else:
        subprocess.check_call([MKINDEX, "build", "--input", os.path.join(d, "raw.fa"), "--prefix", os.path.join(d, "idx"), "--" + mode], stdout=subprocess.DEVNULL)
    os.remove(os.path.join(d, "raw.fa"))
    open(os.path.join(d, "ok_" + mode), "w").write("ok")
    log(f"built index {name} in {time.time() - t0:.1f} s")
    return ref


def ensure_reads(name: str, ref: str, seed: int):
    w = WORKLOADS[name]
    path = os.path.join(cache_dir(name), f"reads_{w['n_reads']}_{w['read_len']}_s{seed}.bin")
    if not (os.path.exists(path) and os.path.exists(path + ".off")):
        if "synth" in w:
            raise SystemExit("bench.py: the reads of a synthetic workload are written with its index; remove the cache directory")
        tmp = path + f".tmp{os.getpid()}"
        subprocess.check_call([MKINDEX, "reads", "--ref", ref, "--n", str(w["n_reads"]), "--len", str(w["read_len"]), "--seed", str(seed),
                               "--bin", tmp] + w["reads"], stdout=subprocess.DEVNULL)
        os.replace(tmp + ".off", path + ".off")
        os.replace(tmp, path)
    return np.fromfile(path, dtype=np.uint8), np.fromfile(path + ".off", dtype=np.uint64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local: int):
    """Run this rank (and first-touch its pinned buffers) on the host NUMA node its GPU hangs off: the end-to-end leg is a
    PCIe + host-memory path.  Best effort; returns a short description for the log."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return f"gpu {local} ({bdf}): no NUMA affinity reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return f"gpu {local} ({bdf}): bound to NUMA node {node} ({len(cpus)} cpus)"
    except Exception as e:  # unknown topology: leave the affinity alone
        return f"gpu {local}: NUMA binding skipped ({type(e).__name__}: {e})"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_binary_seconds(ref, buf, off, n_reads, mode, threads):
    """One `spumoni run -n -t threads` of the REFERENCE's own code (oracle/_ref/spumoni_ref) over the first n_reads reads:
    FASTA in, text files out, as the reference does it. Returns the seconds of its own "processing the patterns/reads"
    timer (compute_ms_pml.cpp:1324-1337, :1365-1378; index loading is outside it), or None without the binary."""
    if not os.path.exists(REF_BIN):
        return None
    d = os.path.dirname(ref)
    fa = os.path.join(d, f"sample_{n_reads}.fa")
    if not os.path.exists(fa):
        with open(fa + ".tmp", "wb") as f:
            for i in range(n_reads):
                f.write(b">r%d\n" % i)
                f.write(buf[int(off[i]): int(off[i + 1])].tobytes())
                f.write(b"\n")
        os.replace(fa + ".tmp", fa)
    db = ref + (".pmlnulldb" if mode == "pml" else ".msnulldb")  # loaded by every run, used only with -c
    if not os.path.exists(db):
        stats = os.path.join(d, "null_stats.bin")
        ((np.arange(1, 2000, dtype=np.uint64) * 7) % 23).tofile(stats)
        subprocess.check_call([MKINDEX, "nulldb", "--stats", stats, "--out", db], stdout=subprocess.DEVNULL)
    if mode == "ms" and not os.path.exists(ref + ".slp"):  # the stand-in for the SLP reads the plain text
        with open(ref) as src, open(ref + ".slp.tmp", "w") as dst:
            for line in src:
                if not line.startswith(">"):
                    dst.write(line.strip())
        os.replace(ref + ".slp.tmp", ref + ".slp")
    r = subprocess.run([REF_BIN, "run", "-r", ref[:-3], "-p", fa, "-P" if mode == "pml" else "-M", "-n", "-t", str(threads)],
                       capture_output=True, text=True)
    for ext in (".pseudo_lengths", ".lengths", ".pointers"):
        if os.path.exists(fa + ext):
            os.remove(fa + ext)
    m = re.search(r"processing the (?:patterns|reads)[^(]*\(([0-9.]+) sec\)", r.stderr)
    if r.returncode != 0 or not m:
        raise SystemExit("bench.py: oracle/_ref/spumoni_ref failed:\n" + r.stderr[-2000:])
    return float(m.group(1))


def run_reference(args, name):
    """Reference arm: the reference's own CPU code on all host threads (oracle/_ref/spumoni_ref, `kind: "reference"`) when
    that binary exists, else the oracle port of compute_ms_pml.cpp:237-286 (`kind: "port"`). Each step is a bounded
    sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    mode = args.mode
    ref = ensure_index(name, mode)
    buf, off = ensure_reads(name, ref, seed=4)
    cores = os.cpu_count() or 1
    w = WORKLOADS[name]
    sample_reads = min(len(off) - 1, max(200, (1000 if mode == "pml" else 400) * cores * 1000 // w["read_len"]))
    idx = oracle.OracleIndex(ref, mode)
    sb, so = buf[: int(off[sample_reads])], off[: sample_reads + 1]
    use_bin = os.path.exists(REF_BIN)
    for _ in range(args.warmup):
        if use_bin:
            reference_binary_seconds(ref, buf, off, max(1, sample_reads // 10), mode, cores)
        else:
            idx.query_batch(sb[: int(so[max(1, sample_reads // 10)])], so[: max(1, sample_reads // 10) + 1], threads=cores, want_output=False)
    secs = []
    for _ in range(args.steps):
        if use_bin:
            secs.append(reference_binary_seconds(ref, buf, off, sample_reads, mode, cores))
        else:
            secs.append(idx.query_batch(sb, so, threads=cores, want_output=False)[0])
    total = sum(secs)
    value = int(so[-1]) * args.steps / total
    port_sec = idx.query_batch(sb, so, threads=cores, want_output=False)[0] if use_bin else None
    sample = (f"first {sample_reads} reads of the workload, {args.steps} steps; " +
              ("`spumoni run -n -t %d` of the reference's own translation units compiled against stand-in third-party containers "
               "(oracle/shim), FASTA in / text out, its own phase timer" % cores if use_bin
               else "oracle port, in-memory, OpenMP dynamic over reads"))
    out = {"impl": "reference", "metric": METRICS[mode], "value": value, "unit": "bases/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": w["desc"].replace("PML", mode.upper()), "sample": f"first {sample_reads} reads ({int(so[-1])} bases) per step"},
           "cpu_baseline": {"value": value, "unit": "bases/s", "cores": cores, "kind": "reference" if use_bin else "port", "sample": sample},
           "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if port_sec:
        out["oracle_port"] = {"value": int(so[-1]) / port_sec, "unit": "bases/s", "cores": cores, "sample": "same reads, in-memory, one step"}
    emit(out)


def cli_seconds(ref, name, seed, threads, extra_env=None, fastq=False):
    """`spumoni run -P -n -c -t threads` of the product CLI on the workload's reads as a FASTA file: the seconds of the phase the
    reference times ("processing the patterns", compute_ms_pml.cpp:1324-1337; index loading is outside it) + the per-stage
    busy times the CLI prints with SPUMONI_B200_TIMING."""
    import spumoni_b200 as sp
    w = WORKLOADS[name]
    d = cache_dir(name)
    fa = os.path.join(d, f"reads_s{seed}" + ("_fq.fa" if fastq else ".fa"))  # the reference insists on a FASTA extension
    if not os.path.exists(fa):
        if "synth" in w:
            return None
        subprocess.check_call([MKINDEX, "reads", "--ref", ref, "--n", str(w["n_reads"]), "--len", str(w["read_len"]), "--seed", str(seed),
                               "--out", fa + ".tmp"] + (["--fastq", "1"] if fastq else []) + w["reads"], stdout=subprocess.DEVNULL)
        os.replace(fa + ".tmp", fa)
    db = ref + ".pmlnulldb"
    if not os.path.exists(db):
        stats = os.path.join(d, "null_stats.bin")
        ((np.arange(1, 2000, dtype=np.uint64) * 7) % 23).tofile(stats)
        subprocess.check_call([MKINDEX, "nulldb", "--stats", stats, "--out", db], stdout=subprocess.DEVNULL)
    env = dict(os.environ, SPUMONI_B200_TIMING="1", **(extra_env or {}))
    t0 = time.perf_counter()
    r = subprocess.run([sp.CLI_PATH, "run", "-r", ref[:-3], "-p", fa, "-P", "-n", "-c", "-t", str(threads)], capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    err = re.sub(r"\x1b\[[0-9;]*m", "", r.stderr)
    m = None
    for m in re.finditer(r"done\.\s+\(([0-9.]+) sec\)", err):  # the last one closes "processing the patterns"
        pass
    if r.returncode != 0 or not m:
        log("CLI run failed:\n" + err[-1500:])
        return None
    out = {"phase_s": float(m.group(1)), "wall_s": wall, "threads": threads, "bytes_out": os.path.getsize(fa + ".pseudo_lengths")}
    st = re.search(r"busy seconds per stage: (.*)", err)
    if st:
        out["stages"] = st.group(1).strip()
    g = re.search(r"\[timing\] devices: (.*)", err)
    if g:
        out["devices"] = g.group(1).strip()
    for ext in (".pseudo_lengths", ".report"):
        if os.path.exists(fa + ext):
            os.remove(fa + ext)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lanes", type=int, default=0, help="lanes per read (tuning; 0 = library default)")
    ap.add_argument("--format", default=None, choices=["packed_bits", "bits", "u8", "u16", "u32"],
                    help="transport of the headline measurement. packed_bits (default for PML): reads travel as 2 bits per base + an exception list "
                         "(spb_pml_query_host_packed; the caller packs while it parses, spb_pack_reads) and the PMLs come back as one bit per base "
                         "(SPB_LENGTH_BITS: lossless for any read length); bits: bytes in, bits out; u8 / u16 / u32: bytes in, lengths of that width out "
                         "(default for --mode ms: u8)")
    ap.add_argument("--mode", default="pml", choices=["pml", "ms"], help="pml = the headline metric (BASELINE configs[1]); ms = configs[2]")
    ap.add_argument("--pointers64", action="store_true", help="--mode ms: keep 64-bit pointers even when 32 bits are lossless")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the secondary measurements (other output formats, MS, -c statistics, CLI)")
    ap.add_argument("--variants", default="u8,u16,u32,bits,packed_bits,pml_stats,ms,cli,cli_fastq", help="which secondary measurements to run (comma-separated)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: a fixed set of 5 x the workload's reads is sharded by bases over the ranks (BASELINE configs[3])")
    args = ap.parse_args()
    want = set() if args.no_variants else set(v for v in args.variants.split(",") if v)
    main_fmt = args.format or ("packed_bits" if args.mode == "pml" else "u8")
    if main_fmt in ("bits", "packed_bits") and args.mode == "ms":
        raise SystemExit("bench.py: one bit per base describes PML lengths only")
    args.length_bytes = {"u8": 1, "u16": 2, "u32": 4, "bits": 0, "packed_bits": 0}[main_fmt]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    name = os.environ.get("SPB_BENCH_WORKLOAD", "c2")
    w = WORKLOADS[name]
    if not os.path.exists(MKINDEX):
        import __graft_entry__
        __graft_entry__.build()
    if args.impl == "reference":
        return run_reference(args, name)

    import torch
    import torch.distributed as dist

    import spumoni_b200 as sp
    from spumoni_b200.distributed import broadcast_index

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; spumoni_b200 has no CPU path to measure")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    log(f"rank {rank}: " + bind_to_gpu_numa_node(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- untimed setup: index (rank 0 builds + parses, NCCL broadcast), reads per rank
    mode = args.mode
    if rank == 0:
        ref = ensure_index(name, mode)
    barrier()
    ref = os.path.join(cache_dir(name), "idx.fa")

    def open_index(m):
        MODE = sp.MODE_PML if m == "pml" else sp.MODE_MS
        t0 = time.time()
        if world > 1:
            ix = broadcast_index(sp.Index.open(ref, MODE) if rank == 0 else None, device=local, src=0)
        else:
            ix = sp.Index.open(ref, MODE).to_device(local)
        log(f"rank {rank}: {m} index resident in {time.time() - t0:.1f} s")
        return ix

    idx = open_index(mode)
    n, r = idx.get_bwt_stats()
    if args.strong:  # one fixed read set for any N: 5 x the workload's reads, this rank's shard by bases
        b1, o1 = ensure_reads(name, ref, seed=4)
        reps = 5
        lens = np.tile(np.diff(o1.astype(np.int64)), reps)
        goff = np.zeros(len(lens) + 1, dtype=np.uint64)
        goff[1:] = np.cumsum(lens)
        cuts = sp.shard_by_bases(goff, world)
        r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
        off = (goff[r0:r1 + 1] - goff[r0]).astype(np.uint64)
        big = np.tile(b1, reps)
        buf = big[int(goff[r0]):int(goff[r1])].copy()
        del big
        total_reads_job, total_bases_job = len(lens), int(goff[-1])
    else:
        buf, off = ensure_reads(name, ref, seed=4 + rank)
        total_reads_job, total_bases_job = None, None
    n_reads, n_bases = len(off) - 1, int(off[-1])
    job_bases = sum_over_ranks(n_bases)

    h_reads = torch.from_numpy(buf).pin_memory()
    h_off = torch.from_numpy(off.astype(np.int64)).pin_memory()
    d_reads = torch.zeros(n_bases + 512, dtype=torch.uint8, device=dev)
    d_reads[:n_bases].copy_(h_reads)
    d_off = h_off.to(dev)
    stream = torch.cuda.current_stream()
    np_reads, np_off = h_reads.numpy(), h_off.numpy().view(np.uint64)
    TDT = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
    NDT = {1: np.uint8, 2: np.uint16, 4: np.uint32}
    packed_host = {}

    def measure(ix, fmt, m="pml", stats=False, device_leg=True, steps=None):
        """Device-resident and end-to-end timing of one transport.  fmt: "u8" | "u16" | "u32" | "bits" (one bit per base) |
        "packed_bits" (end to end the reads travel as 2 bits per base and are unpacked on the device; the device-resident leg is
        the bits kernel on the unpacked resident bytes).  Returns a dict + the host outputs."""
        steps = steps or args.steps
        bits = fmt in ("bits", "packed_bits")
        width = {"u8": 1, "u16": 2, "u32": 4}.get(fmt, 0)
        if bits:
            nwords = sp.bits_bytes(n_bases) // 8
            h_len = torch.zeros(nwords, dtype=torch.int64).pin_memory()
            d_len = torch.zeros(nwords, dtype=torch.int64, device=dev)
            np_len = h_len.numpy().view(np.uint64)
            out_bytes = nwords * 8
        else:
            h_len = torch.empty(n_bases, dtype=TDT[width]).pin_memory()
            d_len = torch.zeros(n_bases, dtype=TDT[width], device=dev)
            np_len = h_len.numpy().view(NDT[width])
            out_bytes = n_bases * width
        d_sat = torch.zeros(1, dtype=torch.int64, device=dev)
        # MS pointers travel as uint32 when the text has fewer than 2^32 - 1 positions (checked lossless like the lengths)
        pw = 4 if (m == "ms" and n < (1 << 32) - 1 and not args.pointers64) else 8
        d_ptr = torch.zeros(n_bases, dtype=torch.int64 if pw == 8 else torch.int32, device=dev) if m == "ms" else None
        h_ptr = torch.empty(n_bases, dtype=torch.int64 if pw == 8 else torch.int32).pin_memory() if m == "ms" else None
        np_ptr = h_ptr.numpy().view(np.uint64 if pw == 8 else np.uint32) if m == "ms" else None
        d_stats = torch.zeros(n_reads * 2, dtype=torch.int64, device=dev) if stats else None
        h_stats = torch.zeros(n_reads * 2, dtype=torch.int64).pin_memory().numpy().view(sp.BinStats) if stats else None  # pinned like every other buffer
        kw = dict(bin_width=150, max_value_thr=9) if stats else {}
        packed = None
        if fmt == "packed_bits":
            if "p" not in packed_host:  # untimed: the caller packs while it parses; pinned like the byte reads
                pk, ep, ev = sp.pack_reads_2bit(np_reads)
                packed_host["p"] = (torch.from_numpy(pk).pin_memory(), torch.from_numpy(ep.astype(np.int64)).pin_memory(), torch.from_numpy(ev).pin_memory())
            tp = packed_host["p"]
            packed = (tp[0].numpy(), tp[1].numpy().view(np.uint64), tp[2].numpy())
        h2d = (packed[0].nbytes + 9 * packed[1].size if packed else n_bases) + 8 * (n_reads + 1)

        def step_device():
            if bits:
                d_len.zero_()
            ix.query_device(d_reads, d_off, n_reads, d_len, d_pointers=d_ptr, d_stats=d_stats, lanes=args.lanes, stream=stream.cuda_stream, d_saturated=d_sat, **kw)

        def step_host():
            # the call a user makes: host buffers in, host buffers out
            ix.query_host(None if packed else np_reads, np_off, lanes=args.lanes, out_lengths=np_len, out_pointers=np_ptr, out_stats=h_stats,
                          length_bytes=sp.LENGTH_BITS if bits else width, pointer_bytes=pw, packed=packed, **kw)

        res = {"format": fmt, "d2h_bytes_per_step": out_bytes + (pw * n_bases if m == "ms" else 0) + (16 * n_reads if stats else 0), "h2d_bytes_per_step": h2d}
        if device_leg:
            for _ in range(args.warmup):
                step_device()
            barrier()
            launches0 = sp.kernel_launches()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            evs[0].record(stream)
            for i in range(steps):
                step_device()
                evs[i + 1].record(stream)
            barrier()
            res["launches"] = int(sp.kernel_launches() - launches0)
            kern_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
            total_ms_max = max_over_ranks(evs[0].elapsed_time(evs[-1]))
            res.update(value=job_bases * steps / (total_ms_max / 1e3), ms_per_step=total_ms_max / steps, kernel_ms=sum(kern_ms) / len(kern_ms),
                       clamped=int(d_sat.item()) // (steps + args.warmup))
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        barrier()
        res["e2e"] = job_bases * steps / max_over_ranks(time.perf_counter() - t0)
        if device_leg:
            dl = d_len.cpu().numpy()
            res["same"] = bool(np.array_equal(np_len, dl.view(np_len.dtype)))
            if m == "ms":
                res["same"] = res["same"] and bool(np.array_equal(np_ptr, d_ptr.cpu().numpy().view(np_ptr.dtype)))
        return res, np_len, np_ptr

    # ---------------- timed region (clock sampling covers the headline measurement)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main_m, np_len, np_ptr = measure(idx, main_fmt, mode)
    main_bits = main_fmt in ("bits", "packed_bits")
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- parity spot check + CPU baseline (rank 0, N=1)
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
        cores = os.cpu_count() or 1
        sample_reads = min(n_reads, max(200, (7000 if mode == "pml" else 6000) * cores * 1000 // w["read_len"]))  # ~10 s of CPU work
        o = oracle.OracleIndex(ref, mode)
        sec, Lo, Po = o.query_batch(buf[: int(off[sample_reads])], off[: sample_reads + 1], threads=cores)
        sb = int(off[sample_reads])
        if main_bits:
            got = sp.expand_bits(np_len, off[: sample_reads + 1])
        else:
            got = sp.widen_lengths(np_len[:sb], off[: sample_reads + 1]) if (mode == "pml" and main_m["clamped"]) else np_len[:sb]
        mism = int(np.count_nonzero(Lo != got.astype(np.uint64)))
        if mode == "ms":
            mism += int(np.count_nonzero(Po != np_ptr[:sb].astype(np.uint64)))
        parity = {"reads_checked": sample_reads, "bases_checked": sb, "mismatches": mism, "host_equals_device_path": main_m["same"]}
        cpu_baseline = {"value": sb / sec, "unit": "bases/s", "cores": cores, "kind": "port",
                        "sample": f"first {sample_reads} reads ({sb} bases) of the same workload, in-memory, OpenMP over reads"}
        ref_sec = reference_binary_seconds(ref, buf, off, sample_reads, mode, cores)
        if ref_sec:  # the reference's own code when it was built: that is the baseline, the port stays next to it
            cpu_baseline = {"value": sb / ref_sec, "unit": "bases/s", "cores": cores, "kind": "reference",
                            "sample": f"first {sample_reads} reads ({sb} bases) of the same workload through `spumoni run -n -t {cores}` of the "
                                      "reference's own translation units (oracle/_ref, stand-in third-party containers), its own phase timer",
                            "oracle_port": cpu_baseline["value"]}
        if mism or not main_m["same"]:
            raise SystemExit(f"bench.py: PARITY FAILURE vs oracle: {parity}")

    # ---------------- secondary measurements: other transports, -c statistics, MS (configs[2]), the CLI
    variants = {}

    def keep(k, v):
        variants[k] = {q: v[q] for q in ("value", "kernel_ms", "e2e", "h2d_bytes_per_step", "d2h_bytes_per_step", "launches", "clamped", "same") if q in v}

    if want and mode == "pml":
        # every transport must give the lengths of the headline run
        L_main = sp.expand_bits(np_len, off) if main_bits else (sp.widen_lengths(np_len, off) if main_m["clamped"] else np_len.astype(np.uint32))
        fmts = [f for f in ("u8", "u16", "u32", "bits", "packed_bits") if f != main_fmt and f in want]
        if world > 1:
            fmts = [f for f in ("u8", "bits", "packed_bits") if f != main_fmt and f in want]
        for f in fmts:
            v, Lv, _ = measure(idx, f, "pml", steps=max(2, args.steps // 2))
            Lx = sp.expand_bits(Lv, off) if f in ("bits", "packed_bits") else (sp.widen_lengths(Lv, off) if v.get("clamped") else Lv.astype(np.uint32))
            v["same"] = bool(v.get("same", True)) and bool(np.array_equal(Lx, L_main))
            if not v["same"]:
                raise SystemExit(f"bench.py: variant {f} disagrees with the headline lengths")
            keep(f, v)
        if world == 1 and "pml_stats" in want:
            v, _, _ = measure(idx, main_fmt, "pml", stats=True, steps=max(2, args.steps // 2))  # -c: per-read bin statistics ride along
            keep("pml_stats", v)
    cli = None
    if world == 1 and mode == "pml" and "synth" not in w and "cli" in want:
        idx = None  # the CLI opens its own copy
        torch.cuda.empty_cache()
        os.sched_setaffinity(0, all_cpus)
        cores = os.cpu_count() or 1
        cli_seconds(ref, name, 4, cores)  # page cache + first-touch warm-up
        cli = cli_seconds(ref, name, 4, cores)
        if cli:
            cli["bases_per_s"] = n_bases / cli["phase_s"]
        if "cli_fastq" in want and cli is not None:  # the same reads as four-line FASTQ records (twice the file bytes per base)
            try:
                cli_seconds(ref, name, 4, cores, fastq=True)
                fq = cli_seconds(ref, name, 4, cores, fastq=True)
                if fq:
                    fq["bases_per_s"] = n_bases / fq["phase_s"]
                    cli["fastq"] = fq
            except Exception as e:  # a secondary measurement never takes the headline line down
                log(f"cli_fastq skipped: {e}")
    if world == 1 and mode == "pml" and "synth" not in w and "ms" in want:
        try:
            idx = None  # close the PML index first (it holds the device's persisting-L2 carve-out, which MS does not want)
            import gc
            gc.collect()
            idx_ms_ref = ensure_index(name, "ms")
            ms_idx = sp.Index.open(idx_ms_ref, sp.MODE_MS).to_device(local)
            n_ms = ms_idx.get_bwt_stats()[0]
            v, _, _ = measure(ms_idx, "u8" if main_bits else main_fmt, "ms", steps=max(2, args.steps // 2))
            keep("ms", v)
            variants["ms"]["roofline_frac_model"] = BYTES_PER_BASE["ms"] * n_bases / (v["kernel_ms"] / 1e3) / 1e9 / peaks()[0]
            del ms_idx
        except SystemExit:
            raise
        except Exception as e:  # the MS line is secondary: report why it is missing instead of losing the headline
            variants["ms"] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        peak, peak_src = peaks()
        kavg_ms = main_m["kernel_ms"]
        achieved = BYTES_PER_BASE[mode] * n_bases / (kavg_ms / 1e3) / 1e9
        pw_run = 4 if (mode == "ms" and n < (1 << 32) - 1 and not args.pointers64) else 8
        as_run = 1 + (args.length_bytes or 0.125) + 64 + ((pw_run + 8 + 32) if mode == "ms" else 0)  # the same model with the widths this run used
        kname = ("spb_move_kernel<PML>" if mode == "pml" else "spb_move_kernel<MS> + spb_ms_lengths_kernel") if (args.lanes or 1) == 1 else "spb_walk_kernel<G>"
        out = {
            "metric": METRICS[mode], "value": main_m["value"], "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_m["ms_per_step"], "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": w["desc"].replace("PML", mode.upper()), "n": n, "r": r, "reads_per_gpu": n_reads, "bases_per_gpu": n_bases,
                       "job_bases": job_bases, "lanes_per_read": args.lanes or 1, "format": main_fmt,
                       "l2": "inputs (1 GB reads per step) and outputs (1-4 GB lengths) larger than L2; no explicit flush",
                       "parallelism": f"reads sharded over {world} GPU(s), index replicated (one NCCL broadcast at load)" +
                                      ("; strong scaling: one fixed set of 5 x the workload's reads cut by bases" if args.strong else "")},
            "clocks": clocks,
            "e2e": {"value": main_m["e2e"], "unit": "bases/s", "h2d_bytes_per_step": main_m["h2d_bytes_per_step"], "d2h_bytes_per_step": main_m["d2h_bytes_per_step"],
                    "api": ((f"spb_pml_query_host_packed, pinned host buffers, reads as 2 bits per base + exception list (packed by the caller, untimed: spb_pack_reads), "
                             if main_fmt == "packed_bits" else f"spb_{mode}_query_host, pinned host buffers, one byte per base in, ") +
                            ("one bit per base out (SPB_LENGTH_BITS: bit set iff the PML is non-zero, which determines every length; spb_pml_expand_bits)" if main_bits else
                             f"uint{8 * args.length_bytes} lengths out ({main_m['clamped']} clamped values per step" +
                             ("; clamped PML lengths are rebuilt exactly by spb_pml_widen_lengths)" if mode == "pml" else ")")))},
            "gpu_launches": int(main_m["launches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": kname, "kernel_ms": kavg_ms,
                         "bytes_per_base": BYTES_PER_BASE[mode], "bytes_per_base_as_run": as_run,
                         "note": "achieved = SURVEY section 8(d)'s algorithmic bytes per base x bases / kernel time (a model: two index sectors per step whatever the layout); "
                                 "traffic = DRAM bytes per launch measured by ncu (profiles/traffic.json), dram_frac = traffic / kernel time / peak",
                         "peak_source": peak_src},
            "cpu_baseline": cpu_baseline, "parity": parity, "variants": variants, "cli": cli,
        }
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                tj = json.load(open(tr)).get(name + ":" + mode)
                if tj and tj.get("kernel") == kname:
                    out["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
                    out["roofline"]["traffic_bases_per_launch"] = tj.get("bases_per_launch")
                    out["roofline"]["dram_frac"] = tj["dram_bytes_per_launch"] * (n_bases / tj["bases_per_launch"]) / (kavg_ms / 1e3) / 1e9 / peak
            except Exception:
                pass
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
