"""CPU: the C-ABI shared library loads and exports exactly what include/spumoni_b200.h declares."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import spumoni_b200 as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "spumoni_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(sp.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spumoni_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", sp.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (spb_[a-z0-9_]+)", out)))
    assert exported == names, "exported spb_* symbols and the header disagree"


def test_abi_version_and_struct_sizes():
    lib = sp.load_library()
    assert lib.spb_abi_version() == 4
    assert ctypes.sizeof(sp.QueryOpts) == 32
    assert sp.BinStats.itemsize == 16
    assert sp.flat_header_bytes() % 8 == 0


def test_kernels_are_sm100a_sass():
    out = subprocess.run(["cuobjdump", "-lelf", sp.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly(small_index):
    """Without a device the product refuses to answer (no CPU fallback)."""
    if sp.device_count() > 0:
        pytest.skip("GPU present")
    idx = sp.Index.open(small_index["ref"], sp.MODE_PML)
    with pytest.raises(sp.SpumoniError) as e:
        idx.to_device(0)
    assert e.value.code == 5
    buf, off = sp.pack_reads([b"ACGT"])
    with pytest.raises(sp.SpumoniError) as e:
        idx.query_host(buf, off)
    assert e.value.code == 5
    with pytest.raises(sp.SpumoniError):
        sp.pml_t(small_index["ref"])


def test_cli_validation_messages(small_index, tmp_path):
    """Exit code 1 + the reference's messages (spumoni_main.hpp:267-329)."""
    cli = sp.CLI_PATH
    prefix = small_index["ref"][:-3]
    r = subprocess.run([cli, "run", "-r", prefix, "-P", "-n"], capture_output=True, text=True)
    assert r.returncode == 1 and "Both a reference file (-r) and pattern file (-p) must be provided." in r.stderr
    r = subprocess.run([cli, "run", "-r", prefix, "-p", small_index["reads_fa"], "-n"], capture_output=True, text=True)
    assert r.returncode == 1 and "An output type with -M or -P must be specified" in r.stderr
    bad = str(tmp_path / "reads.txt")
    open(bad, "w").write(">a\nACGT\n")
    r = subprocess.run([cli, "run", "-r", prefix, "-p", bad, "-P", "-n"], capture_output=True, text=True)
    assert r.returncode == 1 and "does not appear to be a FASTA" in r.stderr
    r = subprocess.run([cli, "run", "-r", prefix, "-p", small_index["reads_fa"], "-P", "-n", "-w", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "bin size used is not optimal" in r.stderr
    r = subprocess.run([cli, "run", "-r", prefix, "-p", small_index["reads_fa"], "-P"], capture_output=True, text=True)
    assert r.returncode == 1 and "A minimizer type must be specified using -m or -a." in r.stderr
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1 and "SPUMONI version: 2.0.9" in r.stderr


def test_no_gpu_replicate_fails_loudly(small_index):
    if sp.device_count() > 0:
        pytest.skip("GPU present")
    idx = sp.Index.open(small_index["ref"], sp.MODE_PML)
    with pytest.raises(sp.SpumoniError) as e:
        idx.replicate([0, 1])
    assert e.value.code == 5 and idx.device_count() == 0
