"""CPU: the algorithm of the GPU index builder (spumoni_b200/csrc/build_kernels.cu), restated step for step in numpy
(tests/build_emulator.py: prefix doubling over (rank, rank + h) keys, PLCP in independent text chunks, leftmost range minima
by 256-element blocks + a sparse table), against the suffix-array construction of tools/mkindex -- whose intermediates the
reference's own constructors turn into the index files (tests/test_reference_pin.py).  The CUDA kernels themselves are
compared with the same intermediates on the GPU (tests/test_gpu_build.py)."""
import os
import random
import subprocess

import numpy as np
import pytest

from build_emulator import LeftmostRmq, build_parts
from conftest import MKINDEX


def _get5(path, cols=1):
    b = np.fromfile(path, dtype=np.uint8).reshape(-1, 5)
    v = np.zeros(len(b), dtype=np.int64)
    for i in range(5):
        v |= b[:, i].astype(np.int64) << (8 * i)
    return v.reshape(-1, cols)


@pytest.mark.parametrize("case", range(6))
def test_emulated_builder_equals_mkindex(tmp_path, case):
    rng = random.Random(case + 1)
    d = str(tmp_path)
    if case < 3:
        unit = "ACGTTGCAAGGCTTAACCGGT"
        text = "".join(unit if rng.random() < 0.9 else rng.choice("ACGT") for _ in range(60 + 40 * case)) + "A" * 50 + "".join(rng.choice("ACGTN") for _ in range(200))
    else:
        text = "".join(rng.choice("AC" if case == 3 else "ACGT") for _ in range(700 * case))
    open(d + "/in.fa", "w").write(">s\n" + text + "\n")
    subprocess.check_call([MKINDEX, "build", "--input", d + "/in.fa", "--prefix", d + "/idx", "--pml", "--keep"] + (["--no-revcomp"] if case % 2 else []),
                          stdout=subprocess.DEVNULL)
    T = np.frombuffer("".join(l.strip() for l in open(d + "/idx.fa") if not l.startswith(">")).encode() + b"\0", dtype=np.uint8)
    P = build_parts(T)
    assert np.array_equal(np.fromfile(d + "/idx.fa.bwt.heads", dtype=np.uint8), P["heads"])
    assert np.array_equal(_get5(d + "/idx.fa.bwt.len")[:, 0], P["lens"])
    assert np.array_equal(_get5(d + "/idx.fa.thr_pos")[:, 0], P["thr"])
    assert np.array_equal(_get5(d + "/idx.fa.ssa", 2)[:, 1], P["ssa"]) and np.array_equal(_get5(d + "/idx.fa.esa", 2)[:, 1], P["esa"])


def test_leftmost_rmq_against_brute_force():
    rng = np.random.default_rng(3)
    v = rng.integers(0, 6, size=2000)  # many ties: the leftmost minimum must win
    rmq = LeftmostRmq(v)
    for _ in range(3000):
        l = int(rng.integers(0, len(v)))
        r = int(rng.integers(l, len(v)))
        assert rmq.query(l, r) == l + int(np.argmin(v[l:r + 1]))
