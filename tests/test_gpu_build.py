"""GPU: index construction on the device (spb_build_parts_device, spumoni_b200/csrc/build_kernels.cu) against the
suffix-array construction of tools/mkindex -- the stand-in for Big-BWT + pfp_thresholds whose intermediates the reference's
own constructors turn into byte-identical index files (tests/test_reference_pin.py)."""
import filecmp
import os
import random

import numpy as np
import pytest

import spumoni_b200 as sp
from conftest import MKINDEX, mkindex

pytestmark = pytest.mark.gpu


def _get5(path, cols=1):
    b = np.fromfile(path, dtype=np.uint8).reshape(-1, 5)
    v = np.zeros(len(b), dtype=np.uint64)
    for i in range(5):
        v |= b[:, i].astype(np.uint64) << np.uint64(8 * i)
    return v.reshape(-1, cols)


def _text_of(ref):
    return np.frombuffer("".join(l.strip() for l in open(ref) if not l.startswith(">")).encode() + b"\0", dtype=np.uint8)


@pytest.mark.parametrize("shape", ["repeats", "binary", "random", "pangenome", "unary"])
def test_device_build_equals_mkindex_intermediates(tmp_path, shape):
    rng = random.Random(3)
    d = str(tmp_path)
    if shape == "pangenome":
        mkindex("genome", "--kind", "pangenome", "--len", 300000, "--haps", 6, "--snp", 0.002, "--indel", 0.0002, "--seed", 5, "--out", d + "/in.fa")
    else:
        if shape == "repeats":
            unit = "ACGTTGCAAGGCTTAACCGGT"
            text = "".join(unit if rng.random() < 0.9 else rng.choice("ACGT") for _ in range(400)) + "A" * 700 + "".join(rng.choice("ACGTN") for _ in range(3000))
        elif shape == "binary":
            text = "".join(rng.choice("AC") for _ in range(20000))
        elif shape == "unary":
            text = "A" * 5000
        else:
            text = "".join(rng.choice("ACGT") for _ in range(100000))
        open(d + "/in.fa", "w").write(">s\n" + text + "\n")
    mkindex("build", "--input", d + "/in.fa", "--prefix", d + "/cpu", "--pml", "--ms", "--keep", *([] if shape == "pangenome" else ["--no-revcomp"]))
    T = _text_of(d + "/cpu.fa")
    P = sp.build_parts_device(T, 0, write_to=d + "/gpu.fa")
    assert P["n"] == T.size and P["r"] == os.path.getsize(d + "/cpu.fa.bwt.heads")
    assert np.array_equal(P["heads"], np.fromfile(d + "/cpu.fa.bwt.heads", dtype=np.uint8))
    assert np.array_equal(P["lens"], _get5(d + "/cpu.fa.bwt.len")[:, 0])
    assert np.array_equal(P["ssa"], _get5(d + "/cpu.fa.ssa", 2)[:, 1]) and np.array_equal(P["esa"], _get5(d + "/cpu.fa.esa", 2)[:, 1])
    assert np.array_equal(P["thr"], _get5(d + "/cpu.fa.thr_pos")[:, 0])
    for ext in (".bwt.heads", ".bwt.len", ".thr_pos", ".ssa", ".esa"):
        assert filecmp.cmp(d + "/gpu.fa" + ext, d + "/cpu.fa" + ext, shallow=False), ext
    # the intermediates become the index through the reference's build side (mkindex from-parts = its constructors + serialize)
    os.link(d + "/cpu.fa", d + "/gpu.fa") if not os.path.exists(d + "/gpu.fa") else None
    mkindex("from-parts", "--prefix", d + "/gpu", "--pml", "--ms")
    for ext in (".thrbv.spumoni", ".thrbv.ms"):
        assert filecmp.cmp(d + "/gpu.fa" + ext, d + "/cpu.fa" + ext, shallow=False), ext


def test_device_build_rejects_bad_texts():
    with pytest.raises(sp.SpumoniError):
        sp.build_parts_device(np.frombuffer(b"ACGT", dtype=np.uint8))           # no terminator
    with pytest.raises(sp.SpumoniError):
        sp.build_parts_device(np.frombuffer(b"AC\x01GT\x00", dtype=np.uint8))   # a byte <= 1 inside the text
