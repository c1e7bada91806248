"""Synthetic run-length BWTs (tools/mkindex `synth`): indexes of any size without a suffix array, serialised like real
ones. They are not the BWT of a text -- the query path is rank/select/threshold arithmetic and does not care -- so the
reference's own code, the oracle, the CPU walk of the flat image and the CUDA path must all still agree. This is the
stand-in for the n > 2^32, r ~ 10^8 configurations (BASELINE configs 4-5). (File name: sorts last, so the round-end
`pytest -x` reaches it after every other GPU test.)"""
import filecmp
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
import spumoni_b200 as sp
from conftest import ROOT, load_bin_reads, mkindex
from flat_emulator import FlatImage

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "spumoni_ref")
SHAPES = {
    # pangenome-like: short runs, a few long ones; 8 % of the steps jump
    "typical": ["--runs", 30000, "--mean-len", 20, "--seed", 3, "--n", 300, "--len", 400, "--sub", 0.08],
    # jump-heavy like the c2 reads (58 % jumps)
    "jumpy": ["--runs", 30000, "--mean-len", 6, "--seed", 4, "--n", 300, "--len", 400, "--sub", 0.55],
    # n > 2^32 (40-bit positions in every generic table) with mostly short runs: 3 % of the runs are ~5 * 10^7 long
    "huge": ["--runs", 6000, "--mean-len", 25, "--tail-frac", 0.03, "--tail-mul", 2000000, "--seed", 5, "--n", 200, "--len", 300,
             "--sub", 0.1],
    # runs longer than 2^32: no move rows at all, the generic kernels serve every step
    "giant": ["--runs", 3000, "--mean-len", 25, "--tail-frac", 0.005, "--tail-mul", 400000000, "--seed", 6, "--n", 100, "--len", 200,
              "--sub", 0.1],
}
# corner shapes, CPU comparisons only: 8 runs, every run of length 1, long runs only, 9 runs of ~10^5, a third of the runs long
CORNERS = {
    "tiny8": ["--runs", 8, "--mean-len", 3, "--seed", 9, "--n", 60, "--len", 120, "--sub", 0.2],
    "len1": ["--runs", 5000, "--mean-len", 1, "--seed", 9, "--n", 60, "--len", 120, "--sub", 0.2],
    "long": ["--runs", 2000, "--mean-len", 1000, "--tail-frac", 0, "--seed", 9, "--n", 60, "--len", 120, "--sub", 0.2],
    "tiny_long": ["--runs", 9, "--mean-len", 100000, "--seed", 9, "--n", 60, "--len", 120, "--sub", 0.2],
    "mixed": ["--runs", 70000, "--mean-len", 2, "--tail-frac", 0.3, "--tail-mul", 300, "--seed", 9, "--n", 60, "--len", 120, "--sub", 0.2],
}


def _make(shape, args, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("synth_" + shape))
    out = mkindex("synth", "--prefix", os.path.join(d, "idx"), "--bin", os.path.join(d, "r.bin"), "--out", os.path.join(d, "reads.fa"), *args)
    n = int(out.split('"n":')[1].split(",")[0])
    assert (n > 2 ** 32) == (shape in ("huge", "giant"))
    shutil.copy(os.path.join(ROOT, "tests", "golden", "reference_run", "idx.fa.pmlnulldb"), os.path.join(d, "idx.fa.pmlnulldb"))
    return {"dir": d, "ref": os.path.join(d, "idx.fa"), "reads_fa": os.path.join(d, "reads.fa"), "reads_bin": os.path.join(d, "r.bin"),
            "shape": shape}


@pytest.fixture(scope="module", params=sorted(SHAPES))
def synth(request, tmp_path_factory):
    return _make(request.param, SHAPES[request.param], tmp_path_factory)


@pytest.fixture(scope="module", params=sorted(CORNERS))
def corner(request, tmp_path_factory):
    return _make(request.param, CORNERS[request.param], tmp_path_factory)


def test_synthetic_oracle_equals_reference_binary(synth, tmp_path):
    _oracle_equals_reference_binary(synth, tmp_path)


def test_corner_shapes_oracle_equals_reference_binary(corner, tmp_path):
    _oracle_equals_reference_binary(corner, tmp_path)


def test_corner_shapes_flat_image_equals_oracle(corner, row_width):
    test_synthetic_flat_image_equals_oracle(corner, row_width)


def _oracle_equals_reference_binary(synth, tmp_path):
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "-s"])
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/spumoni_ref is not built (needs /root/reference)")
    reads = str(tmp_path / "reads.fa")
    shutil.copy(synth["reads_fa"], reads)
    subprocess.run([REF_BIN, "run", "-r", synth["ref"][:-3], "-p", reads, "-P", "-n", "-c"], check=True, capture_output=True)
    want = str(tmp_path / "want")
    os.makedirs(want)
    for e in (".pseudo_lengths", ".report"):
        shutil.move(reads + e, os.path.join(want, "reads.fa" + e))
    oracle.run(synth["ref"], reads, "pml", report=True)
    for e in (".pseudo_lengths", ".report"):
        assert filecmp.cmp(reads + e, os.path.join(want, "reads.fa" + e), shallow=False), e


def test_synthetic_flat_image_equals_oracle(synth, row_width):
    buf, off = load_bin_reads(synth["reads_bin"])
    o = oracle.OracleIndex(synth["ref"], "pml")
    fi = FlatImage(sp.Index.open(synth["ref"], sp.MODE_PML).flat())
    # runs longer than 2^32 are cut into rows by balance_runs (index_host.cpp), so even "giant" gets move rows ...
    assert fi.num_fast > 0 and fi.R >= fi.R_file
    assert bool(fi.off_mrows8) == (row_width == "8") and bool(fi.off_mrows32) == (row_width == "32")
    if synth["shape"] == "giant":  # ... and without the cutting it has none: the generic kernels serve every step
        os.environ["SPUMONI_B200_BALANCE"] = "0"
        try:
            raw = FlatImage(sp.Index.open(synth["ref"], sp.MODE_PML).flat())
        finally:
            del os.environ["SPUMONI_B200_BALANCE"]
        assert raw.num_fast == 0 and raw.R == raw.R_file
        for i in range(0, 10):
            r = bytes(buf[int(off[i]):int(off[i + 1])])
            assert raw.walk(r)[0] == o.matching_statistics(r).tolist(), i
    matched = 0
    for i in range(0, min(40, len(off) - 1)):
        r = bytes(buf[int(off[i]):int(off[i + 1])])
        want = o.matching_statistics(r).tolist()
        assert fi.walk(r)[0] == want, i
        if fi.num_fast:
            assert fi.walk_rows(r)[0] == want, i
        matched += sum(1 for v in want if v > 0)
    assert matched > 1000 or synth["shape"] not in ("typical", "huge", "giant")  # the generated reads do match in stretches


@pytest.mark.gpu
def test_synthetic_cuda_equals_oracle(synth, row_width):
    buf, off = load_bin_reads(synth["reads_bin"])
    _, Lo, _ = oracle.OracleIndex(synth["ref"], "pml").query_batch(buf, off, threads=os.cpu_count() or 1)
    idx = sp.pml_t(synth["ref"], device=0)
    L, _, S = idx.query_host(buf, off, bin_width=150, max_value_thr=7)
    assert np.array_equal(L.astype(np.uint64), Lo)
    for lanes in (8, 32):  # the cooperative generic kernel on the same image
        Lg, _, _ = idx.query_host(buf, off, lanes=lanes)
        assert np.array_equal(Lg, L), lanes
    # bin-max classifier against the literal restatement
    from oracle import naive
    for r in range(0, len(off) - 1, 37):
        _, summ, above, below, _ = naive.classify_bins(Lo[int(off[r]):int(off[r + 1])].tolist(), 150, 7)
        assert (int(S[r]["sum_max"]), int(S[r]["above"]), int(S[r]["below"])) == (summ, above, below)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,flags,exts", [("Pq", ["-P", "-c"], (".pseudo_lengths", ".report")), ("Mq", ["-M"], (".lengths", ".pointers"))])
def test_zz_cli_fastq_records_equal_reference(tmp_path, mode, flags, exts):
    """FASTQ records through the product CLI against the reference's own outputs (tests/golden/reference_run): BatchLoader
    accepts '@' records, and loses the last short batch of a newline-terminated file (21 of 23 records survive). Last test of
    the suite on purpose: it was written after the last GPU minutes of the round were spent (the parser itself is covered on
    the CPU by tests/test_cli_parser.py)."""
    golden = os.path.join(ROOT, "tests", "golden", "reference_run")
    d = str(tmp_path / "g")
    shutil.copytree(golden, d, ignore=shutil.ignore_patterns("expected"))
    r = subprocess.run([sp.CLI_PATH, "run", "-r", os.path.join(d, "idx"), "-p", os.path.join(d, "fastq_reads.fa"), "-n"] + flags,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for e in exts:
        assert open(os.path.join(d, "fastq_reads.fa" + e), "rb").read() == open(os.path.join(golden, "expected", mode, "fastq_reads.fa" + e), "rb").read(), e
    assert "finished processing 21 reads" in r.stderr
