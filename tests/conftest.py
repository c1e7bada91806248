import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MKINDEX = os.path.join(ROOT, "tools", "bin", "spumoni_mkindex")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    if not os.path.exists(MKINDEX):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    import oracle
    oracle.build()
    import spumoni_b200
    if not (os.path.exists(spumoni_b200.LIB_PATH) and os.path.exists(spumoni_b200.CLI_PATH)):
        spumoni_b200.build_library()


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


@pytest.fixture(params=["8", "16"])
def row_width(request, monkeypatch):
    """Both move-row layouts of the flat image: compact 8-byte rows (the default when R < 2^24) and the wide
    16-byte rows (larger indexes); the layout is chosen when an index is opened."""
    monkeypatch.setenv("SPUMONI_B200_ROWS", request.param)
    return request.param


def mkindex(*args):
    out = subprocess.run([MKINDEX] + [str(a) for a in args], check=True, capture_output=True, text=True)
    return out.stdout


def build_index(dirpath, text: bytes = None, fasta: str = None, revcomp=False, keep=False, docs=0):
    """Builds <dir>/idx.fa(.thrbv.spumoni|.thrbv.ms) from raw text or a FASTA file; returns ref_file path."""
    os.makedirs(dirpath, exist_ok=True)
    if fasta is None:
        fasta = os.path.join(dirpath, "in.fa")
        with open(fasta, "wb") as f:
            f.write(b">s\n" + text + b"\n")
    args = ["build", "--input", fasta, "--prefix", os.path.join(dirpath, "idx"), "--pml", "--ms"]
    if not revcomp:
        args.append("--no-revcomp")
    if keep:
        args.append("--keep")
    if docs:
        args += ["--docs", docs]  # also writes <prefix>.fa.fdi and <prefix>.fa.doc (spumoni build -d)
    mkindex(*args)
    return os.path.join(dirpath, "idx.fa")


@pytest.fixture(scope="session")
def small_index(tmp_path_factory):
    """20 kb pangenome-like reference (4 haplotypes, +revcomp, 3 documents), PML + MS index + document array, 300 mixed reads."""
    d = str(tmp_path_factory.mktemp("small"))
    mkindex("genome", "--kind", "pangenome", "--len", 20000, "--haps", 4, "--snp", 0.01, "--indel", 0.001,
            "--seed", 7, "--out", os.path.join(d, "raw.fa"))
    ref = build_index(d, fasta=os.path.join(d, "raw.fa"), revcomp=True, docs=3)
    reads = os.path.join(d, "reads.fa")
    mkindex("reads", "--ref", ref, "--n", 300, "--len", 40, "--len-max", 400, "--sub", 0.05, "--ins", 0.01, "--del", 0.01,
            "--neg-frac", 0.3, "--seed", 9, "--out", reads, "--bin", os.path.join(d, "reads.bin"))
    return {"dir": d, "ref": ref, "reads_fa": reads, "reads_bin": os.path.join(d, "reads.bin")}


def load_bin_reads(path):
    import numpy as np
    return np.fromfile(path, dtype=np.uint8), np.fromfile(path + ".off", dtype=np.uint64)
