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


@pytest.fixture(params=["8", "16", "32"])
def row_width(request, monkeypatch):
    """The move-row layouts of the flat image: compact 8-byte rows (the default when R < 2^24), the wide 16-byte rows
    (larger MS / -d indexes) and the full 32-byte rows (larger PML indexes; MS / -d images ignore the request and keep
    their default); the layout is chosen when an index is opened."""
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


def _read5(path, cols):
    import numpy as np
    b = np.frombuffer(open(path, "rb").read(), dtype=np.uint8).reshape(-1, 5 * cols)
    out = []
    for c in range(cols):
        v = np.zeros(len(b), dtype=np.uint64)
        for j in range(5):
            v |= b[:, 5 * c + j].astype(np.uint64) << np.uint64(8 * j)
        out.append(v.astype(np.int64))
    return out


def _write5(path, *cols):
    import numpy as np
    rows = np.zeros((len(cols[0]), 5 * len(cols)), dtype=np.uint8)
    for c, col in enumerate(cols):
        v = np.asarray(col, dtype=np.uint64)
        for j in range(5):
            rows[:, 5 * c + j] = ((v >> np.uint64(8 * j)) & np.uint64(0xFF)).astype(np.uint8)
    rows.tofile(path)


@pytest.fixture(scope="session")
def huge_periodic_index(tmp_path_factory):
    """A PML index whose BWT positions exceed 2^32 (n = 48 * 90 000 000 + 1 = 4 320 000 001) without suffix-sorting
    4 G characters: for the periodic text U^k every per-run quantity of the RLBWT (run lengths, thresholds, SA samples)
    is affine in k, so the PFP intermediates (.bwt.heads/.bwt.len/.thr_pos/.ssa/.esa) of k = 6 and k = 7 are
    extrapolated (and checked against a real k = 8 build) and serialised with `mkindex from-parts`, the stand-in for
    build_spumoni_main (compute_ms_pml.cpp:1384-1408).  Exercises the 40-bit fields of every table."""
    import random
    import numpy as np
    d = str(tmp_path_factory.mktemp("huge"))
    rng = random.Random(3)
    unit = "".join(rng.choice("ACGT") for _ in range(48))

    def parts(k):
        dk = os.path.join(d, f"k{k}")
        build_index(dk, text=(unit * k).encode(), keep=True)
        ref = os.path.join(dk, "idx.fa")
        heads = open(ref + ".bwt.heads", "rb").read()
        (lens,), (thr,) = _read5(ref + ".bwt.len", 1), _read5(ref + ".thr_pos", 1)
        _, ssa = _read5(ref + ".ssa", 2)
        _, esa = _read5(ref + ".esa", 2)
        return heads, lens, thr, ssa, esa

    p6, p7, p8 = parts(6), parts(7), parts(8)
    assert p6[0] == p7[0] == p8[0]
    K = 90_000_000
    big = [p6[0]]
    for i in range(1, 5):
        slope = p7[i] - p6[i]
        assert np.array_equal(p6[i] + 2 * slope, p8[i]), "the parts are not affine in k"
        big.append(p6[i] + (K - 6) * slope)
    ref = os.path.join(d, "idx.fa")
    open(ref, "w").write(">periodic\n")  # placeholder: PML never reads the text
    open(ref + ".bwt.heads", "wb").write(big[0])
    _write5(ref + ".bwt.len", big[1])
    _write5(ref + ".thr_pos", big[2])
    starts = np.concatenate([[0], np.cumsum(big[1])[:-1]])
    _write5(ref + ".ssa", starts, big[3])
    _write5(ref + ".esa", starts + big[1] - 1, big[4])
    out = mkindex("from-parts", "--prefix", os.path.join(d, "idx"), "--pml")
    n = 48 * K + 1
    assert f'"n": {n}' in out and n > (1 << 32)
    reads = []
    text3 = unit * 3
    for _ in range(120):
        ln = rng.randint(20, 140)
        a = rng.randint(0, len(text3) - ln)
        r = bytearray(text3[a:a + ln].encode())
        for i in range(len(r)):
            if rng.random() < 0.08:
                r[i] = rng.choice(b"ACGTN")
        reads.append(bytes(r))
    reads += [unit.encode() * 2, b"N" * 5, bytes(rng.choice(b"ACGT") for _ in range(300))]
    return {"ref": ref, "n": n, "reads": reads}
