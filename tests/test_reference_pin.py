"""Pins the CPU oracle to the reference's OWN code.

oracle/_ref/spumoni_ref = /root/reference's translation units compiled where they lie against stand-ins for its
un-vendored third-party headers (oracle/shim/, oracle/Makefile.ref). tests/golden/reference_run/ holds what that binary
wrote for a small index (tests/golden/make_reference_golden.py); those files are committed, so the first group of tests
runs anywhere. The second group needs the binary (present where /root/reference is, or shipped prebuilt in oracle/_ref/)
and compares larger, uncommitted fixtures.
"""
import ctypes
import filecmp
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from conftest import MKINDEX, ROOT, build_index, mkindex

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_run")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "spumoni_ref")
MODES = {"P": ("pml", True, False, ["-P", "-n", "-c"]), "Pd": ("pml", False, True, ["-P", "-n", "-d"]),
         "M": ("ms", True, False, ["-M", "-n", "-c"]), "Md": ("ms", True, True, ["-M", "-n", "-d", "-c"])}
# FASTQ records (BatchLoader accepts them, batch_loader.cpp:33-34; the last short batch of a newline-terminated file is lost, :51)
FASTQ = {"Pq": ("pml", True, ["-P", "-n", "-c"]), "Mq": ("ms", False, ["-M", "-n"])}


def _outputs(d, stem):
    return sorted(f for f in os.listdir(d) if f.startswith(stem + "."))


def _assert_same_outputs(got_dir, want_dir, stem):
    want = _outputs(want_dir, stem)
    assert want, "no expected outputs in " + want_dir
    for f in want:
        assert os.path.exists(os.path.join(got_dir, f)), f + " not written"
        assert filecmp.cmp(os.path.join(got_dir, f), os.path.join(want_dir, f), shallow=False), f + " differs from the reference's"


@pytest.fixture()
def golden_copy(tmp_path):
    d = str(tmp_path / "g")
    shutil.copytree(GOLDEN, d, ignore=shutil.ignore_patterns("expected"))
    return d


GENERAL = {"Pg": ("pml", ["-P", "-n", "-g"]), "Mg": ("ms", ["-M", "-n", "-g"])}


def _oracle_general_text(ref, pattern, kind):
    """classify_general_reads_* (compute_ms_pml.cpp:1219-1297) over the oracle's per-read query: reads end at 0x01, are used
    verbatim, bytes after the last separator are dropped; every value is followed by a blank."""
    o = oracle.OracleIndex(ref, kind)
    reads = open(pattern, "rb").read().split(b"\x01")[:-1]
    fmt = lambda v: b"".join(b"%d " % int(x) for x in v) + b"\n"
    if kind == "pml":
        body = b"".join(b">read_%d\n" % i + fmt(o.matching_statistics(r)) for i, r in enumerate(reads))
        open(pattern + ".pseudo_lengths", "wb").write(body)
    else:
        res = [o.matching_statistics(r) for r in reads]
        open(pattern + ".lengths", "wb").write(b"".join(b">read_%d\n" % i + fmt(L) for i, (L, P) in enumerate(res)))
        open(pattern + ".pointers", "wb").write(b"".join(b">read_%d\n" % i + fmt(P) for i, (L, P) in enumerate(res)))


@pytest.mark.parametrize("mode", sorted(GENERAL))
def test_oracle_equals_committed_reference_outputs_general_text(golden_copy, mode):
    kind, _ = GENERAL[mode]
    _oracle_general_text(os.path.join(golden_copy, "idx.fa"), os.path.join(golden_copy, "general.txt"), kind)
    _assert_same_outputs(golden_copy, os.path.join(GOLDEN, "expected", mode), "general.txt")


@pytest.mark.parametrize("mode", sorted(FASTQ))
def test_oracle_equals_committed_reference_outputs_fastq(golden_copy, mode):
    kind, report, _ = FASTQ[mode]
    oracle.run(os.path.join(golden_copy, "idx.fa"), os.path.join(golden_copy, "fastq_reads.fa"), kind, report=report)
    _assert_same_outputs(golden_copy, os.path.join(GOLDEN, "expected", mode), "fastq_reads.fa")
    kept = sum(1 for line in open(os.path.join(golden_copy, "fastq_reads.fa." + ("pseudo_lengths" if kind == "pml" else "lengths"))) if line[0] == ">")
    assert kept == 21  # of 23 records: the quirk is part of the contract


@pytest.mark.parametrize("mode", sorted(MODES))
def test_oracle_equals_committed_reference_outputs(golden_copy, mode):
    """oracle `spumoni run` == the files the reference's code wrote (PML/MS x report x document numbers; reads with lower
    case, N runs, characters absent from the index, one-base reads, the read-id quirk of multi-word headers)."""
    kind, report, use_doc, _ = MODES[mode]
    oracle.run(os.path.join(golden_copy, "idx.fa"), os.path.join(golden_copy, "reads.fa"), kind, report=report, use_doc=use_doc)
    _assert_same_outputs(golden_copy, os.path.join(GOLDEN, "expected", mode), "reads.fa")


def test_product_parser_accepts_reference_written_index():
    """The index files in the golden directory were serialised by the reference's build code: the product's host parser
    (no GPU needed) must accept them and agree on n and r."""
    import spumoni_b200 as sp
    for mode in (sp.MODE_PML, sp.MODE_MS):
        idx = sp.Index.open(os.path.join(GOLDEN, "idx.fa"), mode, use_doc=True)
        assert idx.get_bwt_stats() == (2403, 769)
        assert idx.num_docs() == 2
        idx.close()


# ---------------------------------------------------------------------------------------------------------------
# below: needs the binary


@pytest.fixture(scope="session")
def ref_bin():
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "-s"])
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/spumoni_ref is not built (needs /root/reference)")
    return REF_BIN


def _write_slp(ref):
    with open(ref + ".slp", "wb") as f:  # the stand-in for the SLP random-access structure reads the plain text
        f.write(b"".join(line.rstrip(b"\r\n") for line in open(ref, "rb") if not line.startswith(b">")))


def _run_ref(ref_bin, ref, reads, flags):
    subprocess.run([ref_bin, "run", "-r", ref[:-3], "-p", reads] + flags, check=True, capture_output=True)


def test_reference_binary_reproduces_committed_outputs(ref_bin, golden_copy):
    _write_slp(os.path.join(golden_copy, "idx.fa"))
    runs = [(mode, "reads.fa", flags) for mode, (_, _, _, flags) in MODES.items()]
    runs += [(mode, "general.txt", flags) for mode, (_, flags) in GENERAL.items()]
    runs += [(mode, "fastq_reads.fa", flags) for mode, (_, _, flags) in FASTQ.items()]
    for mode, pattern, flags in runs:
        _run_ref(ref_bin, os.path.join(golden_copy, "idx.fa"), os.path.join(golden_copy, pattern), flags)
        _assert_same_outputs(golden_copy, os.path.join(GOLDEN, "expected", mode), pattern)
        for f in _outputs(golden_copy, pattern):
            os.remove(os.path.join(golden_copy, f))


@pytest.fixture(scope="session")
def pinned_index(ref_bin, tmp_path_factory):
    """40 kb pangenome (+revcomp, 3 documents) whose index files, document array and null databases are all written by
    the reference's code from mkindex's PFP-style intermediates; 400 mixed reads."""
    d = str(tmp_path_factory.mktemp("pinned"))
    mkindex("genome", "--kind", "pangenome", "--len", 20000, "--haps", 4, "--snp", 0.01, "--indel", 0.001, "--seed", 23,
            "--out", os.path.join(d, "raw.fa"))
    ref = build_index(d, fasta=os.path.join(d, "raw.fa"), revcomp=True, keep=True, docs=3)
    mk = {}
    for ext in (".thrbv.spumoni", ".thrbv.ms", ".doc"):
        mk[ext] = open(ref + ext, "rb").read()
        os.remove(ref + ext)
    out = subprocess.check_output([ref_bin, "x-build-pml", ref], text=True)
    runs = int(out.split()[3])
    subprocess.check_call([ref_bin, "x-build-ms", ref], stdout=subprocess.DEVNULL)
    subprocess.check_call([ref_bin, "x-doc", ref, str(runs)])
    _write_slp(ref)
    reads = os.path.join(d, "reads.fa")
    mkindex("reads", "--ref", ref, "--n", 400, "--len", 30, "--len-max", 500, "--sub", 0.05, "--ins", 0.01, "--del", 0.01,
            "--neg-frac", 0.3, "--seed", 5, "--out", reads, "--bin", os.path.join(d, "reads.bin"))
    null_reads = os.path.join(d, "null.fa")
    mkindex("reads", "--ref", ref, "--n", 60, "--len", 150, "--len-max", 150, "--sub", 0.0, "--ins", 0.0, "--del", 0.0,
            "--neg-frac", 0.0, "--seed", 6, "--out", null_reads, "--bin", os.path.join(d, "null.bin"))
    db = {}
    for kind in ("pml", "ms"):
        db[kind] = subprocess.check_output([ref_bin, "x-nulldb", ref, null_reads, kind, "150"], text=True,
                                           stderr=subprocess.DEVNULL).split()
    return {"dir": d, "ref": ref, "reads_fa": reads, "null_bin": os.path.join(d, "null.bin"), "mkindex_bytes": mk, "db": db,
            "runs": runs}


def test_reference_serialisation_equals_fixture_builder(pinned_index):
    """pml_pointers / ms_pointers constructors + serialize (compute_ms_pml.cpp:43-80, :192-213, :355-400, :517-542), thr_bv
    (thresholds_ds.hpp:384-440) and DocumentArray (doc_array.cpp:18-108) run by the reference write the very bytes
    tools/mkindex writes: the fixture builder's reading of the reference's build side is right."""
    for ext, data in pinned_index["mkindex_bytes"].items():
        assert open(pinned_index["ref"] + ext, "rb").read() == data, ext


@pytest.mark.parametrize("mode", sorted(MODES))
def test_oracle_equals_reference_binary(ref_bin, pinned_index, tmp_path, mode):
    kind, report, use_doc, flags = MODES[mode]
    reads = str(tmp_path / "reads.fa")
    shutil.copy(pinned_index["reads_fa"], reads)
    _run_ref(ref_bin, pinned_index["ref"], reads, flags)
    want = str(tmp_path / "want")
    os.makedirs(want)
    for f in _outputs(str(tmp_path), "reads.fa"):
        shutil.move(str(tmp_path / f), os.path.join(want, f))
    oracle.run(pinned_index["ref"], reads, kind, report=report, use_doc=use_doc)
    _assert_same_outputs(str(tmp_path), want, "reads.fa")


@pytest.mark.parametrize("kind", ["pml", "ms"])
def test_oracle_null_database_equals_reference(pinned_index, kind):
    """EmpNullDatabase as the reference builds it (reversed, upper-cased null reads -> statistics -> mean, percentile,
    packed vector) against the oracle's generator on the same reads; ks_stat_threshold is rand()-driven in the
    reference and caller-supplied here, so it is the one field left out."""
    from conftest import load_bin_reads
    buf, off = load_bin_reads(pinned_index["null_bin"])
    o = oracle.OracleIndex(pinned_index["ref"], kind)
    stats, mean, perc = o.nulldb_generate(buf, off)
    rep = pinned_index["db"][kind]  # num_values N ks_stat_threshold T mean M percentile P
    assert int(rep[1]) == len(stats)
    assert float(rep[5]) == mean and float(rep[7]) == perc
    raw = open(pinned_index["ref"] + (".pmlnulldb" if kind == "pml" else ".msnulldb"), "rb").read()
    assert np.frombuffer(raw[:8], dtype="<u8")[0] == len(stats)
    assert np.frombuffer(raw[16:32], dtype="<f8").tolist() == [mean, perc]
    bits = int(np.frombuffer(raw[32:40], dtype="<u8")[0])
    width = raw[40]
    assert bits == len(stats) * width
    words = np.frombuffer(raw[41:], dtype="<u8")
    got = [int((int(words[(i * width) >> 6]) >> ((i * width) & 63) | (int(words[((i * width) >> 6) + 1]) << (64 - ((i * width) & 63))
                if ((i * width) & 63) + width > 64 else 0)) & ((1 << width) - 1)) for i in range(len(stats))]
    assert got == [int(s) for s in stats]
    # the product's writer (spb_nulldb_save) given the same statistics and scalars: the reference's file, byte for byte
    import spumoni_b200 as sp
    db = sp.EmpNullDatabase().load(pinned_index["ref"] + (".pmlnulldb" if kind == "pml" else ".msnulldb"))
    assert (db.num_values, db.mean_null_stat, db.percentile_value) == (len(stats), mean, perc)
    out = pinned_index["ref"] + ".product_" + kind + "nulldb"
    c = sp._NullDb(len(stats), db.ks_stat_threshold, db.mean_null_stat, db.percentile_value)
    s32 = np.ascontiguousarray(stats, dtype=np.uint32)
    assert sp.load_library().spb_nulldb_save(out.encode(), ctypes.byref(c), s32.ctypes.data) == 0
    assert open(out, "rb").read() == raw


@pytest.mark.parametrize("seed,alphabet", [(1, b"ACGT\xe9N"), (2, b"AC"), (3, b"A"), (4, b"ACGTN"), (5, b"\x02\x7f\x80\xffAC"), (6, b"acgtACGT")])
def test_oracle_equals_reference_binary_on_odd_alphabets(ref_bin, tmp_path, seed, alphabet):
    """Random texts over small / non-DNA alphabets (bytes >= 0x80 exercise the reference's signed-char comparisons, App. C-1;
    lower-case text vs upper-cased reads), reads with characters that never occur in the index."""
    import random
    rng = random.Random(seed)
    d = str(tmp_path)
    text = bytes(rng.choice(alphabet) for _ in range(rng.randint(50, 2500)))
    ref = build_index(d, text=text)
    _write_slp(ref)
    for ext in (".pmlnulldb", ".msnulldb"):
        shutil.copy(os.path.join(GOLDEN, "idx.fa.pmlnulldb"), ref + ext)
    pool = bytes(set(alphabet + b"ACGTNY\xe9\xfc"))
    reads = os.path.join(d, "reads.fa")
    with open(reads, "wb") as f:
        for i in range(60):
            ln = rng.randint(1, 90)
            if rng.random() < 0.6 and len(text) > ln:
                s = rng.randint(0, len(text) - ln)
                r = bytearray(text[s:s + ln])
                for k in range(len(r)):
                    if rng.random() < 0.1:
                        r[k] = rng.choice(pool)
            else:
                r = bytearray(rng.choice(pool) for _ in range(ln))
            f.write(b">r%d\n" % i + bytes(r) + b"\n")
    want = os.path.join(d, "want")
    os.makedirs(want)
    for flags in (["-P", "-n", "-c"], ["-M", "-n", "-c"]):
        _run_ref(ref_bin, ref, reads, flags)
        for f in _outputs(d, "reads.fa"):
            shutil.move(os.path.join(d, f), os.path.join(want, f if not f.endswith(".report") else flags[0] + f))
    for kind, flag in (("pml", "-P"), ("ms", "-M")):
        oracle.run(ref, reads, kind, report=True)
        os.replace(reads + ".report", os.path.join(d, flag + "reads.fa.report"))
    for f in sorted(os.listdir(want)):
        assert filecmp.cmp(os.path.join(d, f), os.path.join(want, f), shallow=False), f


def test_oracle_equals_reference_binary_multithreaded(ref_bin, pinned_index, tmp_path):
    """-t 4 on both sides: record order is not deterministic (compute_ms_pml.cpp:1001), the set of records is."""
    def records(path):
        rec, cur = {}, None
        for line in open(path, "rb"):
            if line.startswith(b">"):
                cur = line
            else:
                rec[cur] = line
        return rec
    reads = str(tmp_path / "reads.fa")
    shutil.copy(pinned_index["reads_fa"], reads)
    _run_ref(ref_bin, pinned_index["ref"], reads, ["-M", "-n", "-t", "4"])
    want = {e: records(reads + e) for e in (".lengths", ".pointers")}
    oracle.run(pinned_index["ref"], reads, "ms", threads=4)
    for e in want:
        assert records(reads + e) == want[e] and len(want[e]) == 400


def test_flat_image_reproduces_reference_outputs(row_width):
    """The product's host side (parse of the reference-written index files -> flat device image), walked on the CPU by
    tests/flat_emulator.py, gives the integers the reference's code wrote: PML, MS pointers + lengths, document numbers.
    No oracle and no GPU in the loop."""
    import spumoni_b200 as sp
    from flat_emulator import FlatImage

    def parse(path):
        return [[int(x) for x in line.split()] for line in open(path) if not line.startswith(">")]
    lines = [line.rstrip("\n") for line in open(os.path.join(GOLDEN, "reads.fa")) if not line.startswith(">")]
    reads = [r.upper().encode() for r in lines[:-3] + ["".join(lines[-3:])]]  # the last record spans three lines; :916-917 upper-cases
    ref = os.path.join(GOLDEN, "idx.fa")
    fp = FlatImage(sp.Index.open(ref, sp.MODE_PML, use_doc=True).flat())
    fm = FlatImage(sp.Index.open(ref, sp.MODE_MS, use_doc=True).flat())
    assert bool(fp.off_mrows8) == (row_width in ("8", "32"))  # -d images have no 32-byte rows: the request falls back to the default
    want = {k: parse(os.path.join(GOLDEN, "expected", m, "reads.fa." + k)) for m, k in
            (("Pd", "pseudo_lengths"), ("Pd", "doc_numbers"), ("Md", "lengths"), ("Md", "pointers"))}
    want_md = parse(os.path.join(GOLDEN, "expected", "Md", "reads.fa.doc_numbers"))
    assert len(reads) == len(want["pseudo_lengths"])
    for i, r in enumerate(reads):
        got, _, _ = fp.walk_rows(r)
        assert got == want["pseudo_lengths"][i] and fp.docs == want["doc_numbers"][i], i
        assert fp.walk(r)[0] == want["pseudo_lengths"][i], i
        _, ptr, _ = fm.walk_rows(r)
        assert ptr == want["pointers"][i] and fm.docs == want_md[i], i
        assert fm.ms_lengths(r, ptr) == want["lengths"][i], i


def test_cli_validation_equals_reference_binary(ref_bin, golden_copy, tmp_path):
    """Option parsing and validation of the product's `spumoni run` against the reference's own spumoni.cpp /
    spumoni_main.hpp (:267-329): same exit code and, apart from the version banner, the same stderr text. Every case fails
    (or prints usage) before an index is loaded, so no GPU is involved."""
    import re
    import spumoni_b200 as sp
    P, Q = os.path.join(golden_copy, "idx"), os.path.join(golden_copy, "reads.fa")
    bad = str(tmp_path / "reads.txt")
    open(bad, "w").write(">a\nACGT\n")
    cases = [["run", "-r", P, "-P", "-n"], ["run", "-p", Q, "-P", "-n"], ["run", "-r", P, "-p", Q, "-n"],
             ["run", "-r", P, "-p", Q, "-n", "-M", "-P"], ["run", "-r", P, "-p", bad, "-P", "-n"],
             ["run", "-r", P, "-p", Q, "-P", "-n", "-w", "10"], ["run", "-r", P, "-p", Q, "-P", "-n", "-w", "500"],
             ["run", "-r", P, "-p", Q, "-P"], ["run", "-r", P, "-p", Q, "-P", "-m", "-a"], ["run", "-r", P, "-p", Q, "-P", "-n", "-m"],
             ["run", "-r", P + "_missing", "-p", Q, "-P", "-n"], ["run", "-r", P, "-p", Q + "_missing.fa", "-P", "-n"],
             ["run", "-r", P, "-p", Q, "-P", "-n", "-K", "5"], ["run", "-r", P, "-p", Q, "-P", "-n", "-K", "4", "-W", "3"],
             ["run", "-r", P, "-p", Q, "-P", "-g", "-t", "2"], ["run", "-r", P, "-p", Q, "-P", "-n", "-g", "-t", "2"],
             ["run", "-r", P, "-p", Q, "-P", "-n", "-g", "-c"], ["run", "-r", P, "-p", Q, "-M", "-n", "-d", "-x"], ["run"], ["run", "-h"]]
    os.remove(os.path.join(golden_copy, "idx.fa.doc"))
    cases.append(["run", "-r", P, "-p", Q, "-P", "-n", "-d"])  # document array requested but absent
    os.remove(os.path.join(golden_copy, "idx.fa.thrbv.ms"))
    cases.append(["run", "-r", P, "-p", Q, "-M", "-n"])  # index of the requested type absent

    def clean(s):
        s = re.sub(r"\x1b\[[0-9;]*m", "", s).replace(" (spumoni_b200 query engine)", " ")
        return "\n".join(line.rstrip() for line in s.strip().splitlines())
    for c in cases:
        a = subprocess.run([ref_bin] + c, capture_output=True, text=True)
        b = subprocess.run([sp.CLI_PATH] + c, capture_output=True, text=True)
        assert a.returncode == b.returncode, c
        assert clean(a.stderr) == clean(b.stderr) and a.stdout == b.stdout, (c, clean(a.stderr)[-300:], clean(b.stderr)[-300:])


@pytest.mark.parametrize("docs", [1, 5])
def test_document_array_counts_against_reference_binary(ref_bin, tmp_path, docs):
    """One document (DocumentArray asks sdsl for width 0, which sdsl turns into 64) and five (3-bit entries): the array the
    reference builds equals the fixture builder's, and the oracle's -d outputs equal the reference's."""
    d = str(tmp_path)
    mkindex("genome", "--kind", "pangenome", "--len", 6000, "--haps", 8, "--snp", 0.01, "--indel", 0.001, "--seed", 31, "--out", d + "/raw.fa")
    ref = build_index(d, fasta=d + "/raw.fa", revcomp=True, keep=True, docs=docs)
    mk_doc = open(ref + ".doc", "rb").read()
    runs = subprocess.check_output([ref_bin, "x-build-pml", ref], text=True).split()[3]
    subprocess.check_call([ref_bin, "x-doc", ref, runs])
    assert open(ref + ".doc", "rb").read() == mk_doc
    _write_slp(ref)
    reads = d + "/reads.fa"
    mkindex("reads", "--ref", ref, "--n", 80, "--len", 60, "--len-max", 300, "--sub", 0.05, "--neg-frac", 0.3, "--seed", 5, "--out", reads,
            "--bin", d + "/r.bin")
    for ext in (".pmlnulldb", ".msnulldb"):
        shutil.copy(os.path.join(GOLDEN, "idx.fa.pmlnulldb"), ref + ext)
    for kind, flag, exts in (("pml", "-P", ("pseudo_lengths", "doc_numbers")), ("ms", "-M", ("lengths", "pointers", "doc_numbers"))):
        _run_ref(ref_bin, ref, reads, [flag, "-n", "-d"])
        for e in exts:
            shutil.move(reads + "." + e, d + "/want." + e)
        oracle.run(ref, reads, kind, use_doc=True)
        for e in exts:
            assert filecmp.cmp(reads + "." + e, d + "/want." + e, shallow=False), (kind, e)


def test_report_bin_widths_and_thresholds_against_reference_binary(ref_bin, pinned_index, tmp_path):
    """The classifier's two inputs, swept: bin widths around the read lengths (-w 50 ... 400) and null databases whose
    percentile value lands below the floor of 3, on a power of two, or high (threshold = max(percentile, 3) [+ 4 for PML])."""
    reads = str(tmp_path / "reads.fa")
    shutil.copy(pinned_index["reads_fa"], reads)
    ref = pinned_index["ref"]
    saved = {e: open(ref + e, "rb").read() for e in (".pmlnulldb", ".msnulldb")}
    try:
        for kind, flag in (("pml", "-P"), ("ms", "-M")):
            for w in (50, 51, 150, 399, 400):
                _run_ref(ref_bin, ref, reads, [flag, "-n", "-c", "-w", str(w)])
                shutil.move(reads + ".report", str(tmp_path / "want.report"))
                oracle.run(ref, reads, kind, report=True, bin_width=w)
                assert filecmp.cmp(reads + ".report", str(tmp_path / "want.report"), shallow=False), (kind, w)
        for name, stats in (("ones", [1] * 100), ("high", list(range(10, 40)) * 6), ("rare_high", [3] * 50 + [90] * 4), ("pow2", [16] * 10 + [2] * 10),
                            ("zeros", [0] * 30)):
            np.array(stats, dtype=np.uint64).tofile(str(tmp_path / "st.bin"))
            for e in (".pmlnulldb", ".msnulldb"):
                subprocess.check_call([MKINDEX, "nulldb", "--stats", str(tmp_path / "st.bin"), "--out", ref + e], stdout=subprocess.DEVNULL)
            for kind, flag in (("pml", "-P"), ("ms", "-M")):
                _run_ref(ref_bin, ref, reads, [flag, "-n", "-c"])
                shutil.move(reads + ".report", str(tmp_path / "want.report"))
                oracle.run(ref, reads, kind, report=True)
                assert filecmp.cmp(reads + ".report", str(tmp_path / "want.report"), shallow=False), (name, kind)
                # the product's own loader + threshold rule (host code of the C ABI) against the reference's report header
                import spumoni_b200 as sp
                db = sp.EmpNullDatabase().load(ref + (".pmlnulldb" if kind == "pml" else ".msnulldb"))
                header = open(str(tmp_path / "want.report")).readline()
                assert int(header.split("thr=")[1].split(")")[0]) == sp.max_value_threshold(db.percentile_value, kind == "pml"), (name, kind)
    finally:
        for e, data in saved.items():
            open(ref + e, "wb").write(data)
