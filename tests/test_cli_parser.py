"""CPU: the product CLI's input parser (FastaSource / LineReader in spumoni_b200/csrc/spumoni_run.cpp, compiled as is into
tests/harness/cli_parser_harness.cpp) against the reference's BatchLoader (src/batch_loader.cpp:26-131), observed through
the reference's own outputs: the record ids it wrote and the number of values per record. FASTA and FASTQ records, CRLF,
blank lines, multi-line sequences, headers with blanks, '>' inside a sequence, missing trailing newline, and the inputs
both sides reject."""
import os
import random
import subprocess

import pytest

import spumoni_b200 as sp
from conftest import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_run")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "spumoni_ref")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("harness") / "cli_parser_harness")
    lib_dir = os.path.dirname(sp.LIB_PATH)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "harness", "cli_parser_harness.cpp"), "-L" + lib_dir, "-lspumoni_b200",
                           "-Wl,-rpath," + lib_dir])
    return exe


def _parsed(harness, path):
    r = subprocess.run([harness, path], capture_output=True)
    out = r.stdout.split(b"\n")
    return r.returncode, [(out[i], out[i + 1]) for i in range(0, len(out) - 1, 2)]


def _reference_records(path):
    lines = open(path, "rb").read().split(b"\n")
    return [(lines[i], len(lines[i + 1].split())) for i in range(0, len(lines) - 1, 2)]


@pytest.mark.parametrize("pattern,mode,ext", [("reads.fa", "P", "pseudo_lengths"), ("fastq_reads.fa", "Pq", "pseudo_lengths"),
                                              ("fastq_reads.fa", "Mq", "lengths")])
def test_parser_equals_committed_reference_records(harness, pattern, mode, ext):
    rc, got = _parsed(harness, os.path.join(GOLDEN, pattern))
    want = _reference_records(os.path.join(GOLDEN, "expected", mode, pattern + "." + ext))
    assert rc == 0 and [(i, len(s)) for i, s in got] == want
    if pattern == "fastq_reads.fa":  # sequences are the second line of every kept record
        lines = open(os.path.join(GOLDEN, pattern), "rb").read().split(b"\n")
        assert [s for _, s in got] == [lines[4 * k + 1] for k in range(len(got))] and len(got) == 21


def _nasty_files(d, text):
    rng = random.Random(4)

    def fq(n, length, trail=True):
        out = ""
        for i in range(n):
            p = rng.randint(0, len(text) - length - 1)
            out += "@q%d some comment\n%s\n+\n%s\n" % (i, text[p:p + length], "I" * length)
        return out if trail else out.rstrip("\n")
    t = text
    cases = {
        "crlf": ">r1\r\n" + t[10:60] + "\r\n>r2 desc\r\n" + t[100:140] + "\r\n",
        "blank_lines": "\n\n>r1\n" + t[10:60] + "\n\n>r2\n\n" + t[100:140] + "\n\n",
        "multiline": ">r1\n" + t[10:40] + "\n" + t[40:70] + "\n" + t[70:75] + "\n>r2\n" + t[100:140] + "\n",
        "notrail": ">r1\n" + t[10:60] + "\n>r2\n" + t[100:140],
        "tabs": ">r1\tx y\n" + t[10:60] + "\n>r2  two  spaces\n" + t[100:140] + "\n",
        "gt_in_seq": ">r1\n" + t[10:30] + ">" + t[30:60] + "\n>r2\n" + t[100:140] + "\n",
        "no_header_first": t[10:60] + "\n>r2\n" + t[100:140] + "\n",
        "spaces_in_seq": ">r1\n" + t[10:30] + " " + t[30:60] + "\n>r2\n" + t[100:140] + " \n",
        "lower": ">r1\n" + t[10:60].lower() + "\n",
        "empty_id": ">\n" + t[10:60] + "\n> \n" + t[100:140] + "\n",
        "one_char_id": ">r\n" + t[10:60] + "\n>r\n" + t[100:140] + "\n",
        "long_id": ">" + "x" * 300 + " d\n" + t[10:60] + "\n",
        "short_reads": "".join(">s%d\n%s\n" % (i, t[i * 3:i * 3 + 1 + i % 4]) for i in range(40)),
        "fq_short": fq(23, 150), "fq_short_notrail": fq(23, 150, False), "fq_long": fq(5, 1200), "fq_long_notrail": fq(5, 1200, False),
        "fq_mixed": fq(9, 150) + fq(2, 1500) + fq(5, 60), "fq_one": fq(1, 100), "fq_one_notrail": fq(1, 100, False),
        "fq_crlf": fq(12, 150).replace("\n", "\r\n"), "fq_truncated": fq(10, 400) + "@last\nACGT\n",
        "fq_blank_mid": fq(3, 600) + "\n" + fq(3, 600), "fq_plus_named": fq(4, 500).replace("\n+\n", "\n+same id\n"),
        "fq_bad_header": fq(6, 300) + "Xnot\nACGT\n+\nIIII\n" + fq(6, 300), "fq_two_char_header": fq(6, 300) + "@a\nACGT\n+\nIIII\n" + fq(2, 300),
        "fq_empty_seq": fq(6, 300) + "@e1 x\n\n+\n\n" + fq(6, 300), "fq_tab_ids": fq(30, 90).replace(" some", "\tsome"),
        "fq_bare_ids": fq(30, 90).replace(" some comment", ""), "fq_cut_short_bad_header": fq(10, 400) + "Zlast\nACGT",
        "fq_trailing_blank_lines": fq(10, 400) + "\n\n", "fq_many": fq(700, 37) + fq(40, 1001) + fq(300, 5),
    }
    paths = {}
    for name, body in cases.items():
        paths[name] = os.path.join(d, name + ".fa")
        with open(paths[name], "w", newline="") as f:
            f.write(body)
    return paths


def test_parser_equals_reference_binary_on_odd_inputs(harness, tmp_path):
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "-s"])
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/spumoni_ref is not built (needs /root/reference)")
    import shutil
    d = str(tmp_path)
    for f in os.listdir(GOLDEN):
        if f.startswith("idx.fa"):
            shutil.copy(os.path.join(GOLDEN, f), os.path.join(d, f))
    text = "".join(line.strip() for line in open(os.path.join(d, "idx.fa")) if not line.startswith(">"))
    seen_reject = seen_drop = 0
    for name, path in sorted(_nasty_files(d, text).items()):
        ref = subprocess.run([REF_BIN, "run", "-r", os.path.join(d, "idx"), "-p", path, "-P", "-n"], capture_output=True)
        rc, got = _parsed(harness, path)
        assert (rc != 0) == (ref.returncode != 0), name
        if ref.returncode != 0:
            seen_reject += 1
            continue
        want = _reference_records(path + ".pseudo_lengths")
        assert [(i, len(s)) for i, s in got] == want, name
        seen_drop += name.startswith("fq_") and len(want) < open(path).read().count("\n@") + 1
    assert seen_reject >= 4 and seen_drop >= 4


def test_fast_window_parser_equals_sequential_parser(harness, tmp_path):
    """The fast path of `spumoni run` (memory-mapped FASTA or FASTQ cut into windows, parsed on several threads) must produce exactly the
    records of the sequential BatchLoader-equivalent parser -- ids with their delimiter quirk, stripped multi-line sequences,
    CRLF, '>' inside a line -- for any window size and thread count, and must hand over to the sequential parser (same
    records, same rejection) where a header has no id or a record is empty."""
    d = str(tmp_path)
    text = "".join(line.strip() for line in open(os.path.join(GOLDEN, "idx.fa")) if not line.startswith(">"))
    files = _nasty_files(d, text)
    files["golden"] = os.path.join(GOLDEN, "reads.fa")
    rng = random.Random(9)
    big = os.path.join(d, "big.fa")
    with open(big, "w") as f:
        for i in range(3000):
            ln = rng.choice([1, 7, 60, 150, 400])
            p = rng.randint(0, len(text) - ln - 1)
            body = text[p:p + ln]
            if i % 11 == 0:
                body = body[: ln // 2] + "\n" + body[ln // 2:]
            f.write(">read%d%s\n%s\n" % (i, rng.choice(["", " d", "\tq", " "]), body))
    files["big"] = big
    files["golden_fq"] = os.path.join(GOLDEN, "fastq_reads.fa")
    bigq = os.path.join(d, "big.fq")
    with open(bigq, "w", newline="") as f:  # ragged FASTQ: the batch replay decides which tail records BatchLoader keeps
        for i in range(2500):
            ln = rng.choice([1, 7, 60, 150, 400, 2100])
            p = rng.randint(0, len(text) - ln - 1)
            eol = rng.choice(["\n", "\n", "\r\n"])
            f.write("@read%d%s%s%s%s+%s%s%s" % (i, rng.choice(["", " d", "\tq", " "]), eol, text[p:p + ln], eol, eol, "F" * ln,
                                              eol if i < 2499 else ""))
    files["big_fq"] = bigq
    late = 0
    anomalous = {"fq_bad_header", "fq_two_char_header", "fq_empty_seq", "fq_cut_short_bad_header", "fq_blank_mid"}
    for name, path in sorted(files.items()):
        r0 = subprocess.run([harness, path], capture_output=True)
        for window, threads in ((1, 1), (64, 3), (300, 8), (1 << 20, 4)):
            r1 = subprocess.run([harness, path, str(window), str(threads)], capture_output=True)
            assert r1.returncode == r0.returncode, (name, window, threads)
            if name.startswith("fq_") or name.endswith("_fq"):  # ordinary FASTQ stays on the fast path, anomalies hand over
                assert (b"fastq:sequential" if name in anomalous else b"fastq:fast") in r1.stderr, name
            if r0.returncode == 0:
                assert r1.stdout == r0.stdout, (name, window, threads)
        if name.startswith("fq_") or name.endswith("_fq"):  # FASTQ in segments: each begins at a batch boundary found by the replay
            for segment in (1, 700, 5000, 40000):
                r2 = subprocess.run([harness, path, "900", "3", str(segment)], capture_output=True)
                assert r2.returncode == r0.returncode and (r0.returncode != 0 or r2.stdout == r0.stdout), (name, segment)
                late += b"fastq:sequential from" in r2.stderr and b"from 0\n" not in r2.stderr
    assert late >= 3  # anomalies behind the first segment: the sequential parser took over mid-file, at a batch boundary


def test_fast_fastq_parser_random_files(harness, tmp_path):
    """Random four-line files -- ragged record lengths around the 1000-base batch limit, CRLF, a missing final newline, cut
    short records, and now and then an empty or malformed header or an empty sequence: whatever the batch replay decides
    (fast windows or hand-over to the sequential parser), records and exit status equal the sequential parser's."""
    rng = random.Random(77)
    fast = handed = rejected = pinned = midfile = 0
    ref_prefix = None
    if os.path.exists(REF_BIN):  # where the reference's own binary exists, the sequential parser is held to it on the same files
        import shutil
        for f in os.listdir(GOLDEN):
            if f.startswith("idx.fa"):
                shutil.copy(os.path.join(GOLDEN, f), str(tmp_path / f))
        ref_prefix = str(tmp_path / "idx")
    for trial in range(120):
        n = rng.choice([1, 2, 3, 7, 40, 200])
        eol = rng.choice(["\n", "\n", "\r\n"])
        odd = rng.random() < 0.35
        lines = []
        for i in range(n):
            ln = rng.choice([1, 5, 90, 150, 499, 500, 501, 700, 1999, 2000, 2600])
            hdr = "@r%d%s" % (i, rng.choice(["", " x", "\ty z"]))
            seq = "".join(rng.choice("ACGT") for _ in range(ln))
            if odd and rng.random() < 0.08:
                hdr = rng.choice(["", "@", "@a", "Xr%d" % i, ">r%d" % i])
            if odd and rng.random() < 0.04:
                seq = rng.choice(["", " "])
            lines += [hdr, seq, "+" + rng.choice(["", hdr[1:]]), "I" * len(seq)]
        lines = lines[: len(lines) - rng.choice([0, 0, 0, 1, 2, 3])]
        if not lines or not lines[0].startswith("@"):
            continue  # not recognised as FASTQ at all: the FASTA tests cover that entry
        body = eol.join(lines) + rng.choice([eol, eol, "", eol + eol])
        path = str(tmp_path / ("t%d.fa" % trial))
        with open(path, "w", newline="") as f:
            f.write(body)
        r0 = subprocess.run([harness, path], capture_output=True)
        window, threads = rng.choice([(1, 2), (700, 3), (5000, 8), (1 << 20, 4)])
        segment = rng.choice([1, 300, 2500, 20000, 1 << 40])
        r1 = subprocess.run([harness, path, str(window), str(threads), str(segment)], capture_output=True)
        assert r1.returncode == r0.returncode, (trial, body[:200])
        if r0.returncode == 0:
            assert r1.stdout == r0.stdout, (trial, window, threads, segment)
        else:
            rejected += 1
        if ref_prefix and trial % 2 == 0:
            ref = subprocess.run([REF_BIN, "run", "-r", ref_prefix, "-p", path, "-P", "-n"], capture_output=True)
            assert (ref.returncode != 0) == (r0.returncode != 0), (trial, body[:200])
            if ref.returncode == 0:
                out = r0.stdout.split(b"\n")
                got = [(out[i], len(out[i + 1])) for i in range(0, len(out) - 1, 2)]
                assert got == _reference_records(path + ".pseudo_lengths"), trial
                pinned += 1
        fast += b"fastq:fast" in r1.stderr
        handed += b"fastq:sequential" in r1.stderr
        midfile += b"fastq:sequential from" in r1.stderr and b"from 0\n" not in r1.stderr
    assert midfile >= 5
    assert fast >= 40 and handed >= 10 and rejected >= 5 and (pinned >= 30 or not ref_prefix)
