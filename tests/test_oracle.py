"""CPU: pins the oracle stack (naive Python + C++ restatement) against golden vectors."""
import json
import os
import random

import numpy as np
import pytest

import oracle
from oracle import naive
from conftest import build_index

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_vectors.json")

# SURVEY.md Appendix D (survey-generated golden index + three reads)
APPD_TEXT = b"GATTACAGATTACCGATTAGA"
APPD = {
    b"TTACAGGATTAG": ([4, 3, 2, 1, 0, 0, 3, 2, 1, 0, 0, 0], [2, 3, 4, 5, 6, 14, 14, 15, 16, 17, 6, 14],
                      [6, 5, 4, 3, 2, 1, 6, 5, 4, 3, 2, 1]),
    b"CCGATNAGATT": ([1, 0, 3, 2, 1, 0, 0, 2, 1, 0, 0],
                     [12, 13, 18446744073709551613, 18446744073709551614, 18446744073709551615, 0, 6, 14, 15, 16, 16],
                     [5, 4, 3, 2, 1, 0, 5, 4, 3, 2, 1]),
    b"GATTYACA": ([2, 1, 0, 0, 0, 1, 0, 1], [14, 15, 16, 16, 0, 4, 5, 15], [4, 3, 2, 1, 0, 3, 2, 1]),
}


def test_naive_reproduces_appendix_d_index():
    idx = naive.build_naive_index(APPD_TEXT)
    assert idx.sa == [21, 20, 4, 11, 18, 6, 1, 8, 15, 5, 12, 13, 19, 0, 7, 14, 3, 10, 17, 2, 9, 16]
    assert idx.bwt.replace(b"\x01", b"$") == b"AGTTTCGGGAACA$ACTTTAAA"
    assert idx.heads.replace(b"\x01", b"$") == b"AGTCGACA$ACTA"
    assert idx.lens == [1, 1, 3, 1, 3, 2, 1, 1, 1, 1, 1, 3, 3]
    assert idx.thr == [0, 0, 0, 0, 2, 1, 9, 12, 0, 13, 12, 9, 16]
    assert [idx.F[c] for c in (1, 65, 67, 71, 84)] == [0, 1, 9, 12, 16]


@pytest.mark.parametrize("read", list(APPD))
def test_naive_reproduces_appendix_d_reads(read):
    idx = naive.build_naive_index(APPD_TEXT)
    pml, ptr, ln = APPD[read]
    assert naive.pml_query(idx, read) == pml
    L, P = naive.ms_query(idx, read)
    assert P == ptr and L == ln
    assert ln == naive.brute_force_ms(APPD_TEXT, read)


@pytest.mark.parametrize("read", list(APPD))
def test_cpp_oracle_reproduces_appendix_d(read, tmp_path):
    ref = build_index(str(tmp_path), text=APPD_TEXT)
    pml, ptr, ln = APPD[read]
    assert oracle.OracleIndex(ref, "pml").matching_statistics(read).tolist() == pml
    L, P = oracle.OracleIndex(ref, "ms").matching_statistics(read)
    assert P.tolist() == ptr and L.tolist() == ln
    assert oracle.OracleIndex(ref, "pml").stats() == (22, 13)


def test_cpp_oracle_matches_golden_vectors(tmp_path):
    cases = json.load(open(GOLDEN))["cases"]
    for ci, case in enumerate(cases):
        ref = build_index(str(tmp_path / f"c{ci}"), text=case["text"].encode("latin1"))
        p, m = oracle.OracleIndex(ref, "pml"), oracle.OracleIndex(ref, "ms")
        assert p.stats() == (case["n"], case["r"])
        for rd in case["reads"]:
            read = rd["read"].encode("latin1")
            assert p.matching_statistics(read).tolist() == rd["pml"]
            L, P = m.matching_statistics(read)
            assert [str(x) for x in P.tolist()] == rd["ms_pointers"]
            assert L.tolist() == rd["ms_lengths"]


def test_cpp_oracle_vs_naive_random(tmp_path):
    rng = random.Random(3)
    for trial in range(25):
        n = rng.randint(5, 600)
        alpha = rng.choice(["ACGT", "ACGTN", "AC", "ACGT\xe9"])
        text = "".join(rng.choice(alpha) for _ in range(n)).encode("latin1")
        ref = build_index(str(tmp_path / f"t{trial}"), text=text)
        idx = naive.build_naive_index(text)
        p, m = oracle.OracleIndex(ref, "pml"), oracle.OracleIndex(ref, "ms")
        for _ in range(6):
            ln = rng.randint(1, 70)
            read = bytes(rng.choice(b"ACGTNY\xe9") for _ in range(ln)) if rng.random() < 0.4 else \
                bytes(text[rng.randint(0, max(0, n - ln)):][:ln]) or b"A"
            assert p.matching_statistics(read).tolist() == naive.pml_query(idx, read)
            L, P = m.matching_statistics(read)
            l1, p1 = naive.ms_query(idx, read)
            assert P.tolist() == p1 and L.tolist() == l1


def test_ms_equals_brute_force_and_bounds_pml(small_index):
    """Independent ground truth: with no absent characters the reference MS lengths are true
    matching statistics; PML <= MS; X[i] <= X[i+1] + 1."""
    from conftest import load_bin_reads
    ref = small_index["ref"]
    text = "".join(l.strip() for l in open(ref) if not l.startswith(">")).encode()
    buf, off = load_bin_reads(small_index["reads_bin"])
    p, m = oracle.OracleIndex(ref, "pml"), oracle.OracleIndex(ref, "ms")
    for r in range(0, 40):
        read = bytes(buf[int(off[r]):int(off[r + 1])])
        pml = p.matching_statistics(read)
        L, _ = m.matching_statistics(read)
        if r < 12:
            assert L.tolist() == naive.brute_force_ms(text, read)
        assert np.all(pml <= L)
        assert np.all(L[:-1] <= L[1:] + 1) and np.all(pml[:-1] <= pml[1:] + 1)


def test_bin_rule_and_threshold():
    for L, bw, nb in ((1000, 150, 6), (300, 150, 2), (299, 150, 1), (149, 150, 1), (450, 150, 3), (50, 50, 1), (101, 50, 2)):
        assert naive.classify_bins([1] * L, bw, 3)[4] == nb == max(1, L // bw)
    assert naive.max_value_threshold(2.0, pml=True) == 7
    assert naive.max_value_threshold(9.0, pml=True) == 13
    assert naive.max_value_threshold(9.0, pml=False) == 9
    assert naive.max_value_threshold(9.9, pml=False, use_dna_letters=True) == 10


def test_oracle_run_writes_reference_shaped_files(small_index, tmp_path):
    """orc_run output format: '>' id NL, values each followed by one space, NL (compute_ms_pml.cpp:1003-1010)."""
    import shutil, subprocess
    from conftest import MKINDEX
    reads = str(tmp_path / "q.fa")
    with open(reads, "w") as f:
        f.write(">r1 some description\nacgtacgtNN\nACGTAC\n>r2\tx\nTTTTTTTTTTGGGGG\n")
    stats = np.arange(1, 40, dtype=np.uint64) % 7
    stats.tofile(str(tmp_path / "stats.bin"))
    subprocess.check_call([MKINDEX, "nulldb", "--stats", str(tmp_path / "stats.bin"), "--out", small_index["ref"] + ".pmlnulldb"])
    nr, _ = oracle.run(small_index["ref"], reads, "pml", report=True, bin_width=150, threads=1)
    assert nr == 2
    lines = open(reads + ".pseudo_lengths").read().split("\n")
    assert lines[0] == ">r1 " and lines[2] == ">r2\t"  # id keeps its trailing delimiter (batch_loader.cpp:104-106)
    assert lines[1].endswith(" ") and len(lines[1].split()) == 16
    rep = open(reads + ".report").read().split("\n")
    assert rep[0].startswith("read id:") and "avg max-value (thr=" in rep[0]
    assert rep[1].startswith("r1 ") and ("FOUND" in rep[1] or "NOT_PRESENT" in rep[1])
    assert oracle.nulldb_percentile(small_index["ref"] + ".pmlnulldb") == 6.0


def test_cpp_oracle_document_numbers_vs_naive(tmp_path):
    """-d: the C++ oracle's doc variants (compute_ms_pml.cpp:288-340, :625-682) and the fixture's .doc writer against
    the literal Python restatement and a naive DocumentArray constructor (doc_array.cpp:18-108)."""
    rng = random.Random(5)
    base = "".join(rng.choice("ACGT") for _ in range(90))
    recs = ["".join((ch if rng.random() > 0.1 else rng.choice("ACGT")) for ch in base) for _ in range(5)]
    fa = str(tmp_path / "in.fa")
    open(fa, "w").write("".join(">s%d\n%s\n" % (i, r) for i, r in enumerate(recs)))
    ref = build_index(str(tmp_path), fasta=fa, docs=3)
    text = "".join(recs).encode()
    idx = naive.build_naive_index(text)
    lens = [int(l.split("\t")[1]) for l in open(ref + ".fdi")]
    assert sum(lens) == len(text) and len(lens) == 3
    sd, ed = naive.build_doc_array(idx, lens)
    assert sorted(set(sd + ed)) == [0, 1, 2]
    po, mo = oracle.OracleIndex(ref, "pml", use_doc=True), oracle.OracleIndex(ref, "ms", use_doc=True)
    for _ in range(60):
        ln = rng.randint(1, 80)
        s0 = rng.randint(0, len(text) - ln)
        r = bytearray(text[s0:s0 + ln])
        for i in range(len(r)):
            if rng.random() < 0.1:
                r[i] = rng.choice(b"ACGTN")
        r = bytes(r)
        buf, off = np.frombuffer(r, dtype=np.uint8), np.array([0, len(r)], dtype=np.uint64)
        L, _, D = po.query_batch_doc(buf, off)
        nl, nd = naive.pml_query_doc(idx, r, sd, ed)
        assert L.tolist() == nl and D.tolist() == nd
        assert L.tolist() == naive.pml_query(idx, r)  # the doc variant computes the same lengths
        Lm, Pm, Dm = mo.query_batch_doc(buf, off)
        np_, ndm = naive.ms_pointers_query_doc(idx, r, sd, ed)
        assert Pm.tolist() == np_ and Dm.tolist() == ndm


def test_cpp_oracle_null_database_vs_naive(small_index):
    """generate_null_pml_statistics + EmpNullDatabase summary (compute_ms_pml.cpp:1448-1486, emp_null_database.cpp:42-79)."""
    from conftest import load_bin_reads
    buf, off = load_bin_reads(small_index["reads_bin"])
    buf, off = buf[: int(off[12])], off[:13]
    o = oracle.OracleIndex(small_index["ref"], "pml")
    stats, mean, perc = o.nulldb_generate(buf, off)
    want = []
    for i in range(12):
        r = bytes(buf[int(off[i]):int(off[i + 1])]).upper()[::-1]
        want += o.matching_statistics(r).tolist()
    assert stats.tolist() == want and mean == sum(want) / len(want)
    from collections import Counter
    common = [v for v, c in Counter(want).items() if c >= 5]
    assert perc == (max(common) if common else 0)
