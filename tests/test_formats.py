"""CPU: the product's host side (on-disk parse -> flat image) against the oracle, without a GPU.

The flat image is walked by tests/flat_emulator.py (a Python transcription of the kernel's step)."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

import oracle
import spumoni_b200 as sp
from conftest import MKINDEX, build_index
from conftest import load_bin_reads
from flat_emulator import FlatImage

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_vectors.json")


def test_flat_image_matches_golden_vectors(tmp_path, row_width):
    for ci, case in enumerate(json.load(open(GOLDEN))["cases"]):
        ref = build_index(str(tmp_path / f"c{ci}"), text=case["text"].encode("latin1"))
        pml, ms = sp.Index.open(ref, sp.MODE_PML), sp.Index.open(ref, sp.MODE_MS)
        assert pml.get_bwt_stats() == (case["n"], case["r"]) == ms.get_bwt_stats()
        fp, fm = FlatImage(pml.flat()), FlatImage(ms.flat())
        for rd in case["reads"]:
            read = rd["read"].encode("latin1")
            assert fp.walk(read)[0] == rd["pml"]
            _, ptr = fm.walk(read)
            assert [str(x) for x in ptr] == rd["ms_pointers"]
            assert fm.ms_lengths(read, ptr) == rd["ms_lengths"]
            if fp.num_fast:
                assert bool(fp.off_mrows8) == (row_width == "8")
                assert fp.walk_rows(read)[0] == rd["pml"]
                assert [str(x) for x in fm.walk_rows(read)[1]] == rd["ms_pointers"]


def test_flat_image_vs_oracle_random(tmp_path, row_width):
    rng = random.Random(17)
    for trial in range(12):
        n = rng.randint(50, 4000)
        alpha = rng.choice(["ACGT", "ACGTN", "ACGT\xe9"])
        if rng.random() < 0.4:  # repetitive: long runs, windows that overflow
            unit = "".join(rng.choice(alpha) for _ in range(rng.randint(2, 30)))
            text = "".join(unit if rng.random() < 0.93 else rng.choice(alpha) for _ in range(n // len(unit) + 1))
        else:
            text = "".join(rng.choice(alpha) for _ in range(n))
        tb = text.encode("latin1")
        ref = build_index(str(tmp_path / f"t{trial}"), text=tb, revcomp=(alpha == "ACGT" and trial % 2 == 0))
        po, mo = oracle.OracleIndex(ref, "pml"), oracle.OracleIndex(ref, "ms")
        fp, fm = FlatImage(sp.Index.open(ref, 0).flat()), FlatImage(sp.Index.open(ref, 1).flat())
        for _ in range(8):
            ln = rng.randint(1, 150)
            read = bytes(rng.choice(b"ACGTNY\xe9") for _ in range(ln)) if rng.random() < 0.4 else (tb[rng.randint(0, max(0, len(tb) - ln)):][:ln] or b"A")
            want = po.matching_statistics(read).tolist()
            assert fp.walk(read)[0] == want
            assert fp.walk_rows(read)[0] == want
            L, P = mo.matching_statistics(read)
            _, ptr = fm.walk(read)
            assert ptr == P.tolist() and fm.ms_lengths(read, ptr) == L.tolist()
            assert fm.walk_rows(read)[1] == P.tolist()


def test_flat_roundtrip_and_header(small_index):
    idx = sp.Index.open(small_index["ref"], sp.MODE_PML)
    blob = np.array(idx.flat(), copy=True)
    again = sp.Index.from_flat(blob)
    assert again.get_bwt_stats() == idx.get_bwt_stats()
    assert np.array_equal(again.flat(), blob)
    hdr = sp.Index.from_header(blob[: sp.flat_header_bytes()])
    assert hdr.get_bwt_stats() == idx.get_bwt_stats()
    with pytest.raises(sp.SpumoniError):
        hdr.flat()  # header-only index carries no host image
    with pytest.raises(sp.SpumoniError):
        sp.Index.from_flat(blob[:-8])


def test_corrupt_and_missing_files_fail_loudly(small_index, tmp_path):
    with pytest.raises(sp.SpumoniError) as e:
        sp.Index.open(str(tmp_path / "nope.fa"), sp.MODE_PML)
    assert e.value.code == 1  # SPB_ERR_IO
    src = small_index["ref"] + ".thrbv.spumoni"
    data = open(src, "rb").read()
    bad = str(tmp_path / "bad.fa")
    open(bad + ".thrbv.spumoni", "wb").write(data[: len(data) // 2])  # truncated
    with pytest.raises(sp.SpumoniError) as e:
        sp.Index.open(bad, sp.MODE_PML)
    assert e.value.code == 2  # SPB_ERR_FORMAT
    open(bad + ".thrbv.spumoni", "wb").write(data + b"\0" * 8)  # must be consumed exactly to EOF
    with pytest.raises(sp.SpumoniError):
        sp.Index.open(bad, sp.MODE_PML)
    corrupt = bytearray(data)
    corrupt[8 + 8 + 65 * 8] ^= 0x10  # F['A']
    open(bad + ".thrbv.spumoni", "wb").write(bytes(corrupt))
    with pytest.raises(sp.SpumoniError):
        sp.Index.open(bad, sp.MODE_PML)


def test_ms_index_requires_matching_text(small_index, tmp_path):
    ref = str(tmp_path / "x.fa")
    open(ref, "w").write(">s\nACGT\n")
    open(ref + ".thrbv.ms", "wb").write(open(small_index["ref"] + ".thrbv.ms", "rb").read())
    with pytest.raises(sp.SpumoniError):
        sp.Index.open(ref, sp.MODE_MS)


def test_nulldb_roundtrip(tmp_path):
    vals = np.array([3, 3, 3, 3, 3, 9, 9, 9, 9, 9, 9, 11, 11, 2, 0, 16], dtype=np.uint64)
    vals.tofile(str(tmp_path / "s.bin"))
    out = str(tmp_path / "x.pmlnulldb")
    subprocess.check_call([MKINDEX, "nulldb", "--stats", str(tmp_path / "s.bin"), "--out", out, "--ks-thr", "0.25"])
    db = sp.EmpNullDatabase().load(out)
    assert db.num_values == 16 and db.percentile_value == 9.0 and db.ks_stat_threshold == 0.25
    assert abs(db.mean_null_stat - float(vals.mean())) < 1e-12
    assert oracle.nulldb_percentile(out) == 9.0
    assert sp.max_value_threshold(db.percentile_value, True) == 13
    assert sp.max_value_threshold(1.0, False) == 3
    raw = open(out, "rb").read()  # width = max(ceil(log2(16)),1) = 4 bits (reference quirk: 16 itself does not fit)
    assert raw[32 + 8] == 4


def test_move_rows_saturation_fallbacks_host_side(tmp_path, row_width):
    """The sattest build shrinks the move-row fields (3-bit len/delta, thresholds <= 4, jump distance < 3 /
    one-word head scans); walking its flat image with the same limits must still reproduce the oracle
    (exercises the slow flag, the generic-step threshold class, far jumps and un-normalised states
    entering the generic step)."""
    import subprocess, sys, textwrap
    rng = random.Random(23)
    unit = "ACGTTGCAAGGCTTAACCGGT"
    text = "".join(unit if rng.random() < 0.9 else rng.choice("ACGT") for _ in range(120)) + "A" * 40 + "".join(rng.choice("ACGTN") for _ in range(300))
    ref = build_index(str(tmp_path), text=text.encode())
    env = dict(os.environ, SPUMONI_B200_LIB=sp.SATTEST_LIB_PATH, SPUMONI_B200_ROWS=row_width,
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    imgs = {}
    for mode in (0, 1):
        out = str(tmp_path / f"flat{mode}.npy")
        code = textwrap.dedent(f"""
            import numpy as np, spumoni_b200 as sp
            np.save({out!r}, np.array(sp.Index.open({ref!r}, {mode}).flat(), copy=True))
        """)
        subprocess.check_call([sys.executable, "-c", code], env=env)
        imgs[mode] = FlatImage(np.load(out))
    po, mo = oracle.OracleIndex(ref, "pml"), oracle.OracleIndex(ref, "ms")
    tb = text.encode()
    tot = {"slow": 0, "norm_iters": 0, "jump2": 0, "fast": 0, "end": 0}
    for _ in range(40):
        ln = rng.randint(5, 200)
        s0 = rng.randint(0, len(tb) - ln)
        read = bytearray(tb[s0:s0 + ln])
        for i in range(len(read)):
            if rng.random() < 0.05:
                read[i] = rng.choice(b"ACGTN")
        got, _, st = imgs[0].walk_rows(bytes(read), len_bits=3, dist_far=3, probes=2)
        assert got == po.matching_statistics(bytes(read)).tolist()
        _, ptr, _ = imgs[1].walk_rows(bytes(read), len_bits=3, dist_far=3, probes=2)
        assert ptr == mo.matching_statistics(bytes(read))[1].tolist()
        for key in tot:
            tot[key] += st[key]
    assert tot["slow"] > 0 and tot["norm_iters"] > 0 and tot["jump2"] > 0 and tot["fast"] > 0


def test_document_array_tables_vs_oracle(small_index, row_width):
    """-d: the document tables of the flat image (per move row and per character record), walked on the CPU,
    reproduce the oracle's document numbers (compute_ms_pml.cpp:288-340 / :625-682), both row layouts, PML and MS."""
    ref = small_index["ref"]
    buf, off = load_bin_reads(small_index["reads_bin"])
    fp = FlatImage(sp.Index.open(ref, sp.MODE_PML, use_doc=True).flat())
    fm = FlatImage(sp.Index.open(ref, sp.MODE_MS, use_doc=True).flat())
    assert fp.num_docs == 3 and fm.num_docs == 3
    po, mo = oracle.OracleIndex(ref, "pml", use_doc=True), oracle.OracleIndex(ref, "ms", use_doc=True)
    reads = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(0, 60)] + [b"ACGTNNACGT" * 5, b"N", b"ACG\xe9T" * 8]
    rb, ro = sp.pack_reads(reads)
    Lp, _, Dp = po.query_batch_doc(rb, ro)
    Lm, Pm, Dm = mo.query_batch_doc(rb, ro)
    for i, r in enumerate(reads):
        a, b = int(ro[i]), int(ro[i + 1])
        got, _, _ = fp.walk_rows(r)
        assert got == Lp[a:b].tolist() and fp.docs == Dp[a:b].tolist(), (i, r[:20])
        _, ptr, _ = fm.walk_rows(r)
        assert ptr == Pm[a:b].tolist() and fm.docs == Dm[a:b].tolist(), (i, r[:20])
    # an index opened without the document array has no document tables, and a missing .doc file fails loudly
    assert FlatImage(sp.Index.open(ref, sp.MODE_PML).flat()).num_docs == 0
    os.rename(ref + ".doc", ref + ".doc.bak")
    try:
        with pytest.raises(sp.SpumoniError):
            sp.Index.open(ref, sp.MODE_PML, use_doc=True)
    finally:
        os.rename(ref + ".doc.bak", ref + ".doc")


def test_ms_sample_layouts(small_index, monkeypatch):
    """MS over compact rows: the SA samples either ride inside 16-byte row slots (text positions < 2^32, the
    default) or live in the separate per-row table (SPUMONI_B200_MSAMP=wide, what n >= 2^32 would use)."""
    buf, off = load_bin_reads(small_index["reads_bin"])
    mo = oracle.OracleIndex(small_index["ref"], "ms")
    monkeypatch.setenv("SPUMONI_B200_ROWS", "8")
    imgs = {}
    for layout in ("inline", "wide", "split32"):
        monkeypatch.setenv("SPUMONI_B200_MSAMP", layout)
        imgs[layout] = FlatImage(sp.Index.open(small_index["ref"], sp.MODE_MS).flat())
    assert imgs["inline"].mrows8_stride == 16 and imgs["inline"].off_msamp == 0
    assert imgs["wide"].mrows8_stride == 8 and imgs["wide"].off_msamp != 0 and imgs["wide"].msamp_stride == 16
    assert imgs["split32"].mrows8_stride == 8 and imgs["split32"].off_msamp != 0 and imgs["split32"].msamp_stride == 8
    assert imgs["split32"].total_bytes < imgs["wide"].total_bytes
    for i in range(40):
        r = bytes(buf[int(off[i]):int(off[i + 1])])
        want = mo.matching_statistics(r)[1].tolist()
        for layout in imgs:
            assert imgs[layout].walk_rows(r)[1] == want, layout


def test_nulldb_save_matches_fixture_writer(tmp_path):
    """spb_nulldb_save (EmpNullDatabase::serialize, emp_null_database.cpp:82-102) writes the bytes the fixture writer does,
    including the width quirk (a maximum that is an exact power of two is truncated)."""
    import ctypes
    lib = sp.load_library()
    for vals in ([3, 3, 3, 3, 3, 9, 9, 9, 9, 9, 9, 11, 11, 2, 0, 16], [0, 0, 0], [1, 0, 1, 1, 1, 1], [5] * 7 + [700]):
        a = np.array(vals, dtype=np.uint64)
        a.tofile(str(tmp_path / "s.bin"))
        subprocess.check_call([MKINDEX, "nulldb", "--stats", str(tmp_path / "s.bin"), "--out", str(tmp_path / "f.db"), "--ks-thr", "0.5"])
        ref = sp.EmpNullDatabase().load(str(tmp_path / "f.db"))
        db = sp._NullDb(len(vals), 0.5, ref.mean_null_stat, ref.percentile_value)
        s32 = a.astype(np.uint32)
        assert lib.spb_nulldb_save(str(tmp_path / "g.db").encode(), ctypes.byref(db), s32.ctypes.data) == 0
        assert open(tmp_path / "g.db", "rb").read() == open(tmp_path / "f.db", "rb").read(), vals


def test_positions_beyond_2_32(huge_periodic_index, row_width):
    """n > 2^32: on-disk parse (sd_vectors over a 4.3 G universe), 40-bit fields of the run / character records,
    saturated move rows falling back to the generic step -- the flat image walked on the CPU equals the oracle."""
    ref, reads = huge_periodic_index["ref"], huge_periodic_index["reads"]
    idx = sp.Index.open(ref, sp.MODE_PML)
    assert idx.get_bwt_stats()[0] == huge_periodic_index["n"]
    fp = FlatImage(idx.flat())
    po = oracle.OracleIndex(ref, "pml")
    assert po.stats()[0] == huge_periodic_index["n"]
    slow = 0
    for r in reads:
        want = po.matching_statistics(r).tolist()
        assert fp.walk(r)[0] == want
        got, _, st = fp.walk_rows(r)
        assert got == want
        slow += st["slow"]
    assert slow > 0  # run lengths of ~10^8 do not fit the row fields


@pytest.mark.parametrize("balance", ["2,1,3", "3,2,0", "0"])
def test_balanced_rows_keep_the_semantics(tmp_path, row_width, balance, monkeypatch):
    """balance_runs (index_host.cpp) cuts runs into rows -- here far more aggressively than the defaults: any row whose LF
    image overlaps more than 2 rows, and any row longer than 3 -- and every table of the flat image is then built over the
    rows.  The walk over the rows (move rows and generic tables, PML and MS incl. samples) must still equal the oracle on the
    unmodified index, get_bwt_stats must still report the runs of the file, and the walk must get shorter."""
    rng = random.Random(5)
    unit = "ACGTTGCAAGGCTTAACCGGT"
    text = "".join(unit if rng.random() < 0.9 else rng.choice("ACGT") for _ in range(150)) + "A" * 70 + "".join(rng.choice("ACGTN") for _ in range(300)) + "C" * 50
    ref = build_index(str(tmp_path), text=text.encode())
    po, mo = oracle.OracleIndex(ref, "pml"), oracle.OracleIndex(ref, "ms")
    monkeypatch.setenv("SPUMONI_B200_BALANCE", balance)
    ip, im = sp.Index.open(ref, sp.MODE_PML), sp.Index.open(ref, sp.MODE_MS)
    fp, fm = FlatImage(ip.flat()), FlatImage(im.flat())
    assert ip.get_bwt_stats() == (fp.n, fp.R_file) and fp.R_file == po.stats()[1]
    assert (fp.R > fp.R_file) == (balance != "0")
    tb = text.encode()
    walked = 0
    for _ in range(40):
        ln = rng.randint(5, 250)
        s0 = rng.randint(0, len(tb) - ln)
        read = bytearray(tb[s0:s0 + ln])
        for i in range(len(read)):
            if rng.random() < 0.04:
                read[i] = rng.choice(b"ACGTN")
        want = po.matching_statistics(bytes(read)).tolist()
        got, _, st = fp.walk_rows(bytes(read))
        assert got == want and fp.walk(bytes(read))[0] == want
        wl, wp = mo.matching_statistics(bytes(read))
        _, ptr, _ = fm.walk_rows(bytes(read))
        assert ptr == wp.tolist() and fm.walk(bytes(read))[1] == wp.tolist()
        walked += st["norm_iters"]
    if balance == "2,1,3":
        assert walked <= 40 * 250  # at most one extra row per step


def test_deep_verification_of_stored_supports(small_index, tmp_path):
    """spb_index_verify re-derives every select_support_mcl / rank_support_v stored in the index from the decoded bits and
    compares with the file: green on indexes written by tools/mkindex and by the reference's own serialize() (golden),
    red -- naming the field -- when a byte inside one of those (normally skipped) sections is flipped."""
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run", "idx.fa")
    for ref in (small_index["ref"], golden):
        sp.verify_index(ref, sp.MODE_PML)
        sp.verify_index(ref, sp.MODE_MS)
    import shutil
    ref = str(tmp_path / "idx.fa")
    raw = bytearray(open(small_index["ref"] + ".thrbv.spumoni", "rb").read())
    # the first sparse_sd_vector (bwt.runs) starts after terminator_position, F (8 + 2048) and n, R, B: u, n, size, wl, low, high, select1 ...
    base = 8 + 8 + 2048 + 24
    seen = set()
    for probe in range(base + 200, base + 6000, 97):
        bad = bytearray(raw)
        bad[probe] ^= 0x10
        open(ref + ".thrbv.spumoni", "wb").write(bad)
        plain_ok = True
        try:
            sp.Index.open(ref, sp.MODE_PML)
        except sp.SpumoniError:
            plain_ok = False
        try:
            sp.verify_index(ref, sp.MODE_PML)
            assert plain_ok, probe  # what the plain loader rejects, the deep check rejects too
            seen.add("clean")
        except sp.SpumoniError as e:
            seen.add("support" if "derived from the decoded bits" in str(e) else "data")
            assert "bwt." in str(e) or "thresholds" in str(e) or "run-length" in str(e) or "F" in str(e), str(e)
    assert "support" in seen  # some probes landed in a select support: invisible to the plain loader, caught by the deep check
