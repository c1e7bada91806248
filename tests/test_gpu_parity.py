"""GPU: parity of the CUDA path (through the C ABI) with the oracle. Run with `-m gpu` on a B200."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

import oracle
import spumoni_b200 as sp
from oracle import naive
from conftest import MKINDEX, build_index, load_bin_reads

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_vectors.json")
LANES = [1, 4, 8, 16, 32]  # 1 = thread-per-read kernels over the move rows (PML and MS); others = cooperative kernel


def ms_lanes(lanes):
    return lanes


def oracle_batch(ref, mode, buf, off):
    idx = oracle.OracleIndex(ref, mode)
    _, L, P = idx.query_batch(buf, off, threads=os.cpu_count() or 1)
    return L, P


@pytest.mark.parametrize("lanes", LANES)
def test_golden_vectors(tmp_path, lanes, row_width):
    for ci, case in enumerate(json.load(open(GOLDEN))["cases"]):
        ref = build_index(str(tmp_path / f"c{ci}"), text=case["text"].encode("latin1"))
        pml, ms = sp.pml_t(ref), sp.ms_t(ref)
        assert pml.get_bwt_stats() == (case["n"], case["r"])
        reads = [rd["read"].encode("latin1") for rd in case["reads"]]
        buf, off = sp.pack_reads(reads)
        L, _, _ = pml.query_host(buf, off, lanes=lanes)
        Lm, Pm, _ = ms.query_host(buf, off, lanes=ms_lanes(lanes))
        for i, rd in enumerate(case["reads"]):
            a, b = int(off[i]), int(off[i + 1])
            assert L[a:b].tolist() == rd["pml"]
            assert [str(x) for x in Pm[a:b].tolist()] == rd["ms_pointers"]
            assert Lm[a:b].tolist() == rd["ms_lengths"]
            # reference-shaped single-read call
            assert pml.matching_statistics(reads[i], lanes=lanes).tolist() == rd["pml"]


@pytest.mark.parametrize("lanes", LANES)
def test_small_index_all_reads(small_index, lanes, row_width):
    buf, off = load_bin_reads(small_index["reads_bin"])
    Lo, _ = oracle_batch(small_index["ref"], "pml", buf, off)
    L, _, _ = sp.pml_t(small_index["ref"]).query_host(buf, off, lanes=lanes)
    assert np.array_equal(L.astype(np.uint64), Lo)
    Lo, Po = oracle_batch(small_index["ref"], "ms", buf, off)
    L, P, _ = sp.ms_t(small_index["ref"]).query_host(buf, off, lanes=ms_lanes(lanes))
    assert np.array_equal(P, Po)
    assert np.array_equal(L.astype(np.uint64), Lo)


def test_edge_cases(small_index, row_width):
    ref = small_index["ref"]
    pml, ms = sp.pml_t(ref), sp.ms_t(ref)
    # empty batch
    L, _, _ = pml.query_host(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert L.size == 0
    # single bases, absent characters, bytes >= 0x80, NUL, the terminator byte, empty reads inside a batch
    reads = [b"A", b"N", b"\xe9", b"\x00", b"\x01", b"", b"ACGT" * 30, b"", b"NNNNACGTNNNN", b"ACG\xffT\x80ACGTACGTA", b"Y" * 40,
             b"ACGTTGCA" * 200 + b"N" + b"TTGACCA" * 150]
    buf, off = sp.pack_reads(reads)
    po, mo = oracle.OracleIndex(ref, "pml"), oracle.OracleIndex(ref, "ms")
    for lanes in LANES:
        L, _, _ = pml.query_host(buf, off, lanes=lanes)
        Lm, Pm, _ = ms.query_host(buf, off, lanes=ms_lanes(lanes))
        for i, r in enumerate(reads):
            a, b = int(off[i]), int(off[i + 1])
            if not r:
                continue
            assert L[a:b].tolist() == po.matching_statistics(r).tolist(), (lanes, r[:20])
            lo, p_o = mo.matching_statistics(r)
            assert Pm[a:b].tolist() == p_o.tolist() and Lm[a:b].tolist() == lo.tolist(), (lanes, r[:20])


def test_toupper_flag(small_index):
    buf, off = load_bin_reads(small_index["reads_bin"])
    pml = sp.pml_t(small_index["ref"])
    lower = np.where((buf >= 65) & (buf <= 90), buf + 32, buf).astype(np.uint8)
    L0, _, _ = pml.query_host(buf, off)
    L1, _, _ = pml.query_host(lower, off, toupper=True)
    L2, _, _ = pml.query_host(lower, off, toupper=False)
    assert np.array_equal(L0, L1) and not np.array_equal(L0, L2)


def test_window_overflow_fallback(tmp_path, row_width):
    """Long BWT runs whose LF image spans hundreds of runs: exercises locate_run's binary search."""
    rng = random.Random(5)
    w = "ACGTACGTTTGACAGGATTACA"
    text = "".join(rng.choice("ACT") + "G" + w + "".join(rng.choice("ACGT") for _ in range(rng.randint(3, 9))) for _ in range(600))
    ref = build_index(str(tmp_path), text=text.encode())
    reads = []
    for _ in range(200):
        s = rng.randint(0, len(text) - 80)
        reads.append(text[s:s + rng.randint(30, 80)].encode())
    reads += [("C" + "G" + w).encode(), ("TG" + w + "A").encode(), (w * 3).encode()]
    buf, off = sp.pack_reads(reads)
    Lo, _ = oracle_batch(ref, "pml", buf, off)
    Lmo, Pmo = oracle_batch(ref, "ms", buf, off)
    for lanes in LANES:
        L, _, _ = sp.pml_t(ref).query_host(buf, off, lanes=lanes)
        assert np.array_equal(L.astype(np.uint64), Lo)
        Lm, Pm, _ = sp.ms_t(ref).query_host(buf, off, lanes=ms_lanes(lanes))
        assert np.array_equal(Pm, Pmo) and np.array_equal(Lm.astype(np.uint64), Lmo)


@pytest.mark.parametrize("mode", ["pml", "ms"])
def test_classifier_stats(small_index, mode, row_width):
    buf, off = load_bin_reads(small_index["reads_bin"])
    idx = sp.pml_t(small_index["ref"]) if mode == "pml" else sp.ms_t(small_index["ref"])
    for bw, thr in ((50, 7), (150, 13), (400, 3)):
        for lanes in [1, 8]:
            L, _, S = idx.query_host(buf, off, bin_width=bw, max_value_thr=thr, lanes=lanes)
            for r in range(len(off) - 1):
                found, summ, above, below, nb = naive.classify_bins(L[int(off[r]):int(off[r + 1])].tolist(), bw, thr)
                assert (int(S[r]["sum_max"]), int(S[r]["above"]), int(S[r]["below"])) == (summ, above, below)


def test_multi_chunk_pipeline_and_ragged_lengths(small_index):
    """> 32 Mi bases forces several pipelined sub-batches; expected output tiles the oracle's."""
    buf, off = load_bin_reads(small_index["reads_bin"])
    Lo, _ = oracle_batch(small_index["ref"], "pml", buf, off)
    reps = (40 << 20) // len(buf) + 1
    big = np.tile(buf, reps)
    lens = np.diff(off)
    big_off = np.concatenate([[0], np.cumsum(np.tile(lens, reps))]).astype(np.uint64)
    L, _, _ = sp.pml_t(small_index["ref"]).query_host(big, big_off)
    assert L.size == len(buf) * reps
    assert np.array_equal(L.reshape(reps, -1).astype(np.uint64), np.broadcast_to(Lo, (reps, len(buf))))


def test_device_api_with_torch_tensors(small_index):
    import torch
    buf, off = load_bin_reads(small_index["reads_bin"])
    Lo, _ = oracle_batch(small_index["ref"], "pml", buf, off)
    idx = sp.pml_t(small_index["ref"])
    dev = torch.device("cuda", 0)
    d_reads = torch.zeros(len(buf) + 512, dtype=torch.uint8, device=dev)
    d_reads[: len(buf)] = torch.from_numpy(buf).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_len = torch.zeros(len(buf), dtype=torch.int32, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        idx.query_device(d_reads, d_off, len(off) - 1, d_len, stream=s.cuda_stream)
    s.synchronize()
    assert np.array_equal(d_len.cpu().numpy().astype(np.uint64), Lo)
    # narrow device output + device-side clamp counter
    d_len8 = torch.zeros(len(buf), dtype=torch.uint8, device=dev)
    d_sat = torch.zeros(1, dtype=torch.int64, device=dev)
    idx.query_device(d_reads, d_off, len(off) - 1, d_len8, d_saturated=d_sat)
    torch.cuda.synchronize()
    assert int(d_sat.item()) == int(np.count_nonzero(Lo > 255))
    assert np.array_equal(d_len8.cpu().numpy().astype(np.uint64), np.minimum(Lo, 255))
    with pytest.raises(sp.SpumoniError):  # alignment contract
        idx.query_device(d_reads[16:], d_off, len(off) - 1, d_len)
    assert sp.kernel_launches() > 0


def test_attach_device_from_broadcast_buffer(small_index):
    import torch
    buf, off = load_bin_reads(small_index["reads_bin"])
    base = sp.Index.open(small_index["ref"], sp.MODE_PML)
    blob = np.array(base.flat(), copy=True)
    dev_blob = torch.from_numpy(blob).to("cuda:0")
    idx = sp.Index.from_header(blob[: sp.flat_header_bytes()]).attach_device(0, dev_blob)
    L, _, _ = idx.query_host(buf, off)
    L2, _, _ = base.to_device(0).query_host(buf, off)
    assert np.array_equal(L, L2)


@pytest.mark.parametrize("flags", [["-P"], ["-P", "-c"], ["-M"], ["-M", "-c", "-w", "100"], ["-P", "-d"], ["-M", "-d", "-c"]])
def test_cli_outputs_byte_identical_to_oracle(small_index, tmp_path, flags):
    ref = small_index["ref"]
    for ext in (".pmlnulldb", ".msnulldb"):
        stats = (np.arange(1, 200, dtype=np.uint64) * 7) % 11
        stats.tofile(str(tmp_path / "stats.bin"))
        subprocess.check_call([MKINDEX, "nulldb", "--stats", str(tmp_path / "stats.bin"), "--out", ref + ext])
    qa, qb = str(tmp_path / "a.fa"), str(tmp_path / "b.fa")
    body = open(small_index["reads_fa"]).read()
    body = body.replace(">read_3\n", ">read_3 with description\n").replace(">read_5\n", ">read_5\tx\n")
    body = body.replace("ACGT", "acgt", 5)
    open(qa, "w").write(body)
    open(qb, "w").write(body)
    ms = "-M" in flags
    bw = int(flags[flags.index("-w") + 1]) if "-w" in flags else 150
    oracle.run(ref, qa, "ms" if ms else "pml", report="-c" in flags, bin_width=bw, threads=1, use_doc="-d" in flags)
    r = subprocess.run([sp.CLI_PATH, "run", "-r", ref[:-3], "-p", qb, "-n"] + flags, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if flags[0] == "-P" and "-d" not in flags:  # this path formats on the device; the host formatter must give the same bytes
        qc = str(tmp_path / "c.fa")
        open(qc, "w").write(body)
        r2 = subprocess.run([sp.CLI_PATH, "run", "-r", ref[:-3], "-p", qc, "-n", "-t", "3"] + flags, capture_output=True, text=True,
                            env=dict(os.environ, SPUMONI_B200_HOST_FORMAT="1"))
        assert r2.returncode == 0, r2.stderr
        assert open(qc + ".pseudo_lengths", "rb").read() == open(qb + ".pseudo_lengths", "rb").read()
    exts = [".lengths", ".pointers"] if ms else [".pseudo_lengths"]
    if "-c" in flags:
        exts.append(".report")
    if "-d" in flags:
        exts.append(".doc_numbers")
    for e in exts:
        assert open(qa + e, "rb").read() == open(qb + e, "rb").read(), e
    assert "finished processing 300 reads" in r.stderr


REFERENCE_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_run")


@pytest.mark.parametrize("mode,flags", [("P", ["-P", "-c"]), ("Pd", ["-P", "-d"]), ("M", ["-M", "-c"]), ("Md", ["-M", "-d", "-c"]),
                                        ("Pg", ["-P", "-g"]), ("Mg", ["-M", "-g"])])
def test_cli_outputs_byte_identical_to_reference(tmp_path, row_width, mode, flags):
    """The product CLI against files written by the REFERENCE's own code (tests/golden/reference_run, produced by
    oracle/_ref/spumoni_ref: index, document array and null databases from the reference's build functions, outputs from
    its `spumoni run`). No oracle in between."""
    import shutil
    d = str(tmp_path / "g")
    shutil.copytree(REFERENCE_GOLDEN, d, ignore=shutil.ignore_patterns("expected"))
    pattern = "general.txt" if "-g" in flags else "reads.fa"  # -g: reads separated by 0x01, used verbatim
    r = subprocess.run([sp.CLI_PATH, "run", "-r", os.path.join(d, "idx"), "-p", os.path.join(d, pattern), "-n"] + flags,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = os.path.join(REFERENCE_GOLDEN, "expected", mode)
    names = sorted(os.listdir(want))
    assert names
    for f in names:
        assert open(os.path.join(d, f), "rb").read() == open(os.path.join(want, f), "rb").read(), f


def test_abi_values_equal_reference_golden(row_width):
    """C ABI (host entry points, PML + MS + document numbers) against the integers in the reference-written files."""
    def parse(path):
        return [np.array(line.split(), dtype=np.uint64) for line in open(path) if not line.startswith(">")]
    reads = [line.strip().encode() for line in open(os.path.join(REFERENCE_GOLDEN, "reads.fa")) if not line.startswith(">")]
    # the last record of the file spans three lines
    reads = reads[:-3] + [b"".join(reads[-3:])]
    buf = np.frombuffer(b"".join(reads), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    ref = os.path.join(REFERENCE_GOLDEN, "idx.fa")
    pml = sp.pml_t(ref, device=0, use_doc=True)
    L, _, _, D = pml.query_host_doc(buf, off, toupper=True)
    want_l = np.concatenate(parse(os.path.join(REFERENCE_GOLDEN, "expected", "Pd", "reads.fa.pseudo_lengths")))
    want_d = np.concatenate(parse(os.path.join(REFERENCE_GOLDEN, "expected", "Pd", "reads.fa.doc_numbers")))
    assert np.array_equal(L.astype(np.uint64), want_l) and np.array_equal(D.astype(np.uint64), want_d)
    ms = sp.ms_t(ref, device=0, use_doc=True)
    L, P, _, D = ms.query_host_doc(buf, off, toupper=True)
    want_l = np.concatenate(parse(os.path.join(REFERENCE_GOLDEN, "expected", "Md", "reads.fa.lengths")))
    want_p = np.concatenate(parse(os.path.join(REFERENCE_GOLDEN, "expected", "Md", "reads.fa.pointers")))
    want_d = np.concatenate(parse(os.path.join(REFERENCE_GOLDEN, "expected", "Md", "reads.fa.doc_numbers")))
    assert np.array_equal(L.astype(np.uint64), want_l) and np.array_equal(P.astype(np.uint64), want_p)
    assert np.array_equal(D.astype(np.uint64), want_d)


def test_cli_empty_read_aborts_like_reference(small_index, tmp_path):
    q = str(tmp_path / "q.fa")
    open(q, "w").write(">r_a\nACGT\n>r_b\n>r_c\nACGT\n")
    r = subprocess.run([sp.CLI_PATH, "run", "-r", small_index["ref"][:-3], "-p", q, "-P", "-n"], capture_output=True, text=True)
    assert r.returncode == 1 and "was empty after digestion" in r.stderr


def test_cli_general_text(small_index, tmp_path):
    q = str(tmp_path / "q.txt")
    open(q, "wb").write(b"ACGTACGTAGGT\x01acgtNNACGT\x01TTTT\x01tail-without-separator")
    r = subprocess.run([sp.CLI_PATH, "run", "-r", small_index["ref"][:-3], "-p", q, "-P", "-n", "-g"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    po = oracle.OracleIndex(small_index["ref"], "pml")
    exp = b""
    for i, rd in enumerate([b"ACGTACGTAGGT", b"acgtNNACGT", b"TTTT"]):
        exp += b">read_%d\n" % i + b"".join(b"%d " % v for v in po.matching_statistics(rd).tolist()) + b"\n"
    assert open(q + ".pseudo_lengths", "rb").read() == exp


def test_medium_scale_properties(tmp_path, row_width):
    """2 Mb pangenome, 20k x 500 bp reads: full comparison with the (multi-threaded) oracle plus
    size-independent properties (PML <= MS, X[i] <= X[i+1]+1, pointers in range when no char is absent)."""
    d = str(tmp_path)
    from conftest import mkindex
    mkindex("genome", "--kind", "pangenome", "--len", 2000000, "--haps", 10, "--snp", 0.002, "--indel", 0.0002, "--seed", 21, "--out", d + "/raw.fa")
    ref = build_index(d, fasta=d + "/raw.fa", revcomp=True)
    mkindex("reads", "--ref", ref, "--n", 20000, "--len", 500, "--sub", 0.06, "--ins", 0.02, "--del", 0.02, "--neg-frac", 0.1, "--seed", 22, "--bin", d + "/r.bin")
    buf, off = load_bin_reads(d + "/r.bin")
    Lp, _, _ = sp.pml_t(ref).query_host(buf, off)
    Lm, Pm, _ = sp.ms_t(ref).query_host(buf, off)
    n = sp.Index.open(ref, 0).get_bwt_stats()[0]
    assert np.all(Lp <= Lm)
    inner = np.ones(len(buf), dtype=bool)
    inner[(off[1:] - 1).astype(np.int64)] = False  # last position of each read has no successor
    assert np.all(Lm[:-1][inner[:-1]] <= Lm[1:][inner[:-1]] + 1)
    assert np.all(Lp[:-1][inner[:-1]] <= Lp[1:][inner[:-1]] + 1)
    assert np.all(Pm < n)
    Lo, _ = oracle_batch(ref, "pml", buf, off)
    assert np.array_equal(Lp.astype(np.uint64), Lo)
    Lmo, Pmo = oracle_batch(ref, "ms", buf, off)
    assert np.array_equal(Pm, Pmo) and np.array_equal(Lm.astype(np.uint64), Lmo)


def test_move_rows_saturation_fallbacks_on_gpu(small_index, tmp_path, row_width):
    """Same inputs through the test-only library build whose move-row fields are tiny (3-bit len/delta,
    thresholds <= 4, jump distance < 3): every fall-back of the thread-per-read kernels (slow rows, generic
    step from un-normalised states, far jumps) runs and must stay bit-exact, PML and MS."""
    import sys, textwrap
    buf, off = load_bin_reads(small_index["reads_bin"])
    Lo, _ = oracle_batch(small_index["ref"], "pml", buf, off)
    Lmo, Pmo = oracle_batch(small_index["ref"], "ms", buf, off)
    out = str(tmp_path / "out.npz")
    code = textwrap.dedent(f"""
        import numpy as np, spumoni_b200 as sp
        buf = np.fromfile({small_index["reads_bin"]!r}, dtype=np.uint8); off = np.fromfile({small_index["reads_bin"] + ".off"!r}, dtype=np.uint64)
        L, _, S = sp.pml_t({small_index["ref"]!r}).query_host(buf, off, lanes=1, bin_width=150, max_value_thr=9)
        Lm, Pm, _ = sp.ms_t({small_index["ref"]!r}).query_host(buf, off, lanes=1)
        np.savez({out!r}, L=L, Lm=Lm, Pm=Pm)
    """)
    env = dict(os.environ, SPUMONI_B200_LIB=sp.SATTEST_LIB_PATH, SPUMONI_B200_ROWS=row_width,
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    subprocess.check_call([sys.executable, "-c", code], env=env)
    got = np.load(out)
    assert np.array_equal(got["L"].astype(np.uint64), Lo)
    assert np.array_equal(got["Pm"], Pmo) and np.array_equal(got["Lm"].astype(np.uint64), Lmo)


@pytest.mark.parametrize("mode", ["pml", "ms"])
def test_narrow_length_outputs(small_index, mode, row_width):
    """uint16 / uint8 lengths are bit-identical to uint32 when nothing is clamped; a value that does not fit is clamped and
    counted, never silently wrapped."""
    buf, off = load_bin_reads(small_index["reads_bin"])
    idx = sp.pml_t(small_index["ref"]) if mode == "pml" else sp.ms_t(small_index["ref"])
    L4, P4, S4 = idx.query_host(buf, off, bin_width=150, max_value_thr=9)
    L2, P2, S2 = idx.query_host(buf, off, bin_width=150, max_value_thr=9, length_bytes=2)
    assert L2.dtype == np.uint16 and np.array_equal(L2, L4) and np.array_equal(S2, S4)
    if mode == "ms":
        assert np.array_equal(P2, P4)
    over = int(np.count_nonzero(L4 > 255))
    if over == 0:
        L1, _, S1 = idx.query_host(buf, off, bin_width=150, max_value_thr=9, length_bytes=1)
        assert L1.dtype == np.uint8 and np.array_equal(L1, L4) and np.array_equal(S1, S4)
    # a read that matches the reference exactly for > 255 bases does not fit uint8: MS reports it (SPB_ERR_OVERFLOW + count,
    # never silently wrapped); for PML a clamped array is still lossless (zero / non-zero is kept and the non-zero values count
    # up from the right), so the call succeeds, says how many values it clamped, and spb_pml_widen_lengths rebuilds them --
    # one long exact match in a batch no longer costs a re-run
    text = open(small_index["ref"]).read().split("\n")[1][:600].encode()
    rb, ro = sp.pack_reads([text, b"ACGT"])
    Lw, _, _ = idx.query_host(rb, ro)
    assert int(Lw.max()) > 255
    out1 = np.zeros(len(rb), dtype=np.uint8)
    if mode == "ms":
        with pytest.raises(sp.SpumoniError) as ei:
            idx.query_host(rb, ro, length_bytes=1, out_lengths=out1)
        assert ei.value.code == sp.ERR_OVERFLOW and ei.value.saturated == int(np.count_nonzero(Lw > 255))
    else:
        idx.query_host(rb, ro, length_bytes=1, out_lengths=out1)
        assert int(idx.last_saturated.value) == int(np.count_nonzero(Lw > 255))
        assert np.array_equal(sp.widen_lengths(out1, ro), Lw)
    assert np.array_equal(out1, np.minimum(Lw, 255).astype(np.uint8))  # clamped, otherwise complete
    # the cooperative kernels only write uint32
    with pytest.raises(sp.SpumoniError):
        idx.query_host(rb, ro, lanes=8, length_bytes=2)


def test_unaligned_read_boundaries_narrow(small_index, row_width):
    """Reads of every length 1..70 back to back: 16-value output words straddle read boundaries."""
    rng = random.Random(3)
    text = open(small_index["ref"]).read().split("\n")[1]
    reads = []
    for ln in range(1, 71):
        s0 = rng.randint(0, len(text) - ln)
        r = bytearray(text[s0:s0 + ln].encode())
        if ln > 4 and rng.random() < 0.7:
            r[rng.randrange(ln)] = rng.choice(b"ACGT")
        reads.append(bytes(r))
    buf, off = sp.pack_reads(reads)
    Lo, _ = oracle_batch(small_index["ref"], "pml", buf, off)
    Lmo, Pmo = oracle_batch(small_index["ref"], "ms", buf, off)
    pml, ms = sp.pml_t(small_index["ref"]), sp.ms_t(small_index["ref"])
    for w in (1, 2, 4):
        L, _, _ = pml.query_host(buf, off, length_bytes=w)
        assert np.array_equal(L.astype(np.uint64), Lo), w
        Lm, Pm, _ = ms.query_host(buf, off, length_bytes=w)
        assert np.array_equal(Lm.astype(np.uint64), Lmo) and np.array_equal(Pm, Pmo), w


@pytest.mark.parametrize("mode", ["pml", "ms"])
def test_document_numbers(small_index, mode, row_width):
    """-d: document numbers (compute_ms_pml.cpp:288-340 / :625-682) bit-exact against the oracle, together with the
    lengths / pointers / classifier stats of the same call; both row layouts; the fall-back build as well."""
    import sys, textwrap
    ref = small_index["ref"]
    buf, off = load_bin_reads(small_index["reads_bin"])
    extra = [b"ACGTNNACGT" * 5, b"N", b"ACG\xe9T" * 8, b"", b"T"]
    eb, eo = sp.pack_reads(extra)
    buf2 = np.concatenate([buf, eb])
    off2 = np.concatenate([off, eo[1:] + off[-1]]).astype(np.uint64)
    o = oracle.OracleIndex(ref, mode, use_doc=True)
    Lo, Po, Do = o.query_batch_doc(buf2, off2)
    idx = (sp.pml_t if mode == "pml" else sp.ms_t)(ref, use_doc=True)
    assert idx.num_docs() == 3
    L, P, S, D = idx.query_host_doc(buf2, off2, bin_width=150, max_value_thr=9)
    assert np.array_equal(D.astype(np.uint64), Do) and np.array_equal(L.astype(np.uint64), Lo)
    if mode == "ms":
        assert np.array_equal(P, Po)
    L0, _, S0 = idx.query_host(buf2, off2, bin_width=150, max_value_thr=9)  # the same index still answers plain queries
    assert np.array_equal(L0, L) and np.array_equal(S0, S)
    with pytest.raises(sp.SpumoniError):  # index opened without the document array
        (sp.pml_t if mode == "pml" else sp.ms_t)(ref).query_host_doc(buf2, off2)
    with pytest.raises(sp.SpumoniError):  # cooperative kernels have no document support
        idx.query_host_doc(buf2, off2, lanes=8)
    out = str(tmp_path_for(small_index) / f"doc_{mode}_{row_width}.npy")
    code = textwrap.dedent(f"""
        import numpy as np, spumoni_b200 as sp
        buf = np.load({out + ".buf.npy"!r}); off = np.load({out + ".off.npy"!r})
        idx = (sp.pml_t if {mode!r} == "pml" else sp.ms_t)({ref!r}, use_doc=True)
        np.save({out!r}, idx.query_host_doc(buf, off)[3])
    """)
    np.save(out + ".buf.npy", buf2); np.save(out + ".off.npy", off2)
    env = dict(os.environ, SPUMONI_B200_LIB=sp.SATTEST_LIB_PATH, SPUMONI_B200_ROWS=row_width,
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    subprocess.check_call([sys.executable, "-c", code], env=env)
    assert np.array_equal(np.load(out).astype(np.uint64), Do)


def tmp_path_for(small_index):
    import pathlib
    return pathlib.Path(small_index["dir"])


def test_ms_sample_layouts_on_gpu(small_index, monkeypatch):
    """Both homes of the per-row SA samples (inline 32-bit in 16-byte row slots / separate table) give the oracle's pointers."""
    buf, off = load_bin_reads(small_index["reads_bin"])
    Lo, Po = oracle_batch(small_index["ref"], "ms", buf, off)
    monkeypatch.setenv("SPUMONI_B200_ROWS", "8")
    for layout in ("inline", "wide", "split32"):
        monkeypatch.setenv("SPUMONI_B200_MSAMP", layout)
        L, P, _ = sp.ms_t(small_index["ref"]).query_host(buf, off)
        assert np.array_equal(P, Po) and np.array_equal(L.astype(np.uint64), Lo), layout


def test_long_and_mixed_length_reads(small_index, row_width):
    """BASELINE config 5's shape in miniature: 150 bp .. 50 kb reads in one batch (a read is one dependent chain; the
    work counter balances the ragged lengths), uint16 lengths where they fit, classifier stats on every read."""
    rng = random.Random(11)
    text = "".join(l for l in open(small_index["ref"]).read().split("\n") if not l.startswith(">"))
    reads = []
    for ln in (150, 500, 1000, 5000, 10000, 50000, 150, 10000, 33, 65535, 65536):
        parts, tot = [], 0
        while tot < ln:
            a = rng.randint(0, len(text) - 2000)
            seg = bytearray(text[a:a + rng.randint(200, 2000)].encode())
            for i in range(len(seg)):
                if rng.random() < 0.03:
                    seg[i] = rng.choice(b"ACGT")
            parts.append(bytes(seg))
            tot += len(seg)
        reads.append(b"".join(parts)[:ln])
    buf, off = sp.pack_reads(reads)
    Lo, _ = oracle_batch(small_index["ref"], "pml", buf, off)
    Lmo, Pmo = oracle_batch(small_index["ref"], "ms", buf, off)
    pml, ms = sp.pml_t(small_index["ref"]), sp.ms_t(small_index["ref"])
    L, _, S = pml.query_host(buf, off, bin_width=150, max_value_thr=9)
    assert np.array_equal(L.astype(np.uint64), Lo)
    for r in range(len(reads)):
        found, summ, above, below, nb = naive.classify_bins(Lo[int(off[r]):int(off[r + 1])].tolist(), 150, 9)
        assert (int(S[r]["sum_max"]), int(S[r]["above"]), int(S[r]["below"])) == (summ, above, below)
    Lm, Pm, _ = ms.query_host(buf, off)
    assert np.array_equal(Pm, Pmo) and np.array_equal(Lm.astype(np.uint64), Lmo)
    L16, _, _ = pml.query_host(buf, off, length_bytes=2)  # a PML never exceeds the read length; here it stays far below 65536
    assert np.array_equal(L16.astype(np.uint64), Lo)


@pytest.mark.parametrize("mode", ["pml", "ms"])
def test_null_database_generation(small_index, tmp_path, mode):
    """The last stage of `spumoni build` through the GPU path: reversed null reads -> null statistics -> EmpNullDatabase
    (compute_ms_pml.cpp:1410-1486, emp_null_database.cpp:19-102); statistics, summary and the serialised file match the
    oracle / the fixture writer, and the file loads back through EmpNullDatabase::load."""
    buf, off = load_bin_reads(small_index["reads_bin"])
    lower = np.where((buf >= 65) & (buf <= 90), buf + 32, buf).astype(np.uint8)  # null reads are upper-cased first
    o = oracle.OracleIndex(small_index["ref"], mode)
    so, mean_o, perc_o = o.nulldb_generate(lower, off)
    idx = (sp.pml_t if mode == "pml" else sp.ms_t)(small_index["ref"])
    out = str(tmp_path / ("x.pmlnulldb" if mode == "pml" else "x.msnulldb"))
    db, stats = sp.generate_null_database(idx, lower, off, out_path=out, ks_stat_threshold=0.25)
    assert np.array_equal(stats.astype(np.uint64), so)
    assert db.num_values == len(buf) and db.mean_null_stat == mean_o and db.percentile_value == perc_o
    so.tofile(str(tmp_path / "s.bin"))
    subprocess.check_call([MKINDEX, "nulldb", "--stats", str(tmp_path / "s.bin"), "--out", out + ".fixture", "--ks-thr", "0.25"])
    assert open(out, "rb").read() == open(out + ".fixture", "rb").read()
    back = sp.EmpNullDatabase().load(out)
    assert (back.num_values, back.percentile_value, back.mean_null_stat, back.ks_stat_threshold) == (len(buf), perc_o, mean_o, 0.25)


def test_baseline_config1_plumbing(tmp_path):
    """BASELINE.json configs[0]: PML on a 1 Mb single-FASTA index (i.i.d. ACGT, + reverse complement, n = 2 000 001),
    10 000 synthetic 150 bp reads (half mutated substrings, half random), `spumoni run -P -n -c`: every output file is
    byte-for-byte the reference CPU path's (oracle CLI, 1 thread, so that the output order is defined)."""
    d = str(tmp_path)
    from conftest import mkindex
    mkindex("genome", "--kind", "iid", "--len", 1000000, "--seed", 1, "--out", d + "/raw.fa")
    ref = build_index(d, fasta=d + "/raw.fa", revcomp=True)
    n, r = sp.Index.open(ref, 0).get_bwt_stats()
    assert n == 2000001
    mkindex("reads", "--ref", ref, "--n", 10000, "--len", 150, "--sub", 0.01, "--neg-frac", 0.5, "--seed", 2, "--out", d + "/a.fa")
    import shutil
    shutil.copy(d + "/a.fa", d + "/b.fa")
    stats = (np.arange(1, 200, dtype=np.uint64) * 7) % 11
    stats.tofile(d + "/stats.bin")
    subprocess.check_call([MKINDEX, "nulldb", "--stats", d + "/stats.bin", "--out", ref + ".pmlnulldb"])
    nr, _ = oracle.run(ref, d + "/a.fa", "pml", report=True, threads=1)
    assert nr == 10000
    res = subprocess.run([sp.CLI_PATH, "run", "-r", ref[:-3], "-p", d + "/b.fa", "-P", "-n", "-c", "-t", "4"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    for e in (".pseudo_lengths", ".report"):
        assert open(d + "/a.fa" + e, "rb").read() == open(d + "/b.fa" + e, "rb").read(), e


@pytest.mark.parametrize("lanes", [1, 8, 32])
def test_positions_beyond_2_32_on_gpu(huge_periodic_index, lanes, row_width):
    """BASELINE configs 4-5 territory: BWT positions > 2^32 (n = 4 320 000 001), all kernels, bit-exact vs the oracle."""
    ref, reads = huge_periodic_index["ref"], huge_periodic_index["reads"]
    buf, off = sp.pack_reads(reads)
    Lo, _ = oracle_batch(ref, "pml", buf, off)
    L, _, S = sp.pml_t(ref).query_host(buf, off, lanes=lanes, bin_width=50, max_value_thr=7)
    assert np.array_equal(L.astype(np.uint64), Lo)
    for r in range(len(reads)):
        found, summ, above, below, nb = naive.classify_bins(Lo[int(off[r]):int(off[r + 1])].tolist(), 50, 7)
        assert (int(S[r]["sum_max"]), int(S[r]["above"]), int(S[r]["below"])) == (summ, above, below)


def test_on_device_text_formatting(small_index, monkeypatch, row_width):
    """K5: spb_pml_query_host_text returns the bytes of the .pseudo_lengths file (compute_ms_pml.cpp:1003-1010): ids with the
    trailing-delimiter quirk, empty ids, 1-base reads, a read above 65535 bases (uint32 values in that sub-batch), many
    sub-batches in flight (1 Mi-base chunks), classifier stats alongside, and a too-small buffer reported as such."""
    monkeypatch.setenv("SPUMONI_B200_CHUNK_MB", "1")
    rng = random.Random(5)
    buf, off = load_bin_reads(small_index["reads_bin"])
    reads = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    text = "".join(l for l in open(small_index["ref"]).read().split("\n") if not l.startswith(">"))
    reads += [b"A", b"N", text[100:70100].encode(), b"ACGT" * 300] + [text[a:a + 5000].encode() for a in range(0, 600000, 4000)][:400]
    ids = [b"read_%d" % i for i in range(len(reads))]
    ids[3], ids[4], ids[5] = b"r3 ", b"x\t", b""
    rb, ro = sp.pack_reads(reads)
    Lo, _ = oracle_batch(small_index["ref"], "pml", rb, ro)
    want = b"".join(b">" + ids[i] + b"\n" + b"".join(b"%d " % v for v in Lo[int(ro[i]):int(ro[i + 1])].tolist()) + b"\n" for i in range(len(reads)))
    idx = sp.pml_t(small_index["ref"])
    got, S = idx.query_host_text(rb, ro, ids, bin_width=150, max_value_thr=9)
    assert got == want
    _, _, S0 = idx.query_host(rb, ro, bin_width=150, max_value_thr=9)
    assert np.array_equal(S, S0)
    with pytest.raises(sp.SpumoniError) as ei:
        idx.query_host_text(rb, ro, ids, capacity=len(want) - 1)
    assert ei.value.code == 6 and 0 < idx.last_text_bytes <= len(want)
    assert idx.query_host_text(np.zeros(0, np.uint8), np.zeros(1, np.uint64), [])[0] == b""


def test_ms_pointers32(small_index, tmp_path, row_width):
    """SPB_QUERY_POINTERS32: uint32 MS pointers equal the uint64 ones when every value fits; a read with a character absent
    from the index makes the reference's pointer wrap around (App. C-4), which is reported, never truncated silently."""
    buf, off = load_bin_reads(small_index["reads_bin"])
    ms = sp.ms_t(small_index["ref"])
    L8, P8, S8 = ms.query_host(buf, off, bin_width=150, max_value_thr=9)
    clean = int(P8.max()) < (1 << 32) - 1
    if clean:
        L4, P4, S4 = ms.query_host(buf, off, bin_width=150, max_value_thr=9, pointer_bytes=4, length_bytes=2)
        assert P4.dtype == np.uint32 and np.array_equal(P4.astype(np.uint64), P8) and np.array_equal(L4, L8) and np.array_equal(S4, S8)
    # in CCGATNAGATT the N resets the sample to 0 and the matches to its left wrap to 2^64 - 1 ...
    ms = sp.ms_t(build_index(str(tmp_path), text=b"GATTACAGATTACCGATTAGA"))
    rb, ro = sp.pack_reads([b"CCGATNAGATT", b"TTACAGGATTAG"])
    _, Pw, _ = ms.query_host(rb, ro)
    assert Pw[:5].tolist() == [12, 13, 18446744073709551613, 18446744073709551614, 18446744073709551615]
    with pytest.raises(sp.SpumoniError) as ei:
        ms.query_host(rb, ro, pointer_bytes=4)
    assert ei.value.code == sp.ERR_OVERFLOW and ei.value.saturated == 3
    with pytest.raises(sp.SpumoniError):
        ms.query_host(rb, ro, pointer_bytes=4, lanes=8)


def test_bit_output_and_packed_input(small_index, row_width):
    """The compact transports of the PML call: one bit per base out (SPB_LENGTH_BITS), 2 bits per base in
    (spb_pml_query_host_packed, non-ACGT bytes through the exception list) -- both lossless, alone and together, with the
    classifier statistics unchanged; several sub-batches (edge blocks of the bit map are merged on the host)."""
    buf, off = load_bin_reads(small_index["reads_bin"])
    buf = buf.copy()
    buf[::97] = ord("N")
    buf[5::211] = ord("a")  # lower case: verbatim without toupper, upper-cased with it
    idx = sp.pml_t(small_index["ref"])
    for up in (False, True):
        L4, _, S4 = idx.query_host(buf, off, bin_width=150, max_value_thr=9, toupper=up)
        bits, _, Sb = idx.query_host(buf, off, bin_width=150, max_value_thr=9, toupper=up, length_bytes=sp.LENGTH_BITS)
        assert np.array_equal(sp.expand_bits(bits, off), L4) and np.array_equal(Sb, S4)
        packed = sp.pack_reads_2bit(buf, toupper=up)
        assert packed[1].size > 0
        Lp, _, Sp = idx.query_host(None, off, bin_width=150, max_value_thr=9, toupper=up, packed=packed)
        assert np.array_equal(Lp, L4) and np.array_equal(Sp, S4)
        bp, _, _ = idx.query_host(None, off, toupper=up, packed=packed, length_bytes=sp.LENGTH_BITS)
        assert np.array_equal(bp, bits)


def test_bit_output_many_sub_batches(small_index, monkeypatch):
    """Sub-batches of a few reads each: every sub-batch boundary falls inside a 32-byte block of the bit map and inside a
    packed word; short and empty reads share words."""
    monkeypatch.setenv("SPUMONI_B200_TEST_CHUNK_BASES", "700")  # sub-batches close after 700 bases
    monkeypatch.setenv("SPUMONI_B200_CHUNK_READS", "3")
    buf, off = load_bin_reads(small_index["reads_bin"])
    reads = [bytes(buf[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    reads = reads[:40] + [b"", b"A", b"", b"ACG", b"T" * 31, b"G" * 33] + reads[40:120]
    rb, ro = sp.pack_reads(reads)
    idx = sp.pml_t(small_index["ref"])
    L4, _, _ = idx.query_host(rb, ro)
    Lo, _ = oracle_batch(small_index["ref"], "pml", rb, ro)
    assert np.array_equal(L4.astype(np.uint64), Lo)
    bits, _, _ = idx.query_host(rb, ro, length_bytes=sp.LENGTH_BITS)
    assert np.array_equal(sp.expand_bits(bits, ro), L4)
    bp, _, _ = idx.query_host(None, ro, packed=sp.pack_reads_2bit(rb), length_bytes=sp.LENGTH_BITS)
    assert np.array_equal(bp, bits)


def test_bits_device_api(small_index):
    import torch
    buf, off = load_bin_reads(small_index["reads_bin"])
    Lo, _ = oracle_batch(small_index["ref"], "pml", buf, off)
    idx = sp.pml_t(small_index["ref"])
    dev = torch.device("cuda", 0)
    d_reads = torch.zeros(len(buf) + 512, dtype=torch.uint8, device=dev)
    d_reads[: len(buf)] = torch.from_numpy(buf).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_bits = torch.zeros(sp.bits_bytes(len(buf)) // 8, dtype=torch.int64, device=dev)
    idx.query_device(d_reads, d_off, len(off) - 1, d_bits)
    torch.cuda.synchronize()
    assert np.array_equal(sp.expand_bits(d_bits.cpu().numpy().view(np.uint64), off).astype(np.uint64), Lo)


def test_mixed_lengths_longest_first(small_index):
    """A sub-batch with very different read lengths is walked longest read first (order array); outputs stay in input order."""
    text = "".join(line.strip() for line in open(small_index["ref"]) if not line.startswith(">")).encode()
    rng = np.random.default_rng(5)
    reads = []
    for i in range(300):
        ln = int(rng.choice([30, 150, 1000, 6000]))
        s0 = int(rng.integers(0, len(text) - ln))
        r = bytearray(text[s0:s0 + ln])
        for j in range(0, ln, 37):
            r[j] = ord("ACGT"[int(rng.integers(0, 4))])
        reads.append(bytes(r))
    rb, ro = sp.pack_reads(reads)
    assert sp.plan_order(ro) is not None
    Lo, _ = oracle_batch(small_index["ref"], "pml", rb, ro)
    for mode_idx in (sp.pml_t(small_index["ref"]),):
        L, _, S = mode_idx.query_host(rb, ro, bin_width=150, max_value_thr=9)
        assert np.array_equal(L.astype(np.uint64), Lo)
    ms = sp.ms_t(small_index["ref"])
    Lm, Pm, _ = ms.query_host(rb, ro)
    Lmo, Pmo = oracle_batch(small_index["ref"], "ms", rb, ro)
    assert np.array_equal(Pm, Pmo) and np.array_equal(Lm.astype(np.uint64), Lmo)


@pytest.mark.parametrize("method", ["nccl", "peer"])
def test_replicated_index_two_gpus(small_index, method, monkeypatch):
    """spb_index_replicate: the flat image is broadcast to a second GPU (ncclBroadcast, or the peer-to-peer copy tree used when
    NCCL cannot be loaded) and host queries are sharded by bases over both; results equal the single-GPU ones (lengths,
    statistics, bit map with its shared edge blocks)."""
    if sp.device_count() < 2:
        pytest.skip("needs two GPUs")
    if method == "peer":
        monkeypatch.setenv("SPUMONI_B200_NO_NCCL", "1")
    buf, off = load_bin_reads(small_index["reads_bin"])
    one = sp.pml_t(small_index["ref"])
    two = sp.Index.open(small_index["ref"], sp.MODE_PML).replicate([0, 1])
    assert two.device_count() == 2
    assert sp.load_library().spb_replicate_method().decode() == ("cudaMemcpyPeer tree" if method == "peer" else "ncclBroadcast")
    L1, _, S1 = one.query_host(buf, off, bin_width=150, max_value_thr=9)
    L2, _, S2 = two.query_host(buf, off, bin_width=150, max_value_thr=9)
    assert np.array_equal(L1, L2) and np.array_equal(S1, S2)
    b1, _, _ = one.query_host(buf, off, length_bytes=sp.LENGTH_BITS)
    b2, _, _ = two.query_host(buf, off, length_bytes=sp.LENGTH_BITS)
    assert np.array_equal(b1, b2)
    ms2 = sp.Index.open(small_index["ref"], sp.MODE_MS).replicate([1, 0])
    Lm, Pm, _ = ms2.query_host(buf, off)
    Lmo, Pmo = oracle_batch(small_index["ref"], "ms", buf, off)
    assert np.array_equal(Pm, Pmo) and np.array_equal(Lm.astype(np.uint64), Lmo)


def _cli_run(env_extra, prefix, reads, flags):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sp.CLI_PATH, "run", "-r", prefix, "-p", reads, "-n"] + flags, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    out = {}
    for ext in (".pseudo_lengths", ".lengths", ".pointers", ".doc_numbers", ".report"):
        if os.path.exists(reads + ext):
            out[ext] = open(reads + ext, "rb").read()
            os.remove(reads + ext)
    return out, r.stderr


@pytest.fixture(scope="module")
def many_reads(small_index, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("many"))
    path = os.path.join(d, "many.fa")
    subprocess.check_call([MKINDEX, "reads", "--ref", small_index["ref"], "--n", "30000", "--len", "60", "--len-max", "400", "--sub", "0.05",
                           "--neg-frac", "0.3", "--seed", "11", "--out", path], stdout=subprocess.DEVNULL)
    with open(path, "a") as f:  # multi-line records, CRLF, ids with descriptions: the window parser's special cases
        f.write(">ml one two\nACGTACGTAC\nGGTTAACC\n\n>crlf\r\nACGTTGCA\r\n>last_no_newline\nACGTAC")
    for ext in (".pmlnulldb", ".msnulldb"):  # -c needs the null database of the index
        if not os.path.exists(small_index["ref"] + ext):
            ((np.arange(1, 200, dtype=np.uint64) * 7) % 11).tofile(os.path.join(d, "stats.bin"))
            subprocess.check_call([MKINDEX, "nulldb", "--stats", os.path.join(d, "stats.bin"), "--out", small_index["ref"] + ext])
    return path


def test_cli_fast_path_equals_sequential_path(small_index, many_reads):
    """`spumoni run -P`: the fast path (memory-mapped FASTA cut into windows, parallel parse into pinned buffers, text
    formatted on the device, parallel writes) gives the bytes of the sequential path -- itself byte-identical to the
    reference's files (test_cli_outputs_byte_identical_to_reference) -- for every window size and write mode."""
    prefix = small_index["ref"][:-3]
    want, _ = _cli_run({"SPUMONI_B200_SEQUENTIAL": "1"}, prefix, many_reads, ["-P", "-c"])
    assert len(want[".pseudo_lengths"]) > 4_000_000 and ".report" in want
    for env in ({}, {"SPUMONI_B200_BATCH_MB": "1"}, {"SPUMONI_B200_BATCH_MB": "1", "SPUMONI_B200_WRITE": "mmap"},
                {"SPUMONI_B200_BATCH_MB": "2", "SPUMONI_B200_WRITE": "pwrite"}):
        for t in ("1", "5"):
            got, err = _cli_run(env, prefix, many_reads, ["-P", "-c", "-t", t])
            assert got == want, (env, t)
            assert "finished processing 30003 reads" in err
    # MS (lengths + pointers formatted on the device) and document numbers, same comparison
    for flags in (["-M", "-c"], ["-P", "-d"], ["-M", "-d"]):
        want, _ = _cli_run({"SPUMONI_B200_SEQUENTIAL": "1"}, prefix, many_reads, flags)
        assert (".pointers" in want) == ("-M" in flags) and (".doc_numbers" in want) == ("-d" in flags)
        for env in ({"SPUMONI_B200_BATCH_MB": "1"}, {"SPUMONI_B200_WRITE": "pwrite"}):
            got, _ = _cli_run(env, prefix, many_reads, flags + ["-t", "3"])
            assert got == want, (flags, env)


def test_cli_fastq_fast_path_equals_sequential_path(small_index, many_reads, tmp_path):
    """FASTQ on the fast path: the line index, the replay of BatchLoader's batch formation (which tail records it loses)
    and the windows of records give the bytes of the sequential parser's run -- itself record-identical to the reference
    (test_zz_cli_fastq_records_equal_reference) -- with and without a final newline, CRLF, and for -P / -M / -d."""
    prefix = small_index["ref"][:-3]
    recs, cur = [], None
    for line in open(many_reads, newline=""):
        if line.startswith(">"):
            cur = [line[1:].rstrip("\r\n"), ""]
            recs.append(cur)
        else:
            cur[1] += line.strip()
    for name, eol, trail in (("a_fq.fa", "\n", True), ("b_fq.fa", "\n", False), ("c_fq.fa", "\r\n", True)):
        path = str(tmp_path / name)
        with open(path, "w", newline="") as f:
            body = "".join("@%s%s%s%s+%s%s%s" % (i, eol, q, eol, eol, "I" * len(q), eol) for i, q in recs if q)
            f.write(body if trail else body.rstrip("\r\n"))
        for flags in (["-P", "-c"], ["-M", "-d"]) if name == "a_fq.fa" else (["-P"],):
            want, err0 = _cli_run({"SPUMONI_B200_SEQUENTIAL": "1"}, prefix, path, flags)
            for env in ({}, {"SPUMONI_B200_BATCH_MB": "1", "SPUMONI_B200_FASTQ_SEGMENT_MB": "3"},
                        {"SPUMONI_B200_BATCH_MB": "2", "SPUMONI_B200_WRITE": "pwrite", "SPUMONI_B200_FASTQ_SEGMENT_MB": "1"}):
                got, err = _cli_run(env, prefix, path, flags + ["-t", "5"])
                assert got == want, (name, flags, env)
                assert err.split("finished processing")[1].split()[0] == err0.split("finished processing")[1].split()[0]
    # an empty header line skips the rest of its batch, a bad one is fatal: the sequential parser takes over -- from the top of
    # the file, or from the batch boundary where the segment with the odd record begins (segments of 1 MB: mid-file)
    odd = str(tmp_path / "odd_fq.fa")
    with open(odd, "w") as f:
        f.write("".join("@r%d\n%s\n+\n%s\n" % (i, q, "I" * len(q)) for i, (_, q) in enumerate(recs[:3000]) if q))
        f.write("\nACGT\n+\nIIII\n")  # four lines, so the batches after this one stay aligned
        f.write("".join("@s%d\n%s\n+\n%s\n" % (i, q, "I" * len(q)) for i, (_, q) in enumerate(recs[:3000]) if q))
    want, _ = _cli_run({"SPUMONI_B200_SEQUENTIAL": "1"}, prefix, odd, ["-P"])
    got, _ = _cli_run({"SPUMONI_B200_BATCH_MB": "1"}, prefix, odd, ["-P", "-t", "4"])
    mid, _ = _cli_run({"SPUMONI_B200_BATCH_MB": "1", "SPUMONI_B200_FASTQ_SEGMENT_MB": "1"}, prefix, odd, ["-P", "-c", "-t", "4"])
    want_c, _ = _cli_run({"SPUMONI_B200_SEQUENTIAL": "1"}, prefix, odd, ["-P", "-c"])
    assert mid == want_c and os.path.getsize(odd) > (2 << 20)
    n_written = got[".pseudo_lengths"].count(b">")
    assert got == want and 3000 < n_written < 2 * sum(1 for _, q in recs[:3000] if q)


def test_cli_fast_path_hands_over_on_odd_records(small_index, tmp_path):
    """A header without an id in a later window: the windows before it are written by the fast path, then the sequential
    parser takes over at that window and fails exactly like BatchLoader does (same message, exit code 1)."""
    prefix = small_index["ref"][:-3]
    text = "".join(l.strip() for l in open(small_index["ref"]) if not l.startswith(">"))
    path = str(tmp_path / "odd.fa")
    with open(path, "w") as f:
        for i in range(20000):
            f.write(">r%d\n%s\n" % (i, text[i:i + 80]))
        f.write(">\nACGT\n>after\nACGT\n")
    outs = []
    for env in ({"SPUMONI_B200_SEQUENTIAL": "1"}, {"SPUMONI_B200_BATCH_MB": "1"}):
        r = subprocess.run([sp.CLI_PATH, "run", "-r", prefix, "-p", path, "-P", "-n"], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 1 and "header line is missing an id" in r.stderr
        outs.append(r.returncode)
        os.remove(path + ".pseudo_lengths")


def test_cli_two_gpus_byte_identical(small_index, many_reads):
    """`spumoni run` on two GPUs (windows go round robin to one query thread per GPU, written in input order) = one GPU."""
    if sp.device_count() < 2:
        pytest.skip("needs two GPUs")
    prefix = small_index["ref"][:-3]
    one, _ = _cli_run({"SPUMONI_B200_DEVICES": "0", "SPUMONI_B200_BATCH_MB": "1"}, prefix, many_reads, ["-P", "-c", "-t", "4"])
    two, err = _cli_run({"SPUMONI_B200_DEVICES": "0,1", "SPUMONI_B200_BATCH_MB": "1", "SPUMONI_B200_TIMING": "1"}, prefix, many_reads, ["-P", "-c", "-t", "4"])
    assert one == two and "devices: 2" in err
    # the sequential path shards every batch over both GPUs inside the C ABI
    seq, _ = _cli_run({"SPUMONI_B200_DEVICES": "1,0", "SPUMONI_B200_SEQUENTIAL": "1"}, prefix, many_reads, ["-P", "-c", "-t", "4"])
    assert seq == one
    # -t as the number of workers: two threads -> two GPUs
    auto, err = _cli_run({"SPUMONI_B200_TIMING": "1"}, prefix, many_reads, ["-P", "-c", "-t", "2"])
    assert auto == one and "devices: 2" in err


def test_general_text_ms_index_against_reference_binary(tmp_path):
    """`spumoni build -g` indexes the bytes of a general-text file as they are; `run -M -g` on such an index must read the text
    verbatim (line ends included), not as FASTA.  The index is made from the raw bytes (tests/build_emulator.py -> the
    reference's constructors via mkindex from-parts) and the product CLI is compared with the reference's own code."""
    ref_bin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "spumoni_ref")
    if not os.path.exists(ref_bin):
        pytest.skip("oracle/_ref/spumoni_ref is not built")
    from build_emulator import build_parts
    rng = random.Random(8)
    raw = "".join("".join(rng.choice("ACGT") for _ in range(rng.randint(20, 70))) + "\n" for _ in range(40)).encode()
    d = str(tmp_path)
    ref = os.path.join(d, "g.fa")
    open(ref, "wb").write(raw)
    open(ref + ".slp", "wb").write(raw)  # the reference binary's stand-in SLP reads the plain text
    P = build_parts(np.frombuffer(raw + b"\0", dtype=np.uint8))

    def put5(path, cols):
        with open(path, "wb") as f:
            for row in zip(*cols):
                for v in row:
                    f.write(int(v).to_bytes(5, "little"))
    starts = np.concatenate([[0], np.cumsum(P["lens"])])
    open(ref + ".bwt.heads", "wb").write(bytes(P["heads"].astype(np.uint8)))
    put5(ref + ".bwt.len", [P["lens"]])
    put5(ref + ".thr_pos", [P["thr"]])
    put5(ref + ".ssa", [starts[:-1], P["ssa"]])
    put5(ref + ".esa", [starts[1:] - 1, P["esa"]])
    subprocess.check_call([MKINDEX, "from-parts", "--prefix", os.path.join(d, "g"), "--pml", "--ms"], stdout=subprocess.DEVNULL)
    for ext, tool in ((".pmlnulldb", None), (".msnulldb", None)):
        ((np.arange(1, 200, dtype=np.uint64) * 7) % 11).tofile(os.path.join(d, "stats.bin"))
        subprocess.check_call([MKINDEX, "nulldb", "--stats", os.path.join(d, "stats.bin"), "--out", ref + ext])
    q = os.path.join(d, "q.txt")
    lines = raw.split(b"\n")
    open(q, "wb").write(lines[3] + b"\n" + lines[4][:10] + b"\x01" + b"ACGTTTGA\x01" + lines[10][5:] + b"\n" + lines[11] + b"\x01")
    outs = {}
    for name, exe in (("ref", ref_bin), ("cli", sp.CLI_PATH)):
        for mode, exts in (("-M", (".lengths", ".pointers")), ("-P", (".pseudo_lengths",))):
            r = subprocess.run([exe, "run", "-r", os.path.join(d, "g"), "-p", q, mode, "-n", "-g"], capture_output=True, text=True)
            assert r.returncode == 0, (name, mode, r.stderr)
            for e in exts:
                outs[(name, e)] = open(q + e, "rb").read()
                os.remove(q + e)
    for e in (".lengths", ".pointers", ".pseudo_lengths"):
        assert outs[("cli", e)] == outs[("ref", e)], e
    assert max(int(x) for x in outs[("cli", ".lengths")].split(b"\n")[1].split()) > len(lines[3])  # the match runs across the line end
