"""Generates tests/golden/reference_run/ by running the REFERENCE's own code (run from the repo root, in the container
that has /root/reference).

oracle/_ref/spumoni_ref is the reference's translation units (compute_ms_pml.cpp, doc_array.cpp, emp_null_database.cpp,
ks_test.cpp, batch_loader.cpp, spumoni.cpp) compiled where they lie against stand-ins for the third-party headers it
does not vendor (oracle/shim/, oracle/Makefile.ref). Everything written here except the input FASTA files comes out of
that binary:
  idx.fa.thrbv.spumoni / .thrbv.ms   build_spumoni_main / build_spumoni_ms_main on mkindex's PFP-style intermediates
  idx.fa.doc                         DocumentArray(ref, runs).serialize
  idx.fa.pmlnulldb / .msnulldb       EmpNullDatabase(...) + find_threshold_based_on_null_*_distribution + serialize
  expected/<mode>/reads.fa.*         `spumoni run -r idx -p reads.fa <flags>`
The index is small on purpose (the files are committed); tests/test_reference_pin.py compares larger, uncommitted
fixtures against the binary whenever it is present.
"""
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
MKINDEX = os.path.join(ROOT, "tools", "bin", "spumoni_mkindex")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "spumoni_ref")
OUT = os.path.join(ROOT, "tests", "golden", "reference_run")

MODES = {"P": ("reads.fa", ["-P", "-n", "-c"]), "Pd": ("reads.fa", ["-P", "-n", "-d"]), "M": ("reads.fa", ["-M", "-n", "-c"]),
         "Md": ("reads.fa", ["-M", "-n", "-d", "-c"]),
         # general text (-g): reads separated by 0x01, used verbatim (no upper-casing), ids are read_<i>
         "Pg": ("general.txt", ["-P", "-n", "-g"]), "Mg": ("general.txt", ["-M", "-n", "-g"]),
         # FASTQ records in a file with a FASTA extension (the extension is all that is validated): BatchLoader takes them,
         # and loses the records of the last, short batch when the file ends with a newline (batch_loader.cpp:51)
         "Pq": ("fastq_reads.fa", ["-P", "-n", "-c"]), "Mq": ("fastq_reads.fa", ["-M", "-n"])}


def fasta_text(path):
    return "".join(line.strip() for line in open(path) if not line.startswith(">"))


def write_reads(path, text, rng):
    reads = []
    for i in range(28):
        ln = rng.randint(20, 160)
        if rng.random() < 0.7:
            s = rng.randint(0, len(text) - ln)
            r = list(text[s:s + ln])
            for k in range(len(r)):
                if rng.random() < 0.06:
                    r[k] = rng.choice("ACGT")
            reads.append(("read%d" % i, "".join(r)))
        else:
            reads.append(("neg%d" % i, "".join(rng.choice("ACGT") for _ in range(ln))))
    reads += [("lower", text[100:260].lower()), ("withN", text[300:360] + "NNNN" + text[364:420]), ("allN", "N" * 30),
              ("absent", "XYZ" + text[500:540] + "X" + text[541:580] + "QQ"), ("one", "A"), ("oneN", "N"), ("oneX", "X"),
              ("two", "GT"), ("startX", "X" + text[600:660]), ("endX", text[600:660] + "X"), ("textend", text[-120:]),
              ("textstart", text[:120]), ("symbols", "AC-GT*AC.GT#ACGT~ACGT"), ("lowabsent", "acgtxacgtnacgt"),
              ("long", text[700:1500])]
    with open(path, "w") as f:
        for name, seq in reads:
            f.write(">%s\n%s\n" % (name, seq))
        f.write(">multi line with comment\n%s\n%s\n%s\n" % (text[900:960], text[960:1000], text[1000:1001]))


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref", "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    rng = random.Random(20261017)
    with tempfile.TemporaryDirectory() as d:
        raw = os.path.join(d, "raw.fa")
        subprocess.check_call([MKINDEX, "genome", "--kind", "pangenome", "--len", "1200", "--haps", "3", "--snp", "0.02",
                               "--indel", "0.002", "--seed", "11", "--out", raw])
        out = subprocess.check_output([MKINDEX, "build", "--input", raw, "--prefix", os.path.join(d, "idx"), "--pml", "--ms",
                                       "--keep", "--docs", "2"], text=True)
        runs = int(out.split('"r":')[1].strip(" }\n"))
        ref = os.path.join(d, "idx.fa")
        for ext in (".thrbv.spumoni", ".thrbv.ms", ".doc"):  # mkindex's own copies: replaced by the reference's
            os.rename(ref + ext, ref + ext + ".mkindex")
        subprocess.check_call([REF_BIN, "x-build-pml", ref])
        subprocess.check_call([REF_BIN, "x-build-ms", ref])
        subprocess.check_call([REF_BIN, "x-doc", ref, str(runs)])
        for ext in (".thrbv.spumoni", ".thrbv.ms", ".doc"):
            assert open(ref + ext, "rb").read() == open(ref + ext + ".mkindex", "rb").read(), ext
        text = fasta_text(ref)
        open(ref + ".slp", "w").write(text)  # the shim's random-access text (stands in for the SLP grammar)
        reads = os.path.join(d, "reads.fa")
        write_reads(reads, text, rng)
        general = os.path.join(d, "general.txt")
        with open(general, "wb") as f:
            for seq in (text[50:170], text[200:260].lower(), "ACGT\nACGT", ">not a header", text[300:330] + "N" + text[331:380], "A", "",
                        text[400:1000], "GATTACA" * 9, "acgtACGTxyz", text[-80:]):
                f.write(seq.encode() + b"\x01")
            f.write(b"trailing bytes without a separator are not a read")
        fq_rng = random.Random(777)  # its own stream: the files made before FASTQ support stay byte-identical
        with open(os.path.join(d, "fastq_reads.fa"), "w") as f:
            for i in range(23):
                p0 = fq_rng.randint(0, len(text) - 151)
                f.write("@fq%d comment\n%s\n+\n%s\n" % (i, text[p0:p0 + 150], "I" * 150))
        null_reads = os.path.join(d, "null_reads.fa")
        with open(null_reads, "w") as f:
            for i in range(12):
                s = rng.randint(0, len(text) - 200)
                f.write(">null%d\n%s\n" % (i, text[s:s + 200]))
        srand_note = []
        for kind in ("pml", "ms"):
            srand_note.append(subprocess.check_output([REF_BIN, "x-nulldb", ref, null_reads, kind, "150"], text=True,
                                                      stderr=subprocess.DEVNULL).strip())
        shutil.rmtree(OUT, ignore_errors=True)
        os.makedirs(OUT)
        for name in ("idx.fa", "idx.fa.fdi", "idx.fa.thrbv.spumoni", "idx.fa.thrbv.ms", "idx.fa.doc", "idx.fa.pmlnulldb",
                     "idx.fa.msnulldb", "reads.fa", "general.txt", "fastq_reads.fa", "null_reads.fa"):
            shutil.copy(os.path.join(d, name), os.path.join(OUT, name))
        for mode, (pattern, flags) in MODES.items():
            subprocess.check_call([REF_BIN, "run", "-r", os.path.join(d, "idx"), "-p", os.path.join(d, pattern)] + flags,
                                  stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
            dst = os.path.join(OUT, "expected", mode)
            os.makedirs(dst)
            for f in sorted(os.listdir(d)):
                if f.startswith(pattern + "."):
                    shutil.move(os.path.join(d, f), os.path.join(dst, f))
        with open(os.path.join(OUT, "README.txt"), "w") as f:
            f.write("Written by tests/golden/make_reference_golden.py from oracle/_ref/spumoni_ref (the reference's own code).\n"
                    "runs = %d\n%s\n" % (runs, "\n".join(srand_note)))
    total = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(OUT) for f in fs)
    print("wrote", OUT, total, "bytes")


if __name__ == "__main__":
    sys.exit(main())
