"""Generates tests/golden/golden_vectors.json from oracle/naive.py (run from the repo root).

The reference ships no golden vectors for this path and cannot be executed
in this image, so these vectors come from the literal restatement in oracle/naive.py, which
itself reproduces SURVEY.md Appendix D (checked in tests/test_oracle.py). Outputs of the reference's own
code live next door: tests/golden/reference_run/ (make_reference_golden.py).
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import naive  # noqa: E402


def main():
    rng = random.Random(20261017)
    cases = []
    texts = ["GATTACAGATTACCGATTAGA"]
    for alpha, n in (("ACGT", 60), ("ACGTN", 200), ("AC", 50), ("ACGT\xe9", 120), ("ACGT", 500)):
        texts.append("".join(rng.choice(alpha) for _ in range(n)))
    texts.append("ACGTACGTTTGACA" * 12 + "GGGCCCAT" * 9)
    for t in texts:
        tb = t.encode("latin1")
        idx = naive.build_naive_index(tb)
        reads = []
        if t == texts[0]:
            reads += [b"TTACAGGATTAG", b"CCGATNAGATT", b"GATTYACA"]
        for _ in range(8):
            ln = rng.randint(1, 50)
            if rng.random() < 0.6:
                s = rng.randint(0, max(0, len(tb) - ln))
                r = bytearray(tb[s:s + ln])
                for k in range(len(r)):
                    if rng.random() < 0.1:
                        r[k] = rng.choice(b"ACGTNY\xe9\xfc")
                reads.append(bytes(r))
            else:
                reads.append(bytes(rng.choice(b"ACGTNY\xe9") for _ in range(ln)))
        out = []
        for r in reads:
            if not r:
                continue
            L, P = naive.ms_query(idx, r)
            out.append({"read": r.decode("latin1"), "pml": naive.pml_query(idx, r), "ms_pointers": [str(p) for p in P],
                        "ms_lengths": L})
        cases.append({"text": t, "n": idx.n, "r": len(idx.heads), "reads": out})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_vectors.json")
    with open(path, "w") as f:
        json.dump({"source": "oracle/naive.py (literal restatement; reproduces SURVEY.md Appendix D)", "cases": cases}, f, indent=0)
    print("wrote", path, len(cases), "cases")


if __name__ == "__main__":
    main()
