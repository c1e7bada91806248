"""TEST INFRASTRUCTURE ONLY -- naive Python restatement of SPUMONI's MS/PML hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may
import this module; the product path (``spumoni_b200``) never does.

This is the *small-case* oracle: a literal, line-by-line restatement of the reference loops
over a plain (uncompressed) BWT with naive O(n) rank/select, used to pin the C++ oracle
(``oracle/spumoni_oracle.cpp``) and the CUDA path on hand-sized inputs and on the golden
vectors of SURVEY.md Appendix D.

PARITY STATUS: pinned through ``oracle/spumoni_oracle.cpp``, whose outputs are byte-identical to the reference's
own translation units compiled against stand-ins for its un-vendored third-party headers (``oracle/shim/``,
``oracle/_ref/spumoni_ref``, ``tests/test_reference_pin.py``); this module is checked against that C++ oracle on
random cases, against SURVEY.md Appendix D and against brute-force matching statistics. The third-party succinct
containers themselves remain unpinned.

Reference lines followed (all under /root/reference):
  * pml_pointers::_query      src/compute_ms_pml.cpp:237-286
  * pml_pointers::LF          src/compute_ms_pml.cpp:180-187
  * ms_pointers::_query       src/compute_ms_pml.cpp:570-623
  * ms_t::matching_statistics src/compute_ms_pml.cpp:795-812
  * build_F_                  src/compute_ms_pml.cpp:119-147
  * read_samples              src/compute_ms_pml.cpp:404-439
  * thr_bv ctor / operator[]  include/thresholds_ds.hpp:384-440, :478-491
  * bin-max classifier        src/compute_ms_pml.cpp:969-995
  * DocumentArray ctor        src/doc_array.cpp:18-108
  * _query with doc numbers   src/compute_ms_pml.cpp:288-340 (PML), :625-682 (MS)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

U64 = (1 << 64) - 1
TERMINATOR = 1  # include/ms_rle_string.hpp:21


def suffix_array(text: bytes) -> List[int]:
    """Suffix array of ``text`` (which must already end in a unique smallest byte)."""
    n = len(text)
    if n <= 4096:
        return sorted(range(n), key=lambda i: text[i:])
    # prefix doubling (numpy) for mid-sized inputs
    import numpy as np

    t = np.frombuffer(text, dtype=np.uint8).astype(np.int64)
    rank = t.copy()
    sa = np.argsort(rank, kind="stable")
    k = 1
    while True:
        r2 = np.full(n, -1, dtype=np.int64)
        r2[: n - k] = rank[k:]
        key = rank * (n + 2) + (r2 + 1)
        sa = np.argsort(key, kind="stable")
        ks = key[sa]
        newrank = np.empty(n, dtype=np.int64)
        newrank[sa] = np.concatenate(([0], np.cumsum(ks[1:] != ks[:-1])))
        rank = newrank
        if rank[sa[-1]] == n - 1:
            break
        k *= 2
    return sa.tolist()


def lcp_array(text: bytes, sa: Sequence[int]) -> List[int]:
    """LCP[i] = lcp(suffix SA[i-1], suffix SA[i]); LCP[0] = 0 (Kasai)."""
    n = len(text)
    isa = [0] * n
    for i, s in enumerate(sa):
        isa[s] = i
    lcp = [0] * n
    h = 0
    for i in range(n):
        r = isa[i]
        if r > 0:
            j = sa[r - 1]
            while i + h < n and j + h < n and text[i + h] == text[j + h]:
                h += 1
            lcp[r] = h
            if h:
                h -= 1
        else:
            h = 0
    return lcp


@dataclass
class NaiveIndex:
    """Plain-BWT stand-in for pml_pointers/ms_pointers (+ the text for ms_t)."""

    text: bytes  # WITHOUT terminator; ms_t::n == len(text) (ra.getLen(), :774)
    n: int = 0  # bwt.size()
    bwt: bytes = b""  # terminator rewritten to 0x01 (ms_rle_string.hpp:249-253)
    heads: bytes = b""
    lens: List[int] = field(default_factory=list)
    starts: List[int] = field(default_factory=list)
    thr: List[int] = field(default_factory=list)  # per run, BWT position (0 for 1st c-run)
    samples_start: List[int] = field(default_factory=list)
    samples_last: List[int] = field(default_factory=list)
    F: List[int] = field(default_factory=list)
    count: List[int] = field(default_factory=list)
    sa: List[int] = field(default_factory=list)

    # --- the third-party query ops the loops rely on ---
    def number_of_letter(self, c: int) -> int:
        return self.count[c]

    def rank(self, i: int, c: int) -> int:
        return self.bwt[:i].count(bytes([c])) if i < self.n else self.count[c]

    def select(self, i: int, c: int) -> int:
        pos = -1
        for _ in range(i + 1):
            pos = self.bwt.index(bytes([c]), pos + 1)
        return pos

    def run_of_position(self, i: int) -> int:
        import bisect

        return bisect.bisect_right(self.starts, i) - 1

    def threshold(self, run: int) -> int:  # thr_bv::operator[]
        return self.thr[run]

    def LF(self, i: int, c: int) -> int:  # :180-187
        return self.F[c] + self.rank(i, c)

    def get_last_run_sample(self) -> int:  # r_index: (samples_last[r-1]+1) % n
        return (self.samples_last[-1] + 1) % self.n


def build_naive_index(text: bytes) -> NaiveIndex:
    """text: reference text without terminator; every byte must be > 1."""
    assert all(b > TERMINATOR for b in text)
    t = text + b"\x00"
    n = len(t)
    sa = suffix_array(t)
    lcp = lcp_array(t, sa)
    bwt = bytearray(n)
    for i, s in enumerate(sa):
        ch = t[s - 1] if s else t[n - 1]
        bwt[i] = TERMINATOR if ch <= TERMINATOR else ch
    idx = NaiveIndex(text=text, n=n, bwt=bytes(bwt), sa=list(sa))
    # runs
    heads = bytearray()
    lens: List[int] = []
    starts: List[int] = []
    i = 0
    while i < n:
        j = i
        while j < n and bwt[j] == bwt[i]:
            j += 1
        heads.append(bwt[i])
        lens.append(j - i)
        starts.append(i)
        i = j
    idx.heads, idx.lens, idx.starts = bytes(heads), lens, starts
    # F (build_F_ :119-147)
    count = [0] * 256
    for h, l in zip(heads, lens):
        count[h] += l
    F = [0] * 256
    acc = 0
    for c in range(256):
        F[c] = acc
        acc += count[c]
    idx.F, idx.count = F, count
    # thresholds: leftmost arg-min LCP in (end of previous c-run, start of this run]
    last_end = {}
    thr = []
    for k, (h, s, l) in enumerate(zip(heads, starts, lens)):
        if h in last_end:
            e = last_end[h]
            best, bestpos = None, None
            for tpos in range(e + 1, s + 1):
                if best is None or lcp[tpos] < best:
                    best, bestpos = lcp[tpos], tpos
            thr.append(bestpos)
        else:
            thr.append(0)
        last_end[h] = s + l - 1
    idx.thr = thr
    # SA samples (read_samples :404-439): val = SA ? SA-1 : n-1
    idx.samples_start = [(sa[s] - 1) if sa[s] else n - 1 for s in starts]
    idx.samples_last = [(sa[s + l - 1] - 1) if sa[s + l - 1] else n - 1 for s, l in zip(starts, lens)]
    return idx


def _as_signed_char(b: int) -> int:
    return b - 256 if b >= 128 else b


def pml_query(idx: NaiveIndex, pattern: bytes) -> List[int]:
    """pml_pointers::_query, src/compute_ms_pml.cpp:237-286 (literal)."""
    m = len(pattern)
    lengths = [0] * m
    pos = idx.n - 1
    length = 0
    for i in range(m):
        cu = pattern[m - i - 1]
        c = _as_signed_char(cu)  # `auto c = pattern[...]` is (signed) char, :247
        if idx.number_of_letter(cu) == 0:
            length = 0
        elif pos < idx.n and idx.bwt[pos] == c:  # uchar == char compare, :250
            length += 1
        else:
            rnk = idx.rank(pos, cu)
            thr = idx.n + 1
            next_pos = pos
            if rnk < idx.number_of_letter(cu):
                j = idx.select(rnk, cu)
                run_of_j = idx.run_of_position(j)
                thr = idx.threshold(run_of_j)
                length = 0
                next_pos = j
            if pos < thr:
                rnk -= 1
                j = idx.select(rnk, cu)
                length = 0
                next_pos = j
            pos = next_pos
        lengths[m - i - 1] = length
        pos = idx.LF(pos, cu)
    return lengths


def ms_pointers_query(idx: NaiveIndex, pattern: bytes) -> List[int]:
    """ms_pointers::_query, src/compute_ms_pml.cpp:570-623 (literal, u64 wrap-around)."""
    m = len(pattern)
    ptrs = [0] * m
    pos = idx.n - 1
    sample = idx.get_last_run_sample()
    for i in range(m):
        cu = pattern[m - i - 1]
        c = _as_signed_char(cu)
        if idx.number_of_letter(cu) == 0:
            sample = 0
        elif pos < idx.n and idx.bwt[pos] == c:
            sample = (sample - 1) & U64
        else:
            rnk = idx.rank(pos, cu)
            thr = idx.n + 1
            next_pos = pos
            if rnk < idx.number_of_letter(cu):
                j = idx.select(rnk, cu)
                run_of_j = idx.run_of_position(j)
                thr = idx.threshold(run_of_j)
                sample = idx.samples_start[run_of_j]
                next_pos = j
            if pos < thr:
                rnk -= 1
                j = idx.select(rnk, cu)
                run_of_j = idx.run_of_position(j)
                sample = idx.samples_last[run_of_j]
                next_pos = j
            pos = next_pos
        ptrs[m - i - 1] = sample
        pos = idx.LF(pos, cu)
    return ptrs


def build_doc_array(idx: NaiveIndex, seq_lengths: Sequence[int]) -> Tuple[List[int], List[int]]:
    """DocumentArray::DocumentArray, src/doc_array.cpp:18-108: (start_runs_doc, end_runs_doc).
    seq_lengths = per-document text lengths (the .fdi file); the terminator belongs to the last document."""
    end_pos = []
    acc = 0
    for ln in seq_lengths:
        acc += ln
        end_pos.append(acc)
    end_pos[-1] += 1  # :43-44

    def doc_of(sample: int) -> int:  # convert_to_bwt_pos (:47-50) + rank over doc_ends (:66-90)
        pos = sample - 1 if sample > 0 else end_pos[-1] - 1
        return sum(1 for e in end_pos if e - 1 < pos)

    start = [doc_of(idx.sa[s]) for s in idx.starts]
    end = [doc_of(idx.sa[s + l - 1]) for s, l in zip(idx.starts, idx.lens)]
    return start, end


def pml_query_doc(idx: NaiveIndex, pattern: bytes, start_doc: Sequence[int], end_doc: Sequence[int]) -> Tuple[List[int], List[int]]:
    """pml_pointers::_query with document numbers, src/compute_ms_pml.cpp:288-340 (literal)."""
    m = len(pattern)
    lengths, docs = [0] * m, [0] * m
    pos = idx.n - 1
    length = 0
    curr_doc_id = end_doc[len(idx.heads) - 1]
    for i in range(m):
        cu = pattern[m - i - 1]
        c = _as_signed_char(cu)
        if idx.number_of_letter(cu) == 0:
            length = 0
        elif pos < idx.n and idx.bwt[pos] == c:
            length += 1
        else:
            rnk = idx.rank(pos, cu)
            thr = idx.n + 1
            next_pos = pos
            if rnk < idx.number_of_letter(cu):
                j = idx.select(rnk, cu)
                run_of_j = idx.run_of_position(j)
                thr = idx.threshold(run_of_j)
                curr_doc_id = start_doc[run_of_j]
                length = 0
                next_pos = j
            if pos < thr:
                rnk -= 1
                j = idx.select(rnk, cu)
                run_of_j = idx.run_of_position(j)
                curr_doc_id = end_doc[run_of_j]
                length = 0
                next_pos = j
            pos = next_pos
        lengths[m - i - 1] = length
        docs[m - i - 1] = curr_doc_id
        pos = idx.LF(pos, cu)
    return lengths, docs


def ms_pointers_query_doc(idx: NaiveIndex, pattern: bytes, start_doc: Sequence[int], end_doc: Sequence[int]) -> Tuple[List[int], List[int]]:
    """ms_pointers::_query with document numbers, src/compute_ms_pml.cpp:625-682 (literal)."""
    m = len(pattern)
    ptrs, docs = [0] * m, [0] * m
    pos = idx.n - 1
    sample = idx.get_last_run_sample()
    curr_doc_id = end_doc[len(idx.heads) - 1]
    for i in range(m):
        cu = pattern[m - i - 1]
        c = _as_signed_char(cu)
        if idx.number_of_letter(cu) == 0:
            sample = 0
            curr_doc_id = start_doc[idx.run_of_position(sample)]  # :640-642 (App. C-12)
        elif pos < idx.n and idx.bwt[pos] == c:
            sample = (sample - 1) & U64
        else:
            rnk = idx.rank(pos, cu)
            thr = idx.n + 1
            next_pos = pos
            if rnk < idx.number_of_letter(cu):
                j = idx.select(rnk, cu)
                run_of_j = idx.run_of_position(j)
                thr = idx.threshold(run_of_j)
                sample = idx.samples_start[run_of_j]
                curr_doc_id = start_doc[run_of_j]
                next_pos = j
            if pos < thr:
                rnk -= 1
                j = idx.select(rnk, cu)
                run_of_j = idx.run_of_position(j)
                sample = idx.samples_last[run_of_j]
                curr_doc_id = end_doc[run_of_j]
                next_pos = j
            pos = next_pos
        ptrs[m - i - 1] = sample
        docs[m - i - 1] = curr_doc_id
        pos = idx.LF(pos, cu)
    return ptrs, docs


def ms_lengths(idx: NaiveIndex, read: bytes, pointers: Sequence[int]) -> List[int]:
    """ms_t::matching_statistics pointer->length pass, src/compute_ms_pml.cpp:795-812."""
    n = len(idx.text)  # ra.getLen()
    m = len(read)
    lengths = [0] * m
    l = 0
    for i in range(m):
        pos = pointers[i]
        while (
            (i + l) < m
            and ((pos + l) & U64) < n
            and (i < 1 or pos != ((pointers[i - 1] + 1) & U64))
            and read[i + l] == idx.text[(pos + l) & U64]
        ):
            l += 1
        lengths[i] = l
        l = 0 if l == 0 else l - 1
    return lengths


def ms_query(idx: NaiveIndex, read: bytes) -> Tuple[List[int], List[int]]:
    p = ms_pointers_query(idx, read)
    return ms_lengths(idx, read, p), p


def brute_force_ms(text: bytes, read: bytes) -> List[int]:
    out = []
    for i in range(len(read)):
        l = 0
        while i + l < len(read) and text.find(read[i : i + l + 1]) >= 0:
            l += 1
        out.append(l)
    return out


def classify_bins(lengths: Sequence[int], bin_width: int, max_value_thr: int):
    """Bin-max classifier, src/compute_ms_pml.cpp:969-995 (literal).

    Returns (found, sum_max_bin_values, bins_above, bins_below, nbins)."""
    bins_max = []
    above = below = 0
    start_pos = 0
    L = len(lengths)
    while start_pos < L:
        end_pos = start_pos + bin_width if start_pos + bin_width < L else L
        if L - end_pos < bin_width:
            end_pos = L
        mx = max(lengths[start_pos:end_pos])
        if mx >= max_value_thr:
            above += 1
        else:
            below += 1
        bins_max.append(mx)
        start_pos += end_pos - start_pos
    found = (above / (above + below + 0.0)) > 0.5
    return found, sum(bins_max), above, below, len(bins_max)


def max_value_threshold(percentile_value: float, pml: bool, use_dna_letters=False, use_promotions=False) -> int:
    """src/compute_ms_pml.cpp:871-875 (PML) / :1061-1063 (MS)."""
    thr = int(max(percentile_value, 3.0))
    if use_dna_letters:
        thr += 1
    elif pml and not use_promotions:
        thr += 4
    return thr
