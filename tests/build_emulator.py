"""Test helper: a numpy restatement of the GPU index builder's algorithm (spumoni_b200/csrc/build_kernels.cu), step for
step -- prefix doubling over (rank, rank + h) keys, PLCP by the Phi algorithm in independent text chunks, leftmost range
minima over LCP by 256-element blocks + a sparse table over the block minima -- so that its arithmetic (terminator
handling, tie breaking, threshold ranges) can be checked against the suffix-array construction of tools/mkindex on a box
without a GPU.  NOT a product path: only tests import it."""
from __future__ import annotations

import numpy as np

BLOCK = 256


def suffix_array(T: np.ndarray) -> np.ndarray:
    """T: uint8 text ending with its unique smallest terminator 0."""
    n = len(T)
    pad = np.concatenate([T, np.zeros(8, dtype=np.uint8)])
    key = np.zeros(n, dtype=np.uint64)
    for b in range(8):
        key = (key << np.uint64(8)) | pad[b:b + n].astype(np.uint64)
    sa = np.argsort(key, kind="stable").astype(np.int64)
    key = key[sa]
    h = 8
    while True:
        flag = np.ones(n, dtype=bool)
        flag[1:] = key[1:] != key[:-1]
        start = np.maximum.accumulate(np.where(flag, np.arange(n), 0))  # rank = index of the group's first element
        if flag.all():
            return sa
        rank = np.zeros(n, dtype=np.int64)
        rank[sa] = start
        nxt = sa + h
        lo = np.where(nxt < n, rank[np.minimum(nxt, n - 1)] + 1, 0)
        key = (rank[sa].astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
        order = np.argsort(key, kind="stable")
        sa, key = sa[order], key[order]
        h *= 2


def plcp_phi(T: np.ndarray, sa: np.ndarray, chunk: int = 64) -> np.ndarray:
    n = len(T)
    phi = np.full(n, -1, dtype=np.int64)
    phi[sa[1:]] = sa[:-1]
    plcp = np.zeros(n, dtype=np.int64)
    for lo in range(0, n, chunk):  # every chunk is an independent GPU thread: it restarts from h = 0
        h = 0
        for i in range(lo, min(n, lo + chunk)):
            j = phi[i]
            if j < 0:
                h = 0
                continue
            while i + h < n and j + h < n and T[i + h] == T[j + h]:
                h += 1
            plcp[i] = h
            if h:
                h -= 1
    return plcp


class LeftmostRmq:
    def __init__(self, v: np.ndarray):
        self.v = v
        n = len(v)
        nb = (n + BLOCK - 1) // BLOCK
        first = np.zeros(nb, dtype=np.int64)
        for b in range(nb):
            seg = v[b * BLOCK:(b + 1) * BLOCK]
            first[b] = b * BLOCK + int(np.argmin(seg))  # np.argmin returns the first minimum
        self.levels = [first]
        j = 1
        while (1 << j) <= nb:
            prev = self.levels[-1]
            half = 1 << (j - 1)
            a, b = prev[: nb - (1 << j) + 1], prev[half: half + nb - (1 << j) + 1]
            self.levels.append(np.where(v[b] < v[a], b, a))  # ties keep the left candidate
            j += 1

    def _better(self, a, b):  # a lies left of b
        return b if self.v[b] < self.v[a] else a

    def query(self, l: int, r: int) -> int:
        """leftmost argmin of v[l..r], inclusive"""
        bl, br = l // BLOCK, r // BLOCK
        if bl == br:
            return l + int(np.argmin(self.v[l:r + 1]))
        best = l + int(np.argmin(self.v[l:(bl + 1) * BLOCK]))
        if br - bl > 1:
            a, b = bl + 1, br - 1
            j = (b - a + 1).bit_length() - 1
            mid = self._better(int(self.levels[j][a]), int(self.levels[j][b - (1 << j) + 1]))
            best = self._better(best, mid)
        tail = br * BLOCK + int(np.argmin(self.v[br * BLOCK:r + 1]))
        return self._better(best, tail)


def build_parts(T: np.ndarray):
    """-> dict(heads, lens, thr, ssa, esa) like tools/mkindex build_rlbwt."""
    n = len(T)
    sa = suffix_array(T)
    plcp = plcp_phi(T, sa)
    lcp = plcp[sa]
    bwt = np.where(sa > 0, T[np.maximum(sa - 1, 0)], T[n - 1])
    flag = np.ones(n, dtype=bool)
    flag[1:] = bwt[1:] != bwt[:-1]
    starts = np.nonzero(flag)[0]
    ends = np.concatenate([starts[1:], [n]]) - 1
    heads = bwt[starts]
    rmq = LeftmostRmq(lcp)
    thr = np.zeros(len(starts), dtype=np.int64)
    prev_end = {}
    for k in range(len(starts)):  # on the GPU: previous run of the same head via a per-symbol scan, one thread per run
        c = int(heads[k])
        if c in prev_end:
            thr[k] = rmq.query(prev_end[c] + 1, int(starts[k]))
        prev_end[c] = int(ends[k])
    return {"heads": heads, "lens": ends - starts + 1, "thr": thr, "ssa": sa[starts], "esa": sa[ends], "n": n}
