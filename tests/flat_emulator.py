"""Test helper: a slow Python walk over the flat device image (spumoni_b200/csrc/flat_index.h).

Follows spb_walk_kernel's per-step logic (kernels.cu) statement by statement so that the host
side of the product (on-disk parse -> flat image) can be checked against the oracle on a box
without a GPU.  It is NOT a product path: only tests import it.
"""
from __future__ import annotations

import struct

import numpy as np

U64 = (1 << 64) - 1
HDR_FMT = "<QIIQQQQIIIIQQQQQQQQQQIII4II"  # up to pad0 (then F[256], chartab[256])


class FlatImage:
    def __init__(self, blob: np.ndarray):
        b = bytes(blob)
        self.b = b
        (self.magic, self.version, self.mode, self.n, self.R, self.n_text, self.last_run_sample, self.sigma,
         self.sigma_pad, self.blk_stride, self.num_blocks, self.num_crec, self.off_runs, self.off_blocks,
         self.off_crec, self.off_csamp, self.off_text, self.total_bytes, self.terminator_position, self.off_fast,
         self.off_rl, self.num_fast, self.init_off, self.init_h, fs0, fs1, fs2, fs3, _pad) = struct.unpack_from(HDR_FMT, b, 0)
        self.fast_sym = [fs0, fs1, fs2, fs3]
        o = struct.calcsize(HDR_FMT)
        self.F = struct.unpack_from("<256Q", b, o)
        o += 256 * 8
        self.chartab = [struct.unpack_from("<4I", b, o + 16 * c) for c in range(256)]
        assert self.total_bytes == len(b)
        nrec = self.R + 64
        self.runs = np.frombuffer(b, dtype=np.uint32, count=4 * nrec, offset=self.off_runs).reshape(nrec, 4)
        self.crec = np.frombuffer(b, dtype=np.uint32, count=4 * (self.num_crec + 1), offset=self.off_crec).reshape(-1, 4)
        if self.num_fast:
            self.rows = np.frombuffer(b, dtype=np.uint32, count=16 * self.R, offset=self.off_fast).reshape(self.R, 4, 4)
            self.rl = np.frombuffer(b, dtype=np.uint32, count=self.R + 16, offset=self.off_rl)
        if self.mode == 1:
            self.csamp = np.frombuffer(b, dtype=np.uint64, count=2 * (self.num_crec + 1), offset=self.off_csamp).reshape(-1, 2)
            self.text = b[self.off_text:self.off_text + self.n_text]

    def run(self, k):
        x, y, z, w = (int(v) for v in self.runs[k])
        return {"start": x | (((w >> 8) & 0xFF) << 32), "lf": y | (((w >> 16) & 0xFF) << 32), "dest": z, "head": w & 0xFF}

    def locate(self, target, w0):
        assert self.run(w0)["start"] <= target, "window does not start at or before the target"
        k = w0
        while self.run(k + 1)["start"] <= target:
            k += 1
        return k

    def crun_ordinal(self, k, c, cidx):
        hdr = self.off_blocks + (k >> 5) * self.blk_stride
        heads = self.b[hdr:hdr + 32]
        base = struct.unpack_from("<I", self.b, hdr + 32 + 4 * cidx)[0]
        return base + sum(1 for t in range(k & 31) if heads[t] == c)

    def walk(self, read: bytes):
        """Returns (pml_lengths, ms_pointers or None)."""
        m = len(read)
        ms = self.mode == 1
        n, R = self.n, self.R
        pos, k = n - 1, R - 1
        kr = self.run(k)
        length, sample = 0, self.last_run_sample
        out_len, out_ptr = [0] * m, [0] * m
        for i in range(m):
            c = read[m - 1 - i]
            ct = self.chartab[c]
            if ct[3] == 0xFFFFFFFF:
                length, sample = 0, 0
                npos = ct[0] | (ct[1] << 32)
                w0 = ct[2]
            elif k < R and kr["head"] == c and c < 0x80:
                length += 1
                sample = (sample - 1) & U64
                npos = kr["lf"] + (pos - kr["start"])
                w0 = kr["dest"]
            else:
                j = self.crun_ordinal(k, c, ct[3]) if k < R else ct[1]
                x, y, z, w = (int(v) for v in self.crec[ct[0] + j])
                thr = x | ((w & 0xFF) << 32)
                lf = y | (((w >> 8) & 0xFF) << 32)
                down = pos < thr
                same = k < R and kr["head"] == c
                length = 0
                if ms:
                    ss, sl = (int(v) for v in self.csamp[ct[0] + j])
                    if j < ct[1]:
                        sample = ss
                    if down:
                        sample = sl
                npos = lf - 1 if down else (lf + (pos - kr["start"]) if same else lf)
                w0 = z - 1 if (down and z > 0) else z
            out_len[m - 1 - i] = length
            out_ptr[m - 1 - i] = sample
            pos = npos
            if pos >= n:
                k = R
            else:
                k = self.locate(pos, w0)
            kr = self.run(k)
        return out_len, (out_ptr if ms else None)

    def ms_lengths(self, read: bytes, ptrs):
        m, n_text = len(read), self.n_text
        out = [0] * m
        l = 0
        for i in range(m):
            p = ptrs[i]
            if i == 0 or p != ((ptrs[i - 1] + 1) & U64):
                while i + l < m and ((p + l) & U64) < n_text and read[i + l] == self.text[(p + l) & U64]:
                    l += 1
            out[i] = l
            l = l - 1 if l else 0
        return out


    # ---- PML fast path (spb_pml_fast_kernel), including its fall-backs; sat26/sat29 let a test lower the
    # saturation limits the way the SPB_FAST_SAT / SPB_RL_SAT test build does
    def slow_step(self, k, off, c, length):
        n, R = self.n, self.R
        kr = self.run(k)
        pos = kr["start"] + off if k < R else n
        ct = self.chartab[c]
        if ct[3] == 0xFFFFFFFF:
            length = 0
            npos, w0 = ct[0] | (ct[1] << 32), ct[2]
        elif k < R and kr["head"] == c and c < 0x80:
            length += 1
            npos, w0 = kr["lf"] + off, kr["dest"]
        else:
            j = self.crun_ordinal(k, c, ct[3]) if k < R else ct[1]
            x, y, z, w = (int(v) for v in self.crec[ct[0] + j])
            thr, lf = x | ((w & 0xFF) << 32), y | (((w >> 8) & 0xFF) << 32)
            down = pos < thr
            same = k < R and kr["head"] == c
            length = 0
            npos = lf - 1 if down else (lf + off if same else lf)
            w0 = z - 1 if (down and z > 0) else z
        if npos >= n:
            return R, 0, length
        k2 = self.locate(npos, w0)
        return k2, npos - self.run(k2)["start"], length

    def fid(self, c):
        ct = self.chartab[c]
        return 5 if ct[3] == 0xFFFFFFFF else ct[2]

    def walk_fast(self, read: bytes, sat26=0x3FFFFFF, sat29=0x1FFFFFFF):
        assert self.num_fast and self.mode == 0
        m, R = len(read), self.R
        k, off, h, length = R - 1, self.init_off, self.init_h, 0
        out = [0] * m
        stats = {"fast": 0, "scan_iters": 0, "slow": 0}
        for gi in range(m - 1, -1, -1):
            c = read[gi]
            s = self.fid(c)
            slow = not (s < 4 and h < 4)
            if not slow:
                match = s == h
                x, y, z, w = (int(v) for v in self.rows[k][0 if match else 1 + s - (1 if s > h else 0)])
                lt = off < x
                if match:
                    if lt:
                        k, off, h, length = y, off + z, (w >> 26) & 7, length + 1
                    else:
                        r2, len1 = off - x, w & 0x3FFFFFF
                        if r2 < len1:
                            k, off, h, length = y + 1, r2, w >> 29, length + 1
                        elif len1 == sat26:
                            slow = True
                        else:
                            length += 1
                            sk, rem = y + 2, r2 - len1
                            while True:
                                stats["scan_iters"] += 1
                                wv = int(self.rl[sk])
                                ln = wv & 0x1FFFFFFF
                                if ln == sat29:
                                    st = self.run(sk)["start"] + rem
                                    k = self.locate(st, sk)
                                    off = st - self.run(k)["start"]
                                    h = min(self.fid(self.run(k)["head"]), 4)
                                    break
                                if rem < ln:
                                    k, off, h = sk, rem, wv >> 29
                                    break
                                rem -= ln
                                sk += 1
                else:
                    if not lt:
                        nk, noff, nh = y, z & 0x3FFFFFF, (z >> 26) & 7
                    else:
                        nk, noff, nh = y - ((z >> 29) & 1), w & 0x3FFFFFF, (w >> 26) & 7
                    if noff == sat26:
                        slow = True
                    else:
                        k, off, h, length = nk, noff, nh, 0
                if not slow:
                    stats["fast"] += 1
            if slow:
                stats["slow"] += 1
                k, off, length = self.slow_step(k, off, c, length)
                h = min(self.fid(self.run(k)["head"]), 4) if k < R else 4
            out[gi] = length
        return out, stats
