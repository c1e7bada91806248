"""Test helper: a slow Python walk over the flat device image (spumoni_b200/csrc/flat_index.h).

Follows spb_walk_kernel's per-step logic (kernels.cu) statement by statement so that the host
side of the product (on-disk parse -> flat image) can be checked against the oracle on a box
without a GPU.  It is NOT a product path: only tests import it.
"""
from __future__ import annotations

import struct

import numpy as np

U64 = (1 << 64) - 1
HDR_FMT = "<QIIQQQQIIIIQQQQQQQQQQI4I4I4I3I4QQQQ3I4IIQQII"  # up to pad3 (then F[256], chartab[256])


class FlatImage:
    def __init__(self, blob: np.ndarray):
        b = bytes(blob)
        self.b = b
        (self.magic, self.version, self.mode, self.n, self.R, self.n_text, self.last_run_sample, self.sigma,
         self.sigma_pad, self.blk_stride, self.num_blocks, self.num_crec, self.off_runs, self.off_blocks,
         self.off_crec, self.off_csamp, self.off_text, self.total_bytes, self.terminator_position, self.off_mrows,
         self.off_msamp, self.num_fast, *rest) = struct.unpack_from(HDR_FMT, b, 0)
        self.fast_sym, self.mv_end_k, self.mv_end_off, self.mv_end_sl = rest[0:4], rest[4:8], rest[8:12], rest[15:19]
        self.off_mrows8, self.off_cdoc, self.off_mdoc = rest[19], rest[20], rest[21]
        self.num_docs, self.doc_init, self.doc_absent_ms = rest[22], rest[23], rest[24]
        self.mv_end_doc = rest[25:29]
        self.mrows8_stride = rest[29]
        self.off_mrows32 = rest[30]
        self.R_file = rest[31]
        self.msamp_stride = rest[32]
        o = struct.calcsize(HDR_FMT)
        self.F = struct.unpack_from("<256Q", b, o)
        o += 256 * 8
        self.chartab = [struct.unpack_from("<4I", b, o + 16 * c) for c in range(256)]
        assert self.total_bytes == len(b)
        nrec = self.R + 64
        self.runs = np.frombuffer(b, dtype=np.uint32, count=4 * nrec, offset=self.off_runs).reshape(nrec, 4)
        self.crec = np.frombuffer(b, dtype=np.uint32, count=4 * (self.num_crec + 1), offset=self.off_crec).reshape(-1, 4)
        if self.num_fast:
            if self.off_mrows8:
                st = self.mrows8_stride // 8
                raw = np.frombuffer(b, dtype=np.uint64, count=(self.R + 64) * st, offset=self.off_mrows8).reshape(-1, st)
                self.rows8 = raw[:, 0]
                if st == 2:  # inline 32-bit samples: {ss, sl} per row, presented like msamp
                    assert self.off_msamp == 0
                    self.msamp = np.stack([raw[:, 1] & 0xFFFFFFFF, raw[:, 1] >> 32], axis=1)
            elif self.off_mrows32:
                self.rows32 = np.frombuffer(b, dtype=np.uint32, count=8 * (self.R + 16), offset=self.off_mrows32).reshape(-1, 8)
            else:
                self.rows = np.frombuffer(b, dtype=np.uint32, count=4 * (self.R + 16), offset=self.off_mrows).reshape(-1, 4)
            if self.mode == 1 and self.off_msamp and self.msamp_stride == 8:  # {u32 ss, u32 sl} per row, presented like msamp
                raw = np.frombuffer(b, dtype=np.uint64, count=self.R + 16, offset=self.off_msamp)
                self.msamp = np.stack([raw & 0xFFFFFFFF, raw >> 32], axis=1)
            elif self.mode == 1 and self.off_msamp:
                self.msamp = np.frombuffer(b, dtype=np.uint64, count=2 * (self.R + 16), offset=self.off_msamp).reshape(-1, 2)
        if self.num_docs:
            self.cdoc = np.frombuffer(b, dtype=np.uint32, count=2 * (self.num_crec + 1), offset=self.off_cdoc).reshape(-1, 2)
            if self.off_mdoc:
                self.mdoc = np.frombuffer(b, dtype=np.uint32, count=2 * (self.R + 64), offset=self.off_mdoc).reshape(-1, 2)
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


    # ---- thread-per-read walk over the move rows (spb_move_kernel), including its fall-backs; the limits
    # let a test decode an image written by the SPB_MV_* test build
    def slow_step(self, k, off, c, length, sample, doc=0):
        """Returns (k, off, length, sample); the document number of the step is left in self.last_doc."""
        n, R = self.n, self.R
        ms = self.mode == 1
        if k < R:
            if off == 0xFFFFFFFF:
                pos = self.run(k + 1)["start"] - 1
            else:
                pos = self.run(k)["start"] + off
                k = self.locate(pos, k)
        else:
            k, pos = R, n
        kr = self.run(k)
        ct = self.chartab[c]
        if ct[3] == 0xFFFFFFFF:
            length, sample = 0, 0
            if ms and self.num_docs:
                doc = self.doc_absent_ms
            npos, w0 = ct[0] | (ct[1] << 32), ct[2]
        elif k < R and kr["head"] == c and c < 0x80:
            length += 1
            sample = (sample - 1) & U64
            npos, w0 = kr["lf"] + (pos - kr["start"]), kr["dest"]
        else:
            j = self.crun_ordinal(k, c, ct[3]) if k < R else ct[1]
            x, y, z, w = (int(v) for v in self.crec[ct[0] + j])
            thr, lf = x | ((w & 0xFF) << 32), y | (((w >> 8) & 0xFF) << 32)
            down = pos < thr
            same = k < R and kr["head"] == c
            length = 0
            if ms:
                ss, sl = (int(v) for v in self.csamp[ct[0] + j])
                if j < ct[1]:
                    sample = ss
                if down:
                    sample = sl
            if self.num_docs:
                sd, ed = (int(v) for v in self.cdoc[ct[0] + j])
                if j < ct[1]:
                    doc = sd
                if down:
                    doc = ed
            npos = lf - 1 if down else (lf + (pos - kr["start"]) if same else lf)
            w0 = z - 1 if (down and z > 0) else z
        self.last_doc = doc
        if npos >= n:
            return R, 0, length, sample
        k2 = self.locate(npos, w0)
        return k2, npos - self.run(k2)["start"], length, sample

    def fid(self, c):
        ct = self.chartab[c]
        return 5 if ct[3] == 0xFFFFFFFF else ct[2]

    def walk_move8(self, read: bytes, probes=24):
        """Compact 8-byte rows (spb_move_kernel<ROW8>). Returns (lengths, pointers or None, stats)."""
        assert self.num_fast and self.off_mrows8
        ms = self.mode == 1
        LAST = 0xFFFFFFFF
        m, R = len(read), self.R
        k, off, length, sample = R - 1, LAST, 0, self.last_run_sample
        doc = self.doc_init if self.num_docs else 0
        out, ptr, self.docs = [0] * m, [0] * m, [0] * m
        stats = {"fast": 0, "norm_iters": 0, "slow": 0, "jump2": 0, "end": 0, "probe": 0}
        for gi in range(m - 1, -1, -1):
            c = read[gi]
            slow = False
            while True:
                v = int(self.rows8[k])
                lo, hi = v & 0xFFFFFFFF, v >> 32
                if lo >> 31:
                    slow = True
                    break
                ln = (lo >> 24) & 0x7F
                if off == LAST:
                    off = ln - 1
                if off >= ln:
                    off -= ln
                    k += 1
                    stats["norm_iters"] += 1
                    continue
                break
            if not slow:
                h, s = (hi >> 8) & 7, self.fid(c)
                if s >= 4:
                    slow = True
                elif s == h:
                    length += 1
                    sample = (sample - 1) & U64
                    k, off = lo & 0xFFFFFF, off + (hi & 0xFF)
                else:
                    slot = s - (1 if s > h else 0)
                    cls, dc = (hi >> (11 + 2 * slot)) & 3, (hi >> (23 + 3 * slot)) & 7
                    if cls == 3:
                        slow = True
                    elif dc == 0:
                        k, off, length, sample = self.mv_end_k[s], self.mv_end_off[s], 0, self.mv_end_sl[s]
                        if self.num_docs:
                            doc = self.mv_end_doc[s]
                        stats["end"] += 1
                    else:
                        down = cls == 1 or (cls == 2 and off < ((hi >> 17) & 0x3F))
                        k2, left = k + dc, probes
                        while True:
                            v2 = int(self.rows8[k2])
                            lo2, hi2 = v2 & 0xFFFFFFFF, v2 >> 32
                            if (hi2 >> 8) & 7 == s:
                                break
                            stats["probe"] += 1
                            left -= 1
                            if left == 0:
                                k2 = None
                                break
                            k2 += 1
                        if k2 is None:
                            slow = True
                        else:
                            stats["jump2"] += 1
                            if lo2 >> 31:
                                slow = True
                            else:
                                d2, de = hi2 & 0xFF, lo2 & 0xFFFFFF
                                if ms:
                                    sample = int(self.msamp[k2][1 if down else 0])
                                if self.num_docs:
                                    doc = int(self.mdoc[k2][1 if down else 0])
                                if not down:
                                    k, off = de, d2
                                elif d2:
                                    k, off = de, d2 - 1
                                else:
                                    k, off = de - 1, LAST
                                length = 0
            if slow:
                stats["slow"] += 1
                k, off, length, sample = self.slow_step(k, off, c, length, sample, doc)
                doc = self.last_doc
            else:
                stats["fast"] += 1
            out[gi], ptr[gi], self.docs[gi] = length, sample, doc
        return out, (ptr if ms else None), stats

    def walk_rows(self, read: bytes, **kw):
        """Whichever row table the image holds."""
        if self.off_mrows8:
            return self.walk_move8(read, **{k: v for k, v in kw.items() if k in ("probes",)})
        if self.off_mrows32:
            return self.walk_move32(read, **{k: v for k, v in kw.items() if k in ("len_bits",)})
        return self.walk_move(read, **{k: v for k, v in kw.items() if k in ("len_bits", "dist_far")})

    def walk_move(self, read: bytes, len_bits=20, dist_far=31):
        """Returns (lengths, pointers or None, stats)."""
        assert self.num_fast and not self.off_mrows8 and not self.off_mrows32
        ms = self.mode == 1
        LAST, lmask = 0xFFFFFFFF, (1 << len_bits) - 1
        m, R = len(read), self.R
        k, off, length, sample = R - 1, LAST, 0, self.last_run_sample
        doc = self.doc_init if self.num_docs else 0
        out, ptr, self.docs = [0] * m, [0] * m, [0] * m
        stats = {"fast": 0, "norm_iters": 0, "slow": 0, "jump2": 0, "end": 0}
        for gi in range(m - 1, -1, -1):
            c = read[gi]
            slow = False
            while True:  # STEP iterations (normalisation walks to the next row)
                x, y, z, w = (int(v) for v in self.rows[k])
                if (w >> 30) & 1:
                    slow = True
                    break
                ln = y & lmask
                if off == LAST:
                    off = ln - 1
                if off >= ln:
                    off -= ln
                    k += 1
                    stats["norm_iters"] += 1
                    continue
                break
            if not slow:
                h, s = (w >> 27) & 7, self.fid(c)
                if s >= 4:
                    slow = True
                elif s == h:
                    length += 1
                    sample = (sample - 1) & U64
                    k, off = x, off + (z & lmask)
                else:
                    slot = s - (1 if s > h else 0)
                    v = (y >> 20) if slot == 0 else ((z >> 20) if slot == 1 else (w & 0xFFF))
                    d = (w >> (12 + 5 * slot)) & 31
                    if v == 0xFFE or d == dist_far:
                        slow = True
                    elif d == 0:
                        k, off, length, sample = self.mv_end_k[s], self.mv_end_off[s], 0, self.mv_end_sl[s]
                        if self.num_docs:
                            doc = self.mv_end_doc[s]
                        stats["end"] += 1
                    else:
                        down = v == 0xFFF or off < v
                        k2 = k + d
                        x2, y2, z2, w2 = (int(t) for t in self.rows[k2])
                        stats["jump2"] += 1
                        if (w2 >> 30) & 1:
                            slow = True
                        else:
                            d2 = z2 & lmask
                            if ms:
                                sample = int(self.msamp[k2][1 if down else 0])
                            if self.num_docs:
                                doc = int(self.mdoc[k2][1 if down else 0])
                            if not down:
                                k, off = x2, d2
                            elif d2:
                                k, off = x2, d2 - 1
                            else:
                                k, off = x2 - 1, LAST
                            length = 0
            if slow:
                stats["slow"] += 1
                k, off, length, sample = self.slow_step(k, off, c, length, sample, doc)
                doc = self.last_doc
            else:
                stats["fast"] += 1
            out[gi], ptr[gi], self.docs[gi] = length, sample, doc
        return out, (ptr if ms else None), stats

    def walk_move32(self, read: bytes, len_bits=20):
        """Full 32-byte rows (spb_move_kernel<RW = 32>, PML without documents): a jump lands from the row itself.
        Returns (lengths, None, stats)."""
        assert self.num_fast and self.off_mrows32 and self.mode == 0
        LAST, lmask = 0xFFFFFFFF, (1 << len_bits) - 1
        m, R = len(read), self.R
        k, off, length = R - 1, LAST, 0
        out = [0] * m
        self.docs = [0] * m
        stats = {"fast": 0, "norm_iters": 0, "slow": 0, "jump2": 0, "end": 0}
        for gi in range(m - 1, -1, -1):
            c = read[gi]
            slow = False
            while True:
                w = [int(v) for v in self.rows32[k]]
                if (w[3] >> 15) & 1:
                    slow = True
                    break
                ln = w[1] & lmask
                if off == LAST:
                    off = ln - 1
                if off >= ln:
                    off -= ln
                    k += 1
                    stats["norm_iters"] += 1
                    continue
                break
            if not slow:
                h, s = (w[3] >> 12) & 7, self.fid(c)
                if s >= 4:
                    slow = True
                elif s == h:
                    length += 1
                    k, off = w[0], off + (w[2] & lmask)
                else:
                    slot = s - (1 if s > h else 0)
                    v = (w[1] >> 20) if slot == 0 else ((w[2] >> 20) if slot == 1 else (w[3] & 0xFFF))
                    if v == 0xFFE:
                        slow = True
                    else:
                        down = v == 0xFFF or off < v
                        de = w[4 + slot]
                        d2 = (w[3] >> 16) if slot == 0 else ((w[7] & 0xFFFF) if slot == 1 else (w[7] >> 16))
                        stats["jump2"] += 1
                        if not down:
                            k, off = de, d2
                        elif d2:
                            k, off = de, d2 - 1
                        else:
                            k, off = de - 1, LAST
                        length = 0
            if slow:
                stats["slow"] += 1
                k, off, length, _ = self.slow_step(k, off, c, length, 0, 0)
            else:
                stats["fast"] += 1
            out[gi] = length
        return out, None, stats
