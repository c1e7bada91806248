"""spumoni_b200 -- B200-native MS/PML hot path of SPUMONI, host-side mirror in Python.

The product is ``lib/libspumoni_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/spumoni_b200.h``).  This module is a thin ctypes mirror of that ABI shaped like the
reference's own objects (``pml_t`` / ``ms_t`` with ``matching_statistics`` and
``get_bwt_stats``, /root/reference/src/compute_ms_pml.cpp:694-838) so that tests read like
reference call sites.  There is no CPU fallback: if the shared library is missing, importing
the bindings raises; if no GPU is present, queries raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

__all__ = ["pml_t", "ms_t", "Index", "EmpNullDatabase", "generate_null_database", "SpumoniError", "load_library", "build_library",
           "max_value_threshold", "kernel_launches", "device_count", "BinStats", "QueryOpts", "LIB_PATH", "CLI_PATH"]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPUMONI_B200_LIB") or os.path.join(_HERE, "lib", "libspumoni_b200.so")
SATTEST_LIB_PATH = os.path.join(_HERE, "lib", "libspumoni_b200_sattest.so")  # test-only build, see csrc/Makefile
CLI_PATH = os.path.join(_HERE, "bin", "spumoni")
MODE_PML, MODE_MS = 0, 1
QUERY_TOUPPER = 1
QUERY_POINTERS32 = 2
OPEN_DOC = 1
ERR_OVERFLOW = 7
LENGTH_BITS = 0x80  # length_bytes value: one bit per base (PML)
_LEN_DTYPES = {4: np.uint32, 2: np.uint16, 1: np.uint8}

BinStats = np.dtype([("sum_max", np.uint64), ("above", np.uint32), ("below", np.uint32)])


class QueryOpts(ctypes.Structure):
    _fields_ = [("bin_width", ctypes.c_uint32), ("lanes_per_read", ctypes.c_uint32),
                ("max_value_thr", ctypes.c_uint64), ("flags", ctypes.c_uint32), ("length_bytes", ctypes.c_uint32),
                ("saturated", ctypes.c_void_p)]


class _BuildParts(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint64), ("r", ctypes.c_uint64), ("heads", ctypes.POINTER(ctypes.c_uint8)),
                ("lens", ctypes.POINTER(ctypes.c_uint64)), ("thr", ctypes.POINTER(ctypes.c_uint64)),
                ("ssa", ctypes.POINTER(ctypes.c_uint64)), ("esa", ctypes.POINTER(ctypes.c_uint64))]


class _NullDb(ctypes.Structure):
    _fields_ = [("num_values", ctypes.c_uint64), ("ks_stat_threshold", ctypes.c_double),
                ("mean_null_stat", ctypes.c_double), ("percentile_value", ctypes.c_double)]


class SpumoniError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"spumoni_b200 error {code}: {message}")
        self.code = code


_lib: Optional[ctypes.CDLL] = None


def build_library(force: bool = False) -> None:
    """Compile libspumoni_b200.so + the `spumoni` CLI for sm_100a (nvcc; no GPU needed)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s"]
    if force:
        args.append("-B")
    subprocess.check_call(args)


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(spumoni_b200 has no fallback path)")
    L = ctypes.CDLL(LIB_PATH)
    vp, u64, i32, cp = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_char_p
    L.spb_last_error.restype = cp
    L.spb_abi_version.restype = i32
    L.spb_index_open.argtypes = [cp, i32, ctypes.POINTER(vp)]
    L.spb_index_open_ex.argtypes = [cp, i32, ctypes.c_uint32, ctypes.POINTER(vp)]
    L.spb_index_num_docs.argtypes = [vp]
    L.spb_index_num_docs.restype = ctypes.c_uint32
    L.spb_pml_query_host_doc.argtypes = [vp, vp, vp, u64, vp, vp, vp, ctypes.POINTER(QueryOpts)]
    L.spb_ms_query_host_doc.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, ctypes.POINTER(QueryOpts)]
    L.spb_pml_query_device_doc.argtypes = [vp, vp, vp, u64, vp, vp, vp, ctypes.POINTER(QueryOpts), vp]
    L.spb_ms_query_device_doc.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, ctypes.POINTER(QueryOpts), vp]
    L.spb_index_stats.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.spb_index_flat.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(u64)]
    L.spb_index_from_flat.argtypes = [vp, u64, ctypes.POINTER(vp)]
    L.spb_flat_header_bytes.restype = u64
    L.spb_index_from_header.argtypes = [vp, u64, ctypes.POINTER(vp)]
    L.spb_index_to_device.argtypes = [vp, i32]
    L.spb_index_attach_device.argtypes = [vp, i32, vp, u64]
    L.spb_index_close.argtypes = [vp]
    L.spb_index_close.restype = None
    L.spb_pml_query_host.argtypes = [vp, vp, vp, u64, vp, vp, ctypes.POINTER(QueryOpts)]
    L.spb_ms_query_host.argtypes = [vp, vp, vp, u64, vp, vp, vp, ctypes.POINTER(QueryOpts)]
    L.spb_pml_query_device.argtypes = [vp, vp, vp, u64, vp, vp, ctypes.POINTER(QueryOpts), vp]
    L.spb_ms_query_device.argtypes = [vp, vp, vp, u64, vp, vp, vp, ctypes.POINTER(QueryOpts), vp]
    L.spb_pml_query_host_text.argtypes = [vp, vp, vp, u64, vp, vp, vp, u64, ctypes.POINTER(u64), vp, ctypes.POINTER(QueryOpts)]
    L.spb_pml_query_host_text_on.argtypes = [vp, i32, vp, vp, u64, vp, vp, vp, u64, ctypes.POINTER(u64), vp, ctypes.POINTER(QueryOpts)]
    L.spb_nulldb_load.argtypes = [cp, ctypes.POINTER(_NullDb)]
    L.spb_nulldb_generate_host.argtypes = [vp, vp, vp, u64, vp, ctypes.POINTER(_NullDb)]
    L.spb_nulldb_save.argtypes = [cp, ctypes.POINTER(_NullDb), vp]
    L.spb_max_value_threshold.argtypes = [ctypes.c_double, i32, i32, i32]
    L.spb_max_value_threshold.restype = u64
    L.spb_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.spb_host_free.argtypes = [vp]
    L.spb_host_free.restype = None
    L.spb_device_count.restype = i32
    L.spb_kernel_launches.restype = u64
    L.spb_build_parts_device.argtypes = [vp, u64, i32, ctypes.POINTER(_BuildParts)]
    L.spb_build_parts_free.argtypes = [ctypes.POINTER(_BuildParts)]
    L.spb_build_parts_free.restype = None
    L.spb_build_parts_write.argtypes = [cp, ctypes.POINTER(_BuildParts)]
    L.spb_index_verify.argtypes = [cp, i32]
    L.spb_index_replicate.argtypes = [vp, ctypes.POINTER(i32), i32]
    L.spb_index_device_count.argtypes = [vp]
    L.spb_replicate_method.restype = cp
    L.spb_pml_query_host_packed.argtypes = [vp, vp, vp, vp, u64, vp, u64, vp, vp, ctypes.POINTER(QueryOpts)]
    L.spb_bits_bytes.argtypes = [u64]
    L.spb_bits_bytes.restype = u64
    L.spb_packed_bytes.argtypes = [u64]
    L.spb_packed_bytes.restype = u64
    L.spb_pack_reads.argtypes = [vp, u64, ctypes.c_uint32, vp, vp, vp, u64, ctypes.POINTER(u64)]
    L.spb_pml_expand_bits.argtypes = [vp, vp, u64, ctypes.c_uint32, vp]
    L.spb_pml_widen_lengths.argtypes = [vp, ctypes.c_uint32, vp, u64, vp]
    L.spb_shard_by_bases.argtypes = [vp, u64, ctypes.c_uint32, vp]
    L.spb_plan_chunks.argtypes = [vp, u64, u64, u64, u64, vp, u64, ctypes.POINTER(u64)]
    L.spb_plan_order.argtypes = [vp, u64, vp, ctypes.POINTER(i32)]
    _lib = L
    return L


def _check(rc: int, saturated=None) -> None:
    if rc != 0:
        e = SpumoniError(rc, load_library().spb_last_error().decode(errors="replace"))
        e.saturated = int(saturated.value) if saturated is not None else 0
        raise e


def max_value_threshold(percentile_value: float, is_pml: bool, use_dna_letters=False, use_promotions=False) -> int:
    """compute_ms_pml.cpp:871-875 (PML) / :1061-1063 (MS)."""
    return int(load_library().spb_max_value_threshold(percentile_value, int(is_pml), int(use_dna_letters),
                                                      int(use_promotions)))


def flat_header_bytes() -> int:
    return int(load_library().spb_flat_header_bytes())


def kernel_launches() -> int:
    return int(load_library().spb_kernel_launches())


def device_count() -> int:
    return int(load_library().spb_device_count())


class EmpNullDatabase:
    """EmpNullDatabase::load (emp_null_database.cpp:104-110)."""

    def __init__(self):
        self.num_values = 0
        self.ks_stat_threshold = 0.0
        self.mean_null_stat = 0.0
        self.percentile_value = 0.0

    def load(self, path: str) -> "EmpNullDatabase":
        db = _NullDb()
        _check(load_library().spb_nulldb_load(path.encode(), ctypes.byref(db)))
        self.num_values, self.ks_stat_threshold = db.num_values, db.ks_stat_threshold
        self.mean_null_stat, self.percentile_value = db.mean_null_stat, db.percentile_value
        return self


def generate_null_database(index: "Index", reads: np.ndarray, offsets: np.ndarray, out_path: Optional[str] = None,
                           ks_stat_threshold: float = 0.0):
    """EmpNullDatabase(ref, null_reads, ...) through the GPU path (emp_null_database.cpp:19-80 with
    compute_ms_pml.cpp:1410-1486): returns (EmpNullDatabase, null_stats); writes the .pmlnulldb/.msnulldb file if asked."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    stats = np.zeros(int(offsets[-1]), dtype=np.uint32)
    db = _NullDb()
    _check(load_library().spb_nulldb_generate_host(index._h, reads.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                                   stats.ctypes.data, ctypes.byref(db)))
    db.ks_stat_threshold = ks_stat_threshold
    if out_path:
        _check(load_library().spb_nulldb_save(out_path.encode(), ctypes.byref(db), stats.ctypes.data))
    out = EmpNullDatabase()
    out.num_values, out.ks_stat_threshold = db.num_values, db.ks_stat_threshold
    out.mean_null_stat, out.percentile_value = db.mean_null_stat, db.percentile_value
    return out, stats


def build_parts_device(text: np.ndarray, device: int = 0, write_to: Optional[str] = None) -> dict:
    """spb_build_parts_device: run heads / lengths / thresholds / SA samples of the BWT of `text` (uint8, ending with its
    terminator 0), computed on the GPU -- what `spumoni build` gets from Big-BWT + pfp_thresholds.  write_to = <ref_file>:
    also writes <ref_file>.bwt.heads / .bwt.len / .thr_pos / .ssa / .esa."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    lib = load_library()
    P = _BuildParts()
    _check(lib.spb_build_parts_device(text.ctypes.data, text.size, device, ctypes.byref(P)))
    try:
        r = int(P.r)
        out = {"n": int(P.n), "r": r, "heads": np.ctypeslib.as_array(P.heads, shape=(r,)).copy()}
        for k in ("lens", "thr", "ssa", "esa"):
            out[k] = np.ctypeslib.as_array(getattr(P, k), shape=(r,)).copy()
        if write_to:
            _check(lib.spb_build_parts_write(write_to.encode(), ctypes.byref(P)))
    finally:
        lib.spb_build_parts_free(ctypes.byref(P))
    return out


def verify_index(ref_file: str, mode: int) -> None:
    """spb_index_verify: deep self-test of an index file (re-derives the stored select / rank supports); raises with the field name."""
    _check(load_library().spb_index_verify(ref_file.encode(), mode))


def bits_bytes(n_bases: int) -> int:
    return int(load_library().spb_bits_bytes(n_bases))


def pack_reads_2bit(reads: np.ndarray, toupper: bool = False):
    """spb_pack_reads: concatenated reads -> (packed uint8[spb_packed_bytes], exc_pos uint64[], exc_val uint8[])."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    lib = load_library()
    packed = np.zeros(int(lib.spb_packed_bytes(reads.size)), dtype=np.uint8)
    cap = max(1024, reads.size // 64)
    while True:
        pos, val, n = np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint8), ctypes.c_uint64(0)
        rc = lib.spb_pack_reads(reads.ctypes.data, reads.size, QUERY_TOUPPER if toupper else 0, packed.ctypes.data,
                                pos.ctypes.data, val.ctypes.data, cap, ctypes.byref(n))
        if rc == 6 and n.value > cap:  # SPB_ERR_LIMIT: the exception list did not fit
            cap = int(n.value)
            continue
        _check(rc)
        return packed, pos[: n.value].copy(), val[: n.value].copy()


def expand_bits(bits: np.ndarray, offsets: np.ndarray, length_bytes: int = 4) -> np.ndarray:
    """spb_pml_expand_bits: bit map (LENGTH_BITS output) -> PML lengths."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.zeros(int(offsets[-1]) if len(offsets) else 0, dtype=_LEN_DTYPES[length_bytes])
    _check(load_library().spb_pml_expand_bits(bits.ctypes.data, offsets.ctypes.data, len(offsets) - 1, length_bytes, out.ctypes.data))
    return out


def widen_lengths(narrow: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """spb_pml_widen_lengths: clamped uint8 / uint16 PML lengths -> exact uint32 lengths."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    narrow = np.ascontiguousarray(narrow)
    out = np.zeros(int(offsets[-1]) if len(offsets) else 0, dtype=np.uint32)
    _check(load_library().spb_pml_widen_lengths(narrow.ctypes.data, narrow.dtype.itemsize, offsets.ctypes.data, len(offsets) - 1, out.ctypes.data))
    return out


def shard_by_bases(offsets: np.ndarray, n_shards: int) -> np.ndarray:
    """spb_shard_by_bases: cuts[s] .. cuts[s+1] = contiguous read range of shard s, balanced by bases."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    cuts = np.zeros(n_shards + 1, dtype=np.uint64)
    _check(load_library().spb_shard_by_bases(offsets.ctypes.data, len(offsets) - 1, n_shards, cuts.ctypes.data))
    return cuts


def plan_chunks(offsets: np.ndarray, want_reads: int, min_bases: int, max_bases: int) -> np.ndarray:
    """Sub-batch boundaries (first read of every sub-batch, then n_reads) of a host call -- test hook."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    out, cnt = np.zeros(n + 2, dtype=np.uint64), ctypes.c_uint64(0)
    _check(load_library().spb_plan_chunks(offsets.ctypes.data, n, want_reads, min_bases, max_bases, out.ctypes.data, len(out), ctypes.byref(cnt)))
    return out[: cnt.value + 1].copy()


def plan_order(offsets: np.ndarray):
    """Longest-first processing order of a sub-batch, or None when its lengths are uniform -- test hook."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    order, flag = np.zeros(max(n, 1), dtype=np.uint32), ctypes.c_int(0)
    _check(load_library().spb_plan_order(offsets.ctypes.data, n, order.ctypes.data, ctypes.byref(flag)))
    return order[:n].copy() if flag.value else None


def pack_reads(reads) -> Tuple[np.ndarray, np.ndarray]:
    """list of bytes -> (concatenated uint8 array, uint64 offsets[n+1])."""
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    if len(reads):
        off[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    buf = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if len(reads) else np.zeros(0, dtype=np.uint8)
    return buf, off


class _OwnedArray(np.ndarray):
    _owner = None


class Index:
    """Owning handle on a flat index (``spb_index``)."""

    def __init__(self, handle: int, mode: int):
        self._h = ctypes.c_void_p(handle)
        self.mode = mode
        self._keepalive = None

    # ---- construction
    @classmethod
    def open(cls, ref_file: str, mode: int, use_doc: bool = False) -> "Index":
        """use_doc: also load <ref_file>.doc (the reference's pml_t/ms_t(filename, use_doc=true))."""
        h = ctypes.c_void_p()
        _check(load_library().spb_index_open_ex(ref_file.encode(), mode, OPEN_DOC if use_doc else 0, ctypes.byref(h)))
        return cls(h.value, mode)

    def num_docs(self) -> int:
        return int(load_library().spb_index_num_docs(self._h))

    @classmethod
    def from_flat(cls, blob: np.ndarray) -> "Index":
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        h = ctypes.c_void_p()
        _check(load_library().spb_index_from_flat(blob.ctypes.data, blob.nbytes, ctypes.byref(h)))
        mode = int(np.frombuffer(blob[12:16].tobytes(), dtype=np.uint32)[0])
        return cls(h.value, mode)

    @classmethod
    def from_header(cls, header: np.ndarray) -> "Index":
        """Header-only index (no host image); pair with attach_device."""
        header = np.ascontiguousarray(header, dtype=np.uint8)
        h = ctypes.c_void_p()
        _check(load_library().spb_index_from_header(header.ctypes.data, header.nbytes, ctypes.byref(h)))
        mode = int(np.frombuffer(header[12:16].tobytes(), dtype=np.uint32)[0])
        return cls(h.value, mode)

    def close(self) -> None:
        if self._h:
            load_library().spb_index_close(self._h)
            self._h = ctypes.c_void_p()
            self._keepalive = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference-shaped accessors
    def get_bwt_stats(self) -> Tuple[int, int]:
        n, r = ctypes.c_uint64(), ctypes.c_uint64()
        _check(load_library().spb_index_stats(self._h, ctypes.byref(n), ctypes.byref(r)))
        return n.value, r.value

    def flat(self) -> np.ndarray:
        """The flat image as a read-only uint8 view (valid while the index is open)."""
        p, b = ctypes.c_void_p(), ctypes.c_uint64()
        _check(load_library().spb_index_flat(self._h, ctypes.byref(p), ctypes.byref(b)))
        arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(b.value,))
        arr.flags.writeable = False
        view = arr.view(_OwnedArray)
        view._owner = self  # the blob lives inside the index: keep it alive with the view
        return view

    # ---- residency
    def to_device(self, device: int = 0) -> "Index":
        _check(load_library().spb_index_to_device(self._h, device))
        return self

    def replicate(self, devices) -> "Index":
        """Resident on several GPUs (one broadcast of the flat image); host queries are then sharded over them by bases."""
        arr = (ctypes.c_int * len(devices))(*devices)
        _check(load_library().spb_index_replicate(self._h, arr, len(devices)))
        return self

    def device_count(self) -> int:
        return int(load_library().spb_index_device_count(self._h))

    def attach_device(self, device: int, dev_tensor) -> "Index":
        """Adopt a torch uint8 CUDA tensor already holding the flat image (e.g. after a broadcast)."""
        _check(load_library().spb_index_attach_device(self._h, device, dev_tensor.data_ptr(), dev_tensor.numel()))
        self._keepalive = dev_tensor
        return self

    # ---- queries (host buffers)
    def _opts(self, bin_width=0, max_value_thr=0, lanes=0, toupper=False, length_bytes=4, saturated=None, pointer_bytes=8) -> QueryOpts:
        o = QueryOpts()
        o.bin_width, o.lanes_per_read, o.max_value_thr = bin_width, lanes, max_value_thr
        o.flags = (QUERY_TOUPPER if toupper else 0) | (QUERY_POINTERS32 if pointer_bytes == 4 else 0)
        o.length_bytes = length_bytes
        o.saturated = saturated
        return o

    def query_host_text(self, reads: np.ndarray, offsets: np.ndarray, ids, bin_width=0, max_value_thr=0, toupper=False,
                        capacity: Optional[int] = None):
        """PML query + on-device decimal formatting: returns (the bytes of the .pseudo_lengths file for these reads, stats)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        idbuf, idoff = pack_reads([bytes(i) for i in ids])
        cap = capacity if capacity is not None else int(offsets[-1]) * 11 + len(idbuf) + 3 * n + 64
        text = np.zeros(cap, dtype=np.uint8)
        S = np.zeros(n, dtype=BinStats) if bin_width else None
        o = self._opts(bin_width, max_value_thr, 0, toupper, 4, None)
        got = ctypes.c_uint64(0)
        rc = load_library().spb_pml_query_host_text(self._h, reads.ctypes.data, offsets.ctypes.data, n, idbuf.ctypes.data,
                                                    idoff.ctypes.data, text.ctypes.data, cap, ctypes.byref(got),
                                                    S.ctypes.data if S is not None else None, ctypes.byref(o))
        self.last_text_bytes = int(got.value)
        _check(rc)
        return text[: got.value].tobytes(), S

    def query_host_doc(self, reads: np.ndarray, offsets: np.ndarray, bin_width=0, max_value_thr=0, lanes=0, toupper=False):
        """Document-number variant (index opened with use_doc=True). Returns (lengths, pointers or None, stats, doc_nums)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        total = int(offsets[-1]) if n >= 0 and len(offsets) else 0
        L, D = np.zeros(total, dtype=np.uint32), np.zeros(total, dtype=np.uint32)
        S = np.zeros(n, dtype=BinStats) if bin_width else None
        o = self._opts(bin_width, max_value_thr, lanes, toupper, 4, None)
        lib = load_library()
        sp_ = S.ctypes.data if S is not None else None
        if self.mode == MODE_MS:
            P = np.zeros(total, dtype=np.uint64)
            _check(lib.spb_ms_query_host_doc(self._h, reads.ctypes.data, offsets.ctypes.data, n, L.ctypes.data, P.ctypes.data,
                                             D.ctypes.data, sp_, ctypes.byref(o)))
            return L, P, S, D
        _check(lib.spb_pml_query_host_doc(self._h, reads.ctypes.data, offsets.ctypes.data, n, L.ctypes.data, D.ctypes.data,
                                          sp_, ctypes.byref(o)))
        return L, None, S, D

    def query_host(self, reads: np.ndarray, offsets: np.ndarray, bin_width=0, max_value_thr=0, lanes=0,
                   toupper=False, out_lengths=None, out_pointers=None, out_stats=None, length_bytes=4, pointer_bytes=8, packed=None):
        """length_bytes 4/2/1 selects uint32/uint16/uint8 lengths (LENGTH_BITS: a uint64 bit map, one bit per base, PML only),
        pointer_bytes 8/4 uint64/uint32 MS pointers.  Clamped PML lengths are not an error (``last_saturated`` counts them,
        widen_lengths rebuilds them); a clamped MS value raises SpumoniError with code ERR_OVERFLOW.
        packed = (packed, exc_pos, exc_val) from pack_reads_2bit: the reads travel as 2 bits per base (PML only)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8) if reads is not None else None
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        total = int(offsets[-1]) if n >= 0 and len(offsets) else 0
        if length_bytes == LENGTH_BITS:
            L = out_lengths if out_lengths is not None else np.zeros(bits_bytes(total) // 8, dtype=np.uint64)
            assert L.dtype.itemsize == 8 and L.nbytes >= bits_bytes(total)
        else:
            L = out_lengths if out_lengths is not None else np.zeros(total, dtype=_LEN_DTYPES[length_bytes])
            assert L.dtype.itemsize == length_bytes and L.size >= total
        S = None
        if bin_width:
            S = out_stats if out_stats is not None else np.zeros(n, dtype=BinStats)
        sat = ctypes.c_uint64(0)
        o = self._opts(bin_width, max_value_thr, lanes, toupper, length_bytes, ctypes.addressof(sat), pointer_bytes)
        self.last_saturated = sat  # keeps the counter alive for the duration of the call
        lib = load_library()
        if self.mode == MODE_MS:
            P = out_pointers if out_pointers is not None else np.zeros(total, dtype=np.uint64 if pointer_bytes == 8 else np.uint32)
            assert P.dtype.itemsize == pointer_bytes and P.size >= total
            _check(lib.spb_ms_query_host(self._h, reads.ctypes.data, offsets.ctypes.data, n, L.ctypes.data,
                                         P.ctypes.data, S.ctypes.data if S is not None else None, ctypes.byref(o)), sat)
            return L, P, S
        if packed is not None:
            pk, ep, ev = packed
            _check(lib.spb_pml_query_host_packed(self._h, pk.ctypes.data, ep.ctypes.data if ep.size else None, ev.ctypes.data if ev.size else None,
                                                 ep.size, offsets.ctypes.data, n, L.ctypes.data, S.ctypes.data if S is not None else None,
                                                 ctypes.byref(o)), sat)
            return L, None, S
        _check(lib.spb_pml_query_host(self._h, reads.ctypes.data, offsets.ctypes.data, n, L.ctypes.data,
                                      S.ctypes.data if S is not None else None, ctypes.byref(o)), sat)
        return L, None, S

    # ---- queries (torch CUDA tensors, asynchronous on `stream`)
    def query_device(self, d_reads, d_offsets, n_reads: int, d_lengths, d_pointers=None, d_stats=None, bin_width=0,
                     max_value_thr=0, lanes=0, toupper=False, stream: int = 0, d_saturated=None) -> None:
        """d_lengths: int32/int16/uint8 CUDA tensor (its element size selects the output width), or an int64 tensor = the
        bit map of LENGTH_BITS (zeroed by the caller); d_saturated: optional 1-element int64 CUDA tensor incremented by the
        number of clamped values."""
        o = self._opts(bin_width, max_value_thr, lanes, toupper, LENGTH_BITS if d_lengths.element_size() == 8 else d_lengths.element_size(),
                       d_saturated.data_ptr() if d_saturated is not None else None,
                       d_pointers.element_size() if d_pointers is not None else 8)
        lib = load_library()
        sp = d_stats.data_ptr() if d_stats is not None else None
        if self.mode == MODE_MS:
            _check(lib.spb_ms_query_device(self._h, d_reads.data_ptr(), d_offsets.data_ptr(), n_reads,
                                           d_lengths.data_ptr(), d_pointers.data_ptr(), sp, ctypes.byref(o), stream))
        else:
            _check(lib.spb_pml_query_device(self._h, d_reads.data_ptr(), d_offsets.data_ptr(), n_reads,
                                            d_lengths.data_ptr(), sp, ctypes.byref(o), stream))


class pml_t(Index):
    """pml_t (compute_ms_pml.cpp:694-746): loads <ref_file>.thrbv.spumoni."""

    def __init__(self, filename: str, device: Optional[int] = 0, use_doc: bool = False):
        idx = Index.open(filename, MODE_PML, use_doc)
        super().__init__(idx._h.value, MODE_PML)
        idx._h = ctypes.c_void_p()
        if device is not None:
            self.to_device(device)

    def matching_statistics(self, read: bytes, **kw) -> np.ndarray:
        buf, off = pack_reads([read])
        return self.query_host(buf, off, **kw)[0]


class ms_t(Index):
    """ms_t (compute_ms_pml.cpp:748-838): loads <ref_file>.thrbv.ms; text from <ref_file>."""

    def __init__(self, filename: str, device: Optional[int] = 0, use_doc: bool = False):
        idx = Index.open(filename, MODE_MS, use_doc)
        super().__init__(idx._h.value, MODE_MS)
        idx._h = ctypes.c_void_p()
        if device is not None:
            self.to_device(device)

    def matching_statistics(self, read: bytes, **kw) -> Tuple[np.ndarray, np.ndarray]:
        buf, off = pack_reads([read])
        L, P, _ = self.query_host(buf, off, **kw)
        return L, P
