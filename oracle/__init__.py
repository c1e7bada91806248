"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle (oracle/spumoni_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package ``spumoni_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
CLI_PATH = os.path.join(_HERE, "_build", "spumoni_oracle")
_lib = None


def build(force: bool = False) -> None:
    if force or not (os.path.exists(LIB_PATH) and os.path.exists(CLI_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.orc_last_error.restype = ctypes.c_char_p
        L.orc_open.restype = ctypes.c_void_p
        L.orc_open.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.orc_close.argtypes = [ctypes.c_void_p]
        L.orc_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_query.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_query_batch.restype = ctypes.c_double
        L.orc_query_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_load_doc.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_query_batch_doc.restype = ctypes.c_double
        L.orc_query_batch_doc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_nulldb_generate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_run.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_nulldb_percentile.restype = ctypes.c_double
        L.orc_nulldb_percentile.argtypes = [ctypes.c_char_p]
        _lib = L
    return _lib


class OracleIndex:
    """Reference-shaped index object: ``pml_t`` (mode 'pml') or ``ms_t`` (mode 'ms')."""

    def __init__(self, ref_file: str, mode: str = "pml", use_doc: bool = False):
        self.mode = mode
        self._h = lib().orc_open(ref_file.encode(), 1 if mode == "ms" else 0)
        if not self._h:
            raise RuntimeError(lib().orc_last_error().decode())
        if use_doc and lib().orc_load_doc(self._h, ref_file.encode()):  # pml_t/ms_t(filename, use_doc=true)
            raise RuntimeError(lib().orc_last_error().decode())

    def close(self):
        if self._h:
            lib().orc_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        n, r = ctypes.c_uint64(), ctypes.c_uint64()
        lib().orc_stats(self._h, ctypes.byref(n), ctypes.byref(r))
        return n.value, r.value

    def matching_statistics(self, read: bytes):
        m = len(read)
        L = np.zeros(m, dtype=np.uint64)
        P = np.zeros(m, dtype=np.uint64)
        lib().orc_query(self._h, read, m, L.ctypes.data, P.ctypes.data if self.mode == "ms" else None)
        return (L, P) if self.mode == "ms" else L

    def query_batch(self, reads: np.ndarray, offsets: np.ndarray, threads: int = 1, want_output: bool = True):
        """reads: uint8 concatenation; offsets: uint64[n+1]. Returns (seconds, lengths, pointers)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        total = int(offsets[-1])
        L = np.zeros(total, dtype=np.uint64) if want_output else None
        P = np.zeros(total, dtype=np.uint64) if (want_output and self.mode == "ms") else None
        sec = lib().orc_query_batch(self._h, reads.ctypes.data, offsets.ctypes.data, n,
                                    L.ctypes.data if L is not None else None,
                                    P.ctypes.data if P is not None else None, threads)
        return sec, L, P


    def nulldb_generate(self, reads: np.ndarray, offsets: np.ndarray):
        """EmpNullDatabase(ref, null_reads): returns (null_stats, mean_null_stat, percentile_value)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        stats = np.zeros(int(offsets[-1]), dtype=np.uint64)
        summary = np.zeros(2, dtype=np.float64)
        lib().orc_nulldb_generate(self._h, reads.ctypes.data, offsets.ctypes.data, len(offsets) - 1, stats.ctypes.data, summary.ctypes.data)
        return stats, float(summary[0]), float(summary[1])

    def query_batch_doc(self, reads: np.ndarray, offsets: np.ndarray, threads: int = 1):
        """Document-number variant (index opened with use_doc). Returns (lengths, pointers or None, doc_nums)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n, total = len(offsets) - 1, int(offsets[-1])
        L, D = np.zeros(total, dtype=np.uint64), np.zeros(total, dtype=np.uint64)
        P = np.zeros(total, dtype=np.uint64) if self.mode == "ms" else None
        sec = lib().orc_query_batch_doc(self._h, reads.ctypes.data, offsets.ctypes.data, n, L.ctypes.data,
                                        P.ctypes.data if P is not None else None, D.ctypes.data, threads)
        if sec < 0:
            raise RuntimeError("oracle index was opened without its document array")
        return L, P, D


def run(ref_file: str, pattern_file: str, mode: str = "pml", report: bool = False, bin_width: int = 150,
        threads: int = 1, use_doc: bool = False):
    """File-to-file `spumoni run` restatement; returns (num_reads, seconds)."""
    nr, sec = ctypes.c_uint64(), ctypes.c_double()
    rc = lib().orc_run(ref_file.encode(), pattern_file.encode(), 1 if mode == "ms" else 0, int(report) | (2 if use_doc else 0), bin_width,
                       threads, ctypes.byref(nr), ctypes.byref(sec))
    if rc:
        raise RuntimeError(lib().orc_last_error().decode())
    return nr.value, sec.value


def nulldb_percentile(path: str) -> float:
    return lib().orc_nulldb_percentile(path.encode())
