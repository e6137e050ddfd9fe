"""ctypes binding of oracle/mc_oracle.c - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (metacache_b200/) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmc_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mc_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", _LIB, src])
    return _LIB


class SketchOpt(C.Structure):
    _fields_ = [("k", C.c_uint32), ("s", C.c_uint32), ("w", C.c_uint32), ("stride", C.c_uint32)]


class Candidate(C.Structure):
    _fields_ = [("tgt", C.c_uint32), ("hits", C.c_uint32), ("beg", C.c_uint32), ("end", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.mco_hash32.restype = C.c_uint32
        L.mco_hash32.argtypes = [C.c_uint32]
        L.mco_revcomp32.restype = C.c_uint32
        L.mco_revcomp32.argtypes = [C.c_uint32, C.c_uint32]
        L.mco_canonical32.restype = C.c_uint32
        L.mco_canonical32.argtypes = [C.c_uint32, C.c_uint32]
        L.mco_num_windows.restype = C.c_uint64
        L.mco_num_windows.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.mco_sketch_window.restype = C.c_int
        L.mco_sketch_window.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(SketchOpt), C.c_void_p]
        L.mco_sketch_sequence.restype = C.c_uint64
        L.mco_sketch_sequence.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(SketchOpt), C.c_void_p,
                                          C.c_void_p, C.c_uint64]
        L.mco_table_build.restype = C.c_void_p
        L.mco_table_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.mco_table_free.argtypes = [C.c_void_p]
        L.mco_max_windows_in_range.restype = C.c_uint32
        L.mco_max_windows_in_range.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        L.mco_candidates.restype = C.c_uint32
        L.mco_candidates.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                     C.c_void_p]
        L.mco_query.restype = C.c_uint32
        L.mco_query.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64,
                                C.POINTER(SketchOpt), C.c_uint32, C.c_uint32, C.c_void_p,
                                C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p]
        L.mco_merge_tops.restype = C.c_uint32
        L.mco_merge_tops.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.mco_classify.restype = C.c_uint32
        L.mco_classify.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float,
                                   C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def hash32(x: int) -> int:
    return lib().mco_hash32(x & 0xFFFFFFFF)


def revcomp32(x: int, k: int) -> int:
    return lib().mco_revcomp32(x & 0xFFFFFFFF, k)


def canonical32(x: int, k: int) -> int:
    return lib().mco_canonical32(x & 0xFFFFFFFF, k)


def num_windows(n: int, w: int = 127, stride: int = 112) -> int:
    return lib().mco_num_windows(n, w, stride)


def sketch_sequence(seq: bytes, k=16, s=16, w=127, stride=112):
    """-> list (one entry per window) of np.uint32 arrays, or None for windows the
    reference does not sketch (fewer than k characters)."""
    opt = SketchOpt(k, s, w, stride)
    nw = num_windows(len(seq), w, stride)
    feats = np.zeros((nw, s), dtype=np.uint32)
    counts = np.zeros(nw, dtype=np.int32)
    lib().mco_sketch_sequence(seq, len(seq), C.byref(opt), feats.ctypes.data, counts.ctypes.data, nw)
    return [None if c < 0 else feats[i, :c].copy() for i, c in enumerate(counts)]


class Table:
    """Exact-match feature -> bucket dictionary built from `.cache` arrays."""

    def __init__(self, keys, sizes, values):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint8)
        values = np.ascontiguousarray(values, dtype=np.uint64)
        assert int(sizes.sum(dtype=np.int64)) == len(values)
        self._h = lib().mco_table_build(keys.ctypes.data, sizes.ctypes.data, len(keys),
                                        values.ctypes.data, len(values))
        self.nkeys, self.nvalues = len(keys), len(values)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mco_table_free(self._h)
            self._h = None


def max_windows_in_range(len1, len2=0, insert_size_max=0, winstride=112) -> int:
    return lib().mco_max_windows_in_range(len1, len2, insert_size_max, winstride)


def candidates(locs, max_win, maxc=2, tax_of_tgt=None):
    locs = np.ascontiguousarray(locs, dtype=np.uint64)
    top = (Candidate * max(maxc, 1))()
    tp = None
    if tax_of_tgt is not None:
        tax_of_tgt = np.ascontiguousarray(tax_of_tgt, dtype=np.uint64)
        tp = tax_of_tgt.ctypes.data
    n = lib().mco_candidates(locs.ctypes.data, len(locs), max_win, maxc, tp, top)
    return [(top[i].tgt, top[i].hits, top[i].beg, top[i].end) for i in range(n)]


def query(tables, seq1: bytes, seq2: bytes = b"", k=16, s=16, w=127, stride=112, maxc=2,
          insert_size_max=0, max_win=None, tax_of_tgt=None, allhits_cap=1 << 16):
    """One read (pair) against `tables` (list of Table = database parts).
    -> (allhits np.uint64 sorted per part, [(tgt, hits, beg, end), ...])"""
    if isinstance(tables, Table):
        tables = [tables]
    opt = SketchOpt(k, s, w, stride)
    if max_win is None:
        max_win = max_windows_in_range(len(seq1), len(seq2), insert_size_max, stride)
    arr = (C.c_void_p * len(tables))(*[t._h for t in tables])
    top = (Candidate * max(maxc, 1))()
    n_all = C.c_uint64(0)
    tp = None
    if tax_of_tgt is not None:
        tax_of_tgt = np.ascontiguousarray(tax_of_tgt, dtype=np.uint64)
        tp = tax_of_tgt.ctypes.data
    while True:
        allh = np.zeros(allhits_cap, dtype=np.uint64)
        n = lib().mco_query(arr, len(tables), seq1, len(seq1), seq2, len(seq2), C.byref(opt),
                            max_win, maxc, tp, allh.ctypes.data, allhits_cap, C.byref(n_all), top)
        if n_all.value <= allhits_cap:
            break
        allhits_cap = int(n_all.value)
    return allh[:n_all.value].copy(), [(top[i].tgt, top[i].hits, top[i].beg, top[i].end)
                                       for i in range(n)]


def merge_tops(parts_tops, maxc=2):
    """Stable part-ordered merge of per-part top lists."""
    nparts = len(parts_tops)
    buf = (Candidate * (max(maxc, 1) * max(nparts, 1)))()
    cnt = np.zeros(max(nparts, 1), dtype=np.uint32)
    for p, tl in enumerate(parts_tops):
        cnt[p] = len(tl)
        for i, c in enumerate(tl):
            buf[p * maxc + i] = Candidate(*c)
    top = (Candidate * max(maxc, 1))()
    n = lib().mco_merge_tops(buf, cnt.ctypes.data, nparts, maxc, top)
    return [(top[i].tgt, top[i].hits, top[i].beg, top[i].end) for i in range(n)]


def classify(top, lineages, hits_min=5, hits_diff_fraction=1.0, lowest=0, highest=19):
    """top: [(tgt, hits, beg, end)...]; lineages: np.uint32 [n_targets, 21] (taxon ordinal + 1).
    -> (taxon ordinal + 1 or 0, rank)"""
    lineages = np.ascontiguousarray(lineages, dtype=np.uint32)
    buf = (Candidate * max(len(top), 1))(*[Candidate(*t) for t in top])
    rank = C.c_uint32(21)
    t = lib().mco_classify(buf, len(top), lineages.ctypes.data, lineages.shape[0], hits_min, hits_diff_fraction,
                           lowest, highest, C.byref(rank))
    return t, rank.value
