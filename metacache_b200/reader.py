"""Python mirror of the reference's `sequence_pair_reader` (sequence_io.hpp:123-190) over the C ABI
(`mcb200_reader_*`, csrc/reader.cpp), plus the multi-threaded file -> batch slots -> top hits loop that
replaces the reader thread + worker loop of `query_batched` (database_query.hpp:170-303)."""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import List, Optional, Tuple

import numpy as np

from ._lib import Mcb200Error, check, lib


class SequenceReader:
    """filename2 None/'' = unpaired; == filename1 = pairs of consecutive sequences (-pairseq); else
    two files in lockstep (-pairfiles).  `byte_range=(begin, end)`: only the records that start in that
    range of an uncompressed, unpaired file."""

    def __init__(self, filename1: str, filename2: Optional[str] = None,
                 byte_range: Optional[Tuple[int, int]] = None):
        L = lib()
        if byte_range is not None:
            if filename2:
                raise ValueError("byte ranges are for unpaired files")
            self._h = L.mcb200_reader_open_range(filename1.encode(), int(byte_range[0]), int(byte_range[1]))
        else:
            self._h = L.mcb200_reader_open(filename1.encode(), (filename2 or "").encode())
        if not self._h:
            raise Mcb200Error(-1, L.mcb200_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().mcb200_reader_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:                      # interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def index(self) -> int:
        return lib().mcb200_reader_index(self._h)

    def next(self):
        """-> (header, seq1, seq2) as bytes, or None at the end"""
        h, s1, s2 = C.c_char_p(), C.c_void_p(), C.c_void_p()
        hl, l1, l2 = C.c_uint64(), C.c_uint64(), C.c_uint64()
        hp = C.c_void_p()
        rc = check(lib().mcb200_reader_next(self._h, C.byref(hp), C.byref(hl), C.byref(s1), C.byref(l1),
                                            C.byref(s2), C.byref(l2)))
        if rc == 0:
            return None
        return (C.string_at(hp, hl.value) if hl.value else b"", C.string_at(s1, l1.value) if l1.value else b"",
                C.string_at(s2, l2.value) if l2.value else b"")

    def skip(self, n: int):
        """sequence_pair_reader::skip: -> (queries skipped, their bases)"""
        nb = C.c_uint64()
        k = check(lib().mcb200_reader_skip(self._h, n, C.byref(nb)))
        return k, nb.value

    def __iter__(self):
        while True:
            r = self.next()
            if r is None:
                return
            yield r

    def fill_batch(self, batch, slot: int, insert_size_max: int, winstride: int, max_reads: int,
                   keep_headers: bool = False, header_bytes: int = 1 << 24):
        """add_paired_read for up to max_reads parsed queries; -> (n added, [headers] or None)"""
        if keep_headers:
            hb = C.create_string_buffer(header_bytes)
            ho = (C.c_uint64 * (max_reads + 1))()
            n = check(lib().mcb200_reader_fill_batch(self._h, batch._h, slot, insert_size_max, winstride, max_reads,
                                                     hb, header_bytes, ho))
            raw = hb.raw
            return n, [raw[ho[i]:ho[i + 1]] for i in range(n)]
        n = check(lib().mcb200_reader_fill_batch(self._h, batch._h, slot, insert_size_max, winstride, max_reads,
                                                 None, 0, None))
        return n, None


def query_file(db, filename1: str, filename2: Optional[str] = None, sketching=None, threads: int = 4,
               batch_queries: int = 1 << 18, max_candidates: int = 2, insert_size_max: int = 0,
               keep_headers: bool = False, timing: Optional[dict] = None):
    """`metacache query <db> <file>` below the printing layer: every host thread owns a reader (a byte
    range of the file when it is an uncompressed unpaired file, else one reader for all) and a batch
    slot: parse -> submit -> wait -> collect.  Returns (top candidates [n, max_candidates, 4] uint32 in
    file order, headers or None)."""
    from .database import QueryBatch
    import time
    t_begin = time.perf_counter()
    sk = sketching or db.target_sketching()
    paired = bool(filename2)
    ranged = not paired and threads > 1 and not _is_gz(filename1)
    size = os.path.getsize(filename1)
    if ranged:
        cuts = [size * i // threads for i in range(threads + 1)]
        readers = [SequenceReader(filename1, byte_range=(cuts[i], cuts[i + 1])) for i in range(threads)]
    else:
        threads = 1
        readers = [SequenceReader(filename1, filename2)]
    avg_guess = 1024
    qb = QueryBatch(db, batch_queries, max(1 << 22, batch_queries * avg_guess // 4), max_candidates, False, threads)
    parts: List[list] = [[] for _ in range(threads)]
    heads: List[list] = [[] for _ in range(threads)]
    errors: List[BaseException] = []
    schedule = threading.Lock()               # database_query.hpp:87-124: query_gpu_async runs under scheduleMtx

    def work(t):
        try:
            hd = qb.host_data(t)
            while True:
                n, hs = readers[t].fill_batch(qb, t, insert_size_max, sk.winstride, batch_queries, keep_headers)
                if n == 0:
                    break
                with schedule:
                    db.query_gpu_async(qb, t, sk)
                hd.wait_for_results()
                parts[t].append(hd.top_candidates_array().copy())
                if keep_headers:
                    heads[t].extend(hs)
                hd.clear()
        except BaseException as ex:            # surfaced by the caller
            errors.append(ex)

    t_run = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    if timing is not None:
        timing["setup_s"] = t_run - t_begin            # readers + pinned / device buffers of the slots
        timing["run_s"] = time.perf_counter() - t_run   # parse + H2D + kernels + D2H, all threads
    for r in readers:
        r.close()
    qb.close()
    if errors:
        raise errors[0]
    tops = [a for p in parts for a in p]
    top = np.concatenate(tops) if tops else np.zeros((0, max_candidates, 4), np.uint32)
    return top, ([h for hs in heads for h in hs] if keep_headers else None)


def _is_gz(path: str) -> bool:
    with open(path, "rb") as f:
        return f.read(2) == b"\x1f\x8b"
